#!/usr/bin/env python
"""bench.py — MFCC + delta + delta-delta feature frames/s on B200 (BASELINE.json metric).

A "step" = one pass of the hot path over one shard of synthetic 16 kHz PCM:
BASELINE.json configs[2]: 10 000 utterances x 10 s, 40 mel, 13 MFCC (12 + c0) + delta + delta-delta, per-utterance
CMN, 9.98 M frames per GPU per step. One process per GPU (torchrun for N > 1), utterance-sharded, weak scaling:
every rank owns a full config-3 shard; no data-path collective. The corpus-CMVN variant (BASELINE configs[3]: one NCCL
all-reduce of 4*39+1 doubles inside the Normalizer) is timed as well and reported under "corpus_cmvn".

  value     whole-job frames/s with PCM and features resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       same metric through the C-ABI call with HOST buffers (afe_batch_run_host: H2D PCM + kernels + D2H features)
  roofline  the one kernel of the step (k_fused_mfcc): algorithmic bytes (2*S + 4*width = 476 B/frame) / its event time,
            against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU classes (oracle/_ref, FFTW-API shim) on this box's host cores, bounded sample

`--impl reference` times that CPU path alone (no CUDA) and prints the same line with "impl": "reference".
torch is used for device memory, streams, events and torch.distributed only; all compute is libafe_cuda.so.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR, SECONDS, W, S, NB = 16000, 10, 400, 160, 40
WIDTH = 39
METRIC = "mfcc_delta_delta2_frames_per_s"
UNIT = "frames/s"


def params_dict():
    return dict(window_size=W, shift=S, num_banks=NB, sample_rate=float(SR), low_freq=64.0, high_freq=SR / 2.0,
                ceps_len=12, want_c0=1, lift_coef=22.0, norm=1, dyn=2, delta_l1=3, delta_l2=3, norm_after_dyn=1, alpha=1.0)


def workload_config(n_utts, n_gpus):
    return {"workload": f"BASELINE configs[2]: synthetic 16 kHz PCM, {n_utts} utterances x {SECONDS} s per GPU, "
                        f"W=400/S=160, 512-pt FFT, 40 mel, 13 MFCC(12+c0)+delta+delta-delta (39-dim), per-utterance CMN",
            "utterances_per_gpu": n_utts, "samples_per_utterance": SR * SECONDS,
            "frames_per_gpu": n_utts * ((SR * SECONDS - (W - S)) // S),
            "sharding": f"utterance-sharded x{n_gpus}, no data-path collective",
            "l2_policy": f"inputs ({n_utts * SR * SECONDS * 2 / 1e9:.2f} GB PCM + {n_utts * ((SR * SECONDS - (W - S)) // S) * WIDTH * 4 / 1e9:.2f} GB "
                         "features per GPU) are larger than L2 (126 MB); no explicit flush"}


def synth_host(n_utts, seed):
    """SURVEY §8(d) generator on the host (used by the CPU arm)."""
    rng = np.random.default_rng(seed)
    n = SR * SECONDS
    t = np.arange(n) / SR
    out = np.empty((n_utts, n), np.int16)
    for u in range(n_utts):
        f = rng.uniform(100.0, 3800.0)
        x = 3000.0 * rng.standard_normal(n).astype(np.float32) + 8000.0 * np.sin(2 * np.pi * f * t)
        out[u] = np.clip(np.round(x), -32767, 32767).astype(np.int16)
    return out


def cpu_arm(n_utts, steps, warmup, threads, pcm=None):
    """Times the reference's own CPU path (or the port when oracle/_ref is absent). Returns dict."""
    import oracle_lib as ol
    kind = "reference" if ol.available("ref") else "port"
    lib = ol.RefLib("ref" if kind == "reference" else "port")
    p = params_dict()
    if kind == "port":
        threads = 1
    if pcm is None:
        pcm = synth_host(n_utts, 1234)
    utts = [pcm[i] for i in range(n_utts)]
    frames = n_utts * ((SR * SECONDS - (W - S)) // S)
    times = []
    for i in range(warmup + steps):
        _, s = lib.extract(p, utts, sample_limit=0, n_threads=threads)   # Q5: MfccCpu sized to the utterance
        if i >= warmup:
            times.append(s)
    t = float(np.mean(times))
    return dict(kind=kind, cores=threads, value=frames / t, seconds_per_step=t, frames_per_step=frames,
                sample=f"{n_utts} utterances x {SECONDS} s of the same synthetic workload ({frames} frames) per step, "
                       f"fresh MfccCpu per utterance sized to it (Q3/Q5), FFT = oracle/fftw_shim.c (FFTW not installed)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import oracle_lib as ol
    lib = ol.RefLib("ref" if ol.available("ref") else "port")
    cal = synth_host(8, 99)
    _, s = lib.extract(params_dict(), [cal[i] for i in range(8)], sample_limit=0, n_threads=1)
    rate1 = 8 * 998 / s
    # ~3 s of wall time per step with all cores
    n_utts = int(max(threads * 4, min(4096, 3.0 * rate1 * threads / 998)))
    r = cpu_arm(n_utts, args.steps, max(1, min(args.warmup, 2)), threads)
    cfg = workload_config(10000, args.gpus)
    cfg["reference_sample"] = r["sample"]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "audio_hours_per_s": r["value"] * S / SR / 3600.0}
    print(json.dumps(line), flush=True)


class ClockSampler:
    """nvidia-smi polled every 20 ms in the background; stop(t0, t1) keeps the samples taken inside the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc:
            self.proc.terminate()
        inside = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1 + 0.03)]
        rows = inside if len(inside) >= 2 else [r for (_, r) in self.rows]
        sm, mx, power, reasons = [], 0, [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_in_timed_region": len(inside), "power_w_max": max(power) if power else None}


_FULL_AFFINITY = None


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated: with 8 ranks
    per box the end-to-end path is bound by host memory / PCIe root-complex locality, not by the kernels."""
    global _FULL_AFFINITY
    _FULL_AFFINITY = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"bound": True, "cpus": len(cpus), "first_cpu": min(cpus)}
    except Exception as e:
        return {"bound": False, "why": str(e)[:120]}
    return {"bound": False, "why": "empty affinity mask"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import afe_loader
    afe = afe_loader.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(torch, local) if world > 1 and not args.no_numa_bind else {"bound": False, "why": "single rank"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_utts, n = args.utts, SR * SECONDS
    T = (n - (W - S)) // S
    frames = n_utts * T

    # ---- synthetic shard, generated on the device (plumbing), seed + rank
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    pcm = torch.empty((n_utts * n + 64,), dtype=torch.int16, device=dev)
    t = torch.arange(n, device=dev, dtype=torch.float32) / SR
    chunk = 500
    for u0 in range(0, n_utts, chunk):
        c = min(chunk, n_utts - u0)
        f = torch.empty((c, 1), device=dev).uniform_(100.0, 3800.0, generator=g)
        x = 3000.0 * torch.randn((c, n), device=dev, generator=g) + 8000.0 * torch.sin(2 * np.pi * f * t)
        pcm[u0 * n:(u0 + c) * n] = x.round_().clamp_(-32767, 32767).to(torch.int16).reshape(-1)
        del x
    out = torch.empty((frames, WIDTH), dtype=torch.float32, device=dev)
    offs = np.arange(n_utts, dtype=np.int64) * n
    lens = np.full(n_utts, n, np.int64)

    p = params_dict()
    ap = afe.make_params(input_buffer_size=1 << 22, **{k: v for k, v in p.items() if k != "alpha"})
    flags = (afe.BATCH_Q1_EXACT | (afe.BATCH_NO_TMA if args.no_tma else 0) | (afe.BATCH_NO_CLUSTER if args.no_cluster else 0))
    # a dedicated (non-default) stream: the library launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    b = afe.BatchMfcc(ap, local, stats_scope=afe.STATS_REFERENCE_BLOCK, flags=flags)
    b.set_stream(stream.cuda_stream)
    assert b.plan(offs, lens) == frames

    def step():
        # ONE kernel: PCM -> normalised rows (the last tile of an utterance finalises its CMN statistics and normalises
        # the utterance in place while its rows are L2 resident)
        b.run_device(pcm.data_ptr(), out.data_ptr())
        ev_mid.record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev_mid = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    t_begin = time.perf_counter()
    k1_ms, launches = [], 0
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev0[0].record(stream)
    for i in range(args.steps):
        ev_mid = mids[i]
        step()
        launches += b.kernel_launches                         # 1 per step (k_fused_mfcc)
        ev0[i + 1].record(stream)
    barrier()
    t_end = time.perf_counter()
    total_ms = ev0[0].elapsed_time(ev0[-1])
    k1_ms = [ev0[i].elapsed_time(mids[i]) for i in range(args.steps)]
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    value = world * frames / (ms_per_step * 1e-3)

    # ---- corpus-CMVN variant: extract -> corpus sums -> NCCL all-reduce (C ABI, inside the Normalizer) -> normalise
    corpus = None
    try:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_char * 128)()
            afe._check(afe.lib().afe_nccl_get_unique_id(raw))
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        if world > 1:
            idd = idbuf.to(dev); dist.broadcast(idd, 0); idbuf = idd.cpu()
        comm = C.c_void_p()
        afe._check(afe.lib().afe_nccl_comm_init(idbuf.numpy().tobytes(), world, rank, local, C.byref(comm)))
        bc = afe.BatchMfcc(ap, local, stats_scope=afe.STATS_CORPUS, flags=flags)
        bc.set_stream(stream.cuda_stream)
        bc.plan(offs, lens)

        def cstep():
            bc.extract_device(pcm.data_ptr(), out.data_ptr())
            bc.corpus_stats()
            bc.allreduce(comm.value)
            bc.normalize_device(out.data_ptr())
        for _ in range(max(1, args.warmup)):
            cstep()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(args.steps):
            cstep()
        c1.record(stream)
        barrier()
        cms = torch.tensor([c0.elapsed_time(c1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(cms, op=dist.ReduceOp.MAX)
        cms = float(cms.item()) / args.steps
        corpus = {"value": world * frames / (cms * 1e-3), "unit": UNIT, "ms_per_step": cms,
                  "collective": "ncclAllReduce x3 grouped (sum 79 | min 39 | max 39 doubles) via afe_normalizer_allreduce",
                  "algorithmic_bytes_per_frame": 2 * S + 12 * WIDTH}
        bc.close()
        afe.lib().afe_nccl_comm_destroy(comm)
    except Exception as e:  # NCCL missing etc.: the headline number does not depend on it
        corpus = {"unavailable": str(e)[:200]}

    # ---- e2e through the C-ABI with HOST (pinned) buffers
    e2e = None
    if not args.no_e2e:
        h_pcm = torch.empty((n_utts * n + 64,), dtype=torch.int16).pin_memory()
        h_pcm.copy_(pcm)
        h_out = torch.empty((frames, WIDTH), dtype=torch.float32).pin_memory()
        be = afe.BatchMfcc(ap, local, stats_scope=afe.STATS_REFERENCE_BLOCK, flags=flags)
        be.plan(offs, lens)
        lib = afe.lib()
        e2e_steps = max(2, min(args.steps, 5))
        afe._check(lib.afe_batch_run_host(be._h, C.c_void_p(h_pcm.data_ptr()), C.c_void_p(h_out.data_ptr())))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            afe._check(lib.afe_batch_run_host(be._h, C.c_void_p(h_pcm.data_ptr()), C.c_void_p(h_out.data_ptr())))
        torch.cuda.synchronize()
        et = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        e2e = {"value": world * frames / float(et.item()), "unit": UNIT, "h2d_bytes_per_step": int(n_utts * n * 2),
               "d2h_bytes_per_step": int(frames * WIDTH * 4), "ms_per_step": float(et.item()) * 1e3, "steps": e2e_steps,
               "api": "afe_batch_run_host (pinned host buffers)", "host_affinity": affinity}
        # the host result must be the device result
        chk = h_out[:998 * 4].to(dev)
        step(); torch.cuda.synchronize()
        e2e["matches_device_path"] = bool(torch.equal(chk, out[:998 * 4]))
        be.close()
        del h_pcm, h_out

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k1 = float(np.mean(k1_ms))
        alg_bytes = frames * (2 * S + 4 * WIDTH)
        achieved = alg_bytes / (k1 * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": b.kernel_name + "<512>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": "measured" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k1, "kernel_share_of_step": k1 / ms_per_step,
                "note": "fp32-issue bound (FFT butterflies), not HBM bound; see DESIGN.md"}
        # what actually bounds the kernel: scheduler issue slots. 441 warp-instructions per frame of which
        # 150 are packed FP32 that hold the issue port for two cycles (profiles/r01_final_k_fused_summary.txt,
        # tools/ubench/issue.cu) = 591 slot-cycles per frame, spread over 4 schedulers per SM at the sampled SM clock.
        try:
            props = torch.cuda.get_device_properties(local)
            slots, mhz = 441.0 + 150.0, float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
            roof["issue_model"] = {"slot_cycles_per_frame": slots, "schedulers": 4 * props.multi_processor_count, "sm_mhz": mhz,
                                   "frac": slots * frames / (4 * props.multi_processor_count * mhz * 1e6 * k1 * 1e-3),
                                   "source": "instruction counts from profiles/r01_final_k_fused_summary.txt (ncu), time and clock live"}
        except Exception:
            pass
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                per_frame = json.load(open(prof)).get("k_fused_mfcc_dram_bytes_per_frame")
                roof["traffic"] = per_frame * frames if per_frame else None   # ncu dram read+write per frame x frames/launch
                roof["traffic_source"] = "profiles/traffic.json (ncu --set full on a 2000-utterance launch, scaled per frame)"
            except Exception:
                pass
        cpu = None
        if not args.no_cpu:
            if _FULL_AFFINITY:
                os.sched_setaffinity(0, _FULL_AFFINITY)   # the CPU baseline uses every host core, not one NUMA node
            threads = os.cpu_count() or 1
            sample = pcm[:64 * n].cpu().numpy().reshape(64, n)
            import oracle_lib as ol
            lib1 = ol.RefLib("ref" if ol.available("ref") else "port")
            _, s1 = lib1.extract(p, [sample[i] for i in range(16)], sample_limit=0, n_threads=1)
            rate1 = 16 * T / s1
            n_cpu = int(max(threads * 2, min(4096, 2.0 * rate1 * threads / T)))
            n_cpu = min(n_cpu, n_utts)
            cs = pcm[:n_cpu * n].cpu().numpy().reshape(n_cpu, n)
            r = cpu_arm(n_cpu, 2, 1, threads, pcm=cs)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                   "value_1core": rate1}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(n_utts, world),
                "audio_hours_per_s": value * S / SR / 3600.0, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "clocks": clocks, "corpus_cmvn": corpus,
                "kernels_per_step": [b.kernel_name],
                "tiles_per_gpu": b.num_tiles, "flags": {"tma": not args.no_tma, "devtools": bool(afe.lib().afe_build_flags() & 1)}}
        print(json.dumps(line), flush=True)
    b.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utts", type=int, default=10000, help="utterances per GPU (BASELINE config 3: 10000)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin ranks to their GPU's NUMA node (A/B, N > 1)")
    ap.add_argument("--no-cluster", action="store_true", help="ticket-scheme normalisation instead of clusters + DSMEM (A/B)")
    ap.add_argument("--no-tma", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""What can this box move between host and GPUs?

One process per GPU (python tools/h2d_d2h_ceiling.py for 1 GPU; torchrun --nproc-per-node N ... for N): every rank pins
the config-3 step's buffers (3.2 GB of int16 PCM in, 1.56 GB of float32 features out), binds itself to its GPU's NUMA node
first, and then moves them with plain cudaMemcpyAsync on two streams (H2D and D2H concurrently, all ranks at once, no kernels,
never a batched-memcpy API), in one piece and in 32 chunks like afe_batch_run_host. Prints one JSON line: the aggregate GB/s
and the frames/s no end-to-end extractor on this box can exceed. bench.py measures the same ceiling in-run
(e2e.copy_ceiling) and reports e2e.frac_of_copy_ceiling.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    aff = bench.bind_to_gpu_numa_node(torch, local) if world > 1 else {"bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_utts, n, T = 10000, 160000, 998
    h_in = torch.empty(n_utts * n, dtype=torch.int16).pin_memory(); h_in.zero_()
    h_out = torch.empty(n_utts * T * 39, dtype=torch.float32).pin_memory(); h_out.zero_()
    d_in, d_out = torch.empty_like(h_in, device=dev), torch.empty_like(h_out, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(chunks, h2d=True, d2h=True):
        best = None
        for _ in range(4):
            barrier()
            t0 = time.perf_counter()
            for c in range(chunks):
                a0, a1 = len(h_in) * c // chunks, len(h_in) * (c + 1) // chunks
                b0, b1 = len(h_out) * c // chunks, len(h_out) * (c + 1) // chunks
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in[a0:a1].copy_(h_in[a0:a1], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out[b0:b1].copy_(d_out[b0:b1], non_blocking=True)
            s1.synchronize(); s2.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best
    gb_in, gb_out = h_in.numel() * 2 / 1e9, h_out.numel() * 4 / 1e9
    res = {"n_gpus": world, "host_affinity": aff, "bytes_per_rank": {"h2d": gb_in * 1e9, "d2h": gb_out * 1e9}}
    for name, kw in (("h2d_only", dict(d2h=False)), ("d2h_only", dict(h2d=False)), ("both_1_chunk", {}), ("both_32_chunks", {})):
        t = run(32 if name.endswith("32_chunks") else 1, **kw)
        moved = (gb_in if kw.get("h2d", True) else 0) + (gb_out if kw.get("d2h", True) else 0)
        res[name] = {"ms": t * 1e3, "aggregate_gb_s": world * moved / t}
    res["frames_per_s_ceiling"] = world * n_utts * T / (res["both_32_chunks"]["ms"] * 1e-3)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

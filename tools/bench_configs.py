#!/usr/bin/env python
"""Device-resident throughput of the fused path on the other BASELINE configurations (parity is covered by tests/):
  c2-like : 16 kHz, 23 mel, 13 MFCC+d+dd, CMN, ragged 3-25 s utterances
  c5      : 8 kHz telephony, 256-pt FFT, 20 mel, ONE 1-hour stream (28.8 M samples), fused deltas, no norm / CMN
Prints one JSON line per case. Usage: python tools/bench_configs.py"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import afe_loader
afe = afe_loader.load()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
g = torch.Generator(device=dev); g.manual_seed(7)

def synth(n, sr):
    t = torch.arange(n, device=dev, dtype=torch.float32) / sr
    f = float(np.random.default_rng(n).uniform(100, 3800))
    x = 3000.0 * torch.randn(n, device=dev, generator=g) + 8000.0 * torch.sin(2 * np.pi * f * t)
    return x.round_().clamp_(-32767, 32767).to(torch.int16)

def run(name, params, lens, sr, steps=10):
    offs, pos = [], 0
    for n in lens:
        offs.append(pos); pos += (n + 7) // 8 * 8
    pcm = torch.zeros(pos + 64, dtype=torch.int16, device=dev)
    for o, n in zip(offs, lens):
        pcm[o:o + n] = synth(n, sr)
    b = afe.BatchMfcc(params, 0, flags=afe.BATCH_Q1_EXACT | (afe.BATCH_WS_KERNEL if os.environ.get("AFE_WS") else 0))
    b.set_stream(stream.cuda_stream)
    frames = b.plan(np.array(offs, np.int64), np.array(lens, np.int64))
    out = torch.empty((frames, b.width), dtype=torch.float32, device=dev)
    for _ in range(3):
        b.run_device(pcm.data_ptr(), out.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(steps):
        b.run_device(pcm.data_ptr(), out.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    S = params.shift
    bytes_alg = frames * (2 * S + 4 * b.width)
    print(json.dumps({"case": name, "frames": frames, "utterances": len(lens), "tiles": b.num_tiles, "ms": ms,
                      "frames_per_s": frames / ms * 1e3, "audio_hours_per_s": frames * S / sr / 3600 / ms * 1e3,
                      "achieved_GBps": bytes_alg / ms / 1e6, "kernel": b.kernel_name, "finite": bool(torch.isfinite(out).all())}), flush=True)
    b.close()

rng = np.random.default_rng(1)
run("c2-like 16k/23mel ragged 3-25s x4000", afe.make_params(num_banks=23, norm=1, dyn=2), [int(x) for x in rng.integers(48000, 400000, 4000)], 16000.0)
run("c3 16k/40mel 10s x10000", afe.make_params(num_banks=40, norm=1, dyn=2), [160000] * 10000, 16000.0)
tel = dict(window_size=200, shift=80, num_banks=20, sample_rate=8000.0, high_freq=4000.0, dyn=2)
run("c5 8k/20mel one 1-hour stream, no norm", afe.make_params(norm=0, **tel), [28800000], 8000.0, steps=20)
run("c5 8k/20mel one 1-hour stream, CMN", afe.make_params(norm=1, **tel), [28800000], 8000.0, steps=20)
run("8k/20mel 10s x10000", afe.make_params(norm=1, **tel), [80000] * 10000, 8000.0)

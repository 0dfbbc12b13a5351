#!/usr/bin/env python
"""Per-file cost of the drop-in streaming object (MfccCuda: set_input -> apply -> get_output_data -> flush -> apply ->
get_output_data, one object reused with reset()), the way the reference driver uses ParamBase*. Launch/latency bound.
Usage: python tools/bench_stream.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import afe_loader
afe = afe_loader.load()
from common import synth_utterances

def run(seconds, n_files=40, **kw):
    pcm = synth_utterances(1, 16000 * seconds, seed=3)[0]
    p = afe.make_params(input_buffer_size=1 << 22, window_size=400, shift=160, num_banks=40, sample_rate=16000.0, low_freq=64.0,
                        high_freq=8000.0, ceps_len=12, want_c0=1, lift_coef=22.0, norm=1, dyn=2, delta_l1=3, delta_l2=3,
                        norm_after_dyn=1)
    m = afe.MfccCuda(p, 0)
    m.set_window(afe.make_window(400))
    def one():
        m.reset()
        wc = m.set_input(pcm); m.apply(); a = m.get_output_data(wc)
        wc2 = m.flush(); m.apply(); b = m.get_output_data(wc2)
        return wc + wc2
    for _ in range(3): frames = one()
    t0 = time.perf_counter()
    for _ in range(n_files): one()
    dt = (time.perf_counter() - t0) / n_files
    m.close()
    print(json.dumps({"case": f"MfccCuda streaming object, one {seconds} s file per iteration", "frames": frames,
                      "ms_per_file": dt * 1e3, "frames_per_s": frames / dt}), flush=True)

for s in (3, 10, 60):
    run(s)

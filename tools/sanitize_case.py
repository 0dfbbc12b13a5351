"""Small cases for compute-sanitizer (memcheck / racecheck): fused batch (512 and 256 point, CMN fused + unfused),
streaming object. Run as: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import afe_loader
afe = afe_loader.load()
rng = np.random.default_rng(0)
utts = [np.clip(np.round(3000 * rng.standard_normal(n)), -32767, 32767).astype(np.int16) for n in (9000, 30001, 5000, 88000)]
for kw, flags in ((dict(num_banks=40, norm=1, dyn=2), afe.BATCH_Q1_EXACT),
                  (dict(num_banks=23, norm=2, dyn=2), afe.BATCH_UNFUSED_NORM),
                  (dict(window_size=200, shift=80, num_banks=20, sample_rate=8000.0, high_freq=4000.0, norm=3, dyn=1), afe.BATCH_NO_TMA),
                  (dict(num_banks=40, ceps_len=0, norm=0, dyn=0), 0)):
    p = afe.make_params(input_buffer_size=1 << 22, **kw)
    b = afe.BatchMfcc(p, 0, flags=flags)
    pcm, offs, lens = afe.pack_utterances(utts)
    b.plan(offs, lens)
    out = b.run_host(np.concatenate([pcm, np.zeros(16, np.int16)]))
    assert np.isfinite(out).all()
    print("batch", kw, out.shape, float(np.abs(out).max()))
    b.close()
p = afe.make_params(input_buffer_size=16000, num_banks=23, norm=1, dyn=2)
print("stream", afe.extract_stream(p, utts[3]).shape)

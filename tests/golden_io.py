"""Loader for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def load_pcm():
    if "pcm" not in _cache:
        z = np.load(os.path.join(HERE, "pcm_v1.npz"))
        pcm = {k: z[k] for k in ("sample1", "a0001", "a1")}
        a1 = pcm["a1"]
        for n in ("a2", "a3", "a4", "a5"):
            res = z[n + "_residual"].astype(np.int32)
            reps = -(-len(res) // len(a1))
            pcm[n] = (np.tile(a1, reps)[:len(res)].astype(np.int32) + res).astype(np.int16)
        pcm["a1_8k"] = a1[::2].copy()
        _cache["pcm"] = pcm
    return _cache["pcm"]


def load_golden():
    """-> {case: dict(utt, params, sample_limit, frames, width, rows, feats)}"""
    if "gold" not in _cache:
        z = np.load(os.path.join(HERE, "golden_v1.npz"))
        meta = json.loads(bytes(z["meta"]).decode())
        for name, m in meta.items():
            m["rows"] = z[name + "/rows"]
            m["feats"] = z[name + "/feats"]
        _cache["gold"] = meta
    return _cache["gold"]

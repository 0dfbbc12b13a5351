"""Generate tests/golden/*.npz from the reference's own CPU path (oracle/_ref/libref_mfcc.so).

Run HERE (the authoring container), where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py
The GPU box has no /root/reference, so the inputs (the reference's 16 kHz fixtures) and
the expected outputs are committed:
  pcm_v1.npz     int16 PCM: sample1, a0001, a1 stored verbatim; a2..a5 stored as the (<=13 LSB) residual against
                 a1 tiled, which is how those files were made -> reconstructed by golden.load_pcm().
  golden_v1.npz  reference outputs for the cases listed in CASES (rows subsampled for the bigger cases, the row
                 index is stored), plus the parameter set of each case as JSON.
Nothing here is executed by the product.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import RefLib, default_params, read_pcm, total_frames  # noqa: E402

REF = "/root/reference"
BIG = 1 << 22  # "sample_limit" larger than any fixture: single set_input + flush (reference default is 10 M; Q5 makes that slow)


def tile_like(a1, n):
    reps = -(-n // len(a1))
    return np.tile(a1, reps)[:n]


def rows_subset(T, full):
    if full:
        return np.arange(T)
    idx = set(range(0, min(T, 24))) | set(range(max(0, T - 24), T)) | set(range(0, T, 7))
    return np.array(sorted(idx))


def main():
    L = RefLib("ref")
    pcm = {"sample1": read_pcm(f"{REF}/sample1.wav")}
    for n in ("a0001", "a1", "a2", "a3", "a4", "a5"):
        pcm[n] = read_pcm(f"{REF}/soundfiles/{n}.wav")
    store = {"sample1": pcm["sample1"], "a0001": pcm["a0001"], "a1": pcm["a1"]}
    for n in ("a2", "a3", "a4", "a5"):
        res = pcm[n].astype(np.int32) - tile_like(pcm["a1"], len(pcm[n])).astype(np.int32)
        assert np.abs(res).max() < 127
        store[n + "_residual"] = res.astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "pcm_v1.npz"), **store)

    # 8 kHz telephony variant (BASELINE config 5 parameters) on decimated a1; aliasing is irrelevant for parity.
    pcm["a1_8k"] = pcm["a1"][::2].copy()

    P = default_params
    c2 = dict(norm="cmn", dyn="acc", delta_l1=3, delta_l2=3)
    tel = dict(window_size=200, shift=80, num_banks=20, sample_rate=8000.0, high_freq=4000.0)
    # name -> (utterance, params, sample_limit, full rows?)
    CASES = {
        "c1_sample1": ("sample1", P(), BIG, True),
        "c1_sample1_13nc0": ("sample1", P(ceps_len=13, want_c0=0), BIG, False),
        "c1_sample1_acc": ("sample1", P(dyn="acc"), BIG, False),
        "c2_a1_single": ("a1", P(**c2), BIG, True),
        "c2_a1_blocks": ("a1", P(**c2), 0, True),
        "c2_a1_stream16k": ("a1", P(**c2), 16000, False),
        "c2_a0001_single": ("a0001", P(**c2), BIG, False),
        "c2_a2_single": ("a2", P(**c2), BIG, False),
        "c2_a3_single": ("a3", P(**c2), BIG, False),
        "c2_a4_single": ("a4", P(**c2), BIG, False),
        "c2_a5_single": ("a5", P(**c2), BIG, False),
        "c2_a5_blocks": ("a5", P(**c2), 0, False),
        "c3_a1_40mel": ("a1", P(num_banks=40, **c2), BIG, False),
        "c3_a1_40mel_nonorm": ("a1", P(num_banks=40, dyn="acc"), BIG, False),
        "v_a1_cvn": ("a1", P(norm="cvn", dyn="acc"), BIG, False),
        "v_a1_minmax": ("a1", P(norm="minmax", dyn="acc"), BIG, False),
        "v_a1_delta_only": ("a1", P(norm="cmn", dyn="delta"), BIG, False),
        "v_a1_l1_2_l2_1": ("a1", P(norm="cmn", dyn="acc", delta_l1=2, delta_l2=1), BIG, False),
        "v_a1_norm_before_dyn": ("a1", P(norm="cvn", dyn="acc", norm_after_dyn=0), BIG, False),
        "v_a1_fbank": ("a1", P(ceps_len=0, norm="cmn", dyn="acc"), BIG, False),
        "v_a1_nolifter_noc0": ("a1", P(want_c0=0, lift_coef=1e9), BIG, False),
        "v_a1_alpha090": ("a1", P(alpha=0.9, **c2), BIG, False),
        "v_a1_alpha112": ("a1", P(alpha=1.12, **c2), BIG, False),
        "c5_a1_8k": ("a1_8k", P(dyn="acc", **tel), BIG, False),
        "c5_a1_8k_cmn_stream": ("a1_8k", P(norm="cmn", dyn="acc", **tel), 8000, False),
    }
    out = {}
    meta = {}
    for name, (utt, p, limit, full) in CASES.items():
        res, _ = L.extract(p, [pcm[utt]], sample_limit=limit)
        feats = res[0]
        T = total_frames(len(pcm[utt]), p)
        assert feats.shape[0] == T, (name, feats.shape, T)
        rows = rows_subset(T, full)
        out[name + "/rows"] = rows.astype(np.int32)
        out[name + "/feats"] = feats[rows]
        meta[name] = dict(utt=utt, params=p, sample_limit=limit, frames=int(T), width=int(feats.shape[1]))
        print(f"{name:28s} T={T:5d} width={feats.shape[1]:3d} row0={feats[0, :3]}")
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    for f in ("pcm_v1.npz", "golden_v1.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()

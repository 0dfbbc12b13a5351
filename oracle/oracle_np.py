"""TEST INFRASTRUCTURE — float64 NumPy restatement of the reference arithmetic, batch form.

Independent of the C oracles (different precision, different structure: whole-utterance, no streaming state),
used to cross-check them at small sizes. Steps cite /root/reference files:
  window            ASR_OCL.cpp:149-152
  framing           segmentercpu.cpp:17-28, frame count parambase.cpp:16-19
  |X|/N2            mfcccpu.cpp:203
  mel bank + log    mfcccpu.cpp:24-60, 192-220
  DCT-II + lifter   mfcccpu.cpp:118-136, 222-232
  deltas            mfcccpu.cpp:234-263, deltacpu.cpp:16-30 (computed on the edge-replicated, EXTENDED axis)
  normalisation     normalizercpu.cpp:22-89, stats over the first T-D rows (one-block utterance)
"""
import numpy as np


def make_window(W):
    i = np.arange(W)
    return ((0.56 - 0.46 * np.cos(2.0 * np.pi * i / W)) / 32768.0)


def mel_tables(nb, N2, sr, lo, hi, alpha=1.0):
    mel = lambda f: 1127.0 * np.log(f / 700.0 + 1.0)
    imel = lambda m: 700.0 * (np.exp(m / 1127.0) - 1.0)
    i = np.arange(nb + 2)
    f = imel(i / (nb + 1.0) * (mel(hi) - mel(lo)) + mel(lo))
    o = 2 * np.pi * f / sr
    o = o + 2 * np.arctan(((1 - alpha) * np.sin(o)) / (1 - (1 - alpha) * np.cos(o)))
    cent = sr * o / (2 * np.pi)
    edge = np.floor(cent * N2 / sr + 0.5).astype(int)
    Wm = np.zeros((nb, N2 // 2 + 1))
    for b in range(nb):
        cl, cc, cr = cent[b], cent[b + 1], cent[b + 2]
        for j in range(edge[b], min(edge[b + 2], N2 // 2 + 1)):
            fj = j * sr / N2
            Wm[b, j] = max(0.0, min((fj - cl) / (cc - cl), (fj - cr) / (cc - cr)))
    return cent, edge, Wm


def dct_matrix(nb, C, want_c0, lift):
    k = np.arange(nb)[:, None]
    i = np.arange(1, C + 1)[None, :]
    M = (1 + lift / 2 * np.sin(np.pi * i / lift)) * np.sqrt(2.0 / nb) * np.cos(np.pi * i * (k + 0.5) / nb)
    if want_c0:
        M = np.concatenate([M, np.full((nb, 1), np.sqrt(2.0 / nb))], axis=1)
    return M


def delta_ext(x, L):
    """x: [R+2L, d] -> [R, d]"""
    R = x.shape[0] - 2 * L
    den = 2.0 * sum(l * l for l in range(1, L + 1))
    out = np.zeros((R, x.shape[1]))
    for l in range(1, L + 1):
        out += l * (x[L + l:L + l + R] - x[L - l:L - l + R])
    return out / den


def mfcc(pcm, p, q1=False, stats_rows=None):
    """Whole-utterance features [T, width] in float64. q1=True reproduces the single-block flush bug."""
    W, S, nb = p["window_size"], p["shift"], p["num_banks"]
    N2 = 1 << (W - 1).bit_length()
    T = (len(pcm) - (W - S)) // S
    w = make_window(W).astype(np.float32).astype(np.float64)
    idx = np.arange(T)[:, None] * S + np.arange(W)[None, :]
    fr = np.zeros((T, N2))
    x = pcm.astype(np.float64)[idx]
    pre = p.get("preemphasis", 0.0)   # NOT in the reference (segmentercpu.cpp:21-27 has none): the new path's optional
    if pre:                           # per-frame pre-emphasis y[j] = x[j] - pre*x[j-1], y[0] = (1-pre)*x[0]
        x = x - pre * np.concatenate([x[:, :1], x[:, :-1]], axis=1)
    fr[:, :W] = x * w
    mag = np.abs(np.fft.rfft(fr, axis=1)) / N2
    _, edge, Wm = mel_tables(nb, N2, p["sample_rate"], p["low_freq"], p["high_freq"], p.get("alpha", 1.0))
    E = np.log(np.maximum(mag @ Wm.T, 1e-30))
    C = p["ceps_len"]
    c = E @ dct_matrix(nb, C, p["want_c0"], p["lift_coef"]) if C > 0 else E
    dyn = p["dyn"]
    l1 = p["delta_l1"] if dyn else 0
    l2 = p["delta_l2"] if dyn == 2 else 0
    D = l1 + l2
    n_stats = (T - D) if stats_rows is None else stats_rows

    def norm(x, ref_rows):
        kind = p["norm"]
        if kind == 0:
            return x
        r = x[:ref_rows]
        mu = r.mean(0)
        if kind == 1:
            return x - mu
        if kind == 2:
            n = ref_rows
            return (x - mu) * np.sqrt((n - 1) / ((r * r).sum(0) - r.sum(0) ** 2 / n))
        return (x - mu) / np.maximum(np.abs(r.min(0) - mu), np.abs(r.max(0) - mu))

    if not p["norm_after_dyn"]:
        c = norm(c, T)
    streams = [c]
    if dyn:
        pad = np.concatenate([np.repeat(c[:1], D, 0), c, np.repeat(c[-1:], D, 0)])
        d1 = delta_ext(pad, l1)            # T + 2*l2 rows on the extended axis
        streams.append(d1[l2:l2 + T])
        if dyn == 2:
            streams.append(delta_ext(d1, l2))
    if q1 and D > 0:
        s = streams[0].copy()
        s[T - D:] = streams[0][T - 2 * D:T - D]
        streams[0] = s
    if p["norm_after_dyn"]:
        if q1 and D > 0:
            # stats are taken before the flush rows exist, i.e. on the un-shifted statics
            base = [c] + streams[1:]
            streams = [norm(np.concatenate([b[:n_stats], s[n_stats:]]), n_stats) for b, s in zip(base, streams)]
        else:
            streams = [norm(s, n_stats) for s in streams]
    return np.concatenate(streams, axis=1)

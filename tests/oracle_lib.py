"""ctypes access to the CPU oracles (TEST INFRASTRUCTURE ONLY).

`RefLib("ref")`   -> oracle/_ref/libref_mfcc.so   : the reference's own CPU classes (oracle/ref_capi.cpp)
`RefLib("port")`  -> oracle/liboracle_port.so     : our C restatement (oracle/mfcc_port.c)
Both export the same `<prefix>_*` C API, so every test can run against either.
Nothing under asr-featext-opencl_b200/ may import this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NORM = {"none": 0, "cmn": 1, "cvn": 2, "minmax": 3}
DYN = {"none": 0, "delta": 1, "acc": 2}

_f32p = C.POINTER(C.c_float)
_i16p = C.POINTER(C.c_short)
_i64p = C.POINTER(C.c_longlong)


class RefParams(C.Structure):
    # mirrors struct ref_params in oracle/ref_capi.cpp and oracle/mfcc_port.c
    _fields_ = [
        ("window_size", C.c_int), ("shift", C.c_int), ("num_banks", C.c_int),
        ("sample_rate", C.c_float), ("low_freq", C.c_float), ("high_freq", C.c_float),
        ("ceps_len", C.c_int), ("want_c0", C.c_int), ("lift_coef", C.c_float),
        ("norm", C.c_int), ("dyn", C.c_int), ("delta_l1", C.c_int), ("delta_l2", C.c_int),
        ("norm_after_dyn", C.c_int), ("alpha", C.c_float),
    ]


def default_params(**kw):
    """Reference defaults for the 16 kHz configs (ASR_OCL.cpp:560 + BASELINE.json configs)."""
    p = dict(window_size=400, shift=160, num_banks=23, sample_rate=16000.0, low_freq=64.0,
             high_freq=8000.0, ceps_len=12, want_c0=1, lift_coef=22.0, norm=0, dyn=0,
             delta_l1=3, delta_l2=3, norm_after_dyn=1, alpha=1.0)
    for k, v in kw.items():
        if k not in p:
            raise KeyError(k)
        p[k] = v
    if isinstance(p["norm"], str):
        p["norm"] = NORM[p["norm"]]
    if isinstance(p["dyn"], str):
        p["dyn"] = DYN[p["dyn"]]
    return p


def cols_of(p):
    return p["ceps_len"] + (1 if p["want_c0"] else 0) if p["ceps_len"] > 0 else p["num_banks"]


def width_of(p):
    return cols_of(p) * {0: 1, 1: 2, 2: 3}[p["dyn"]]


def total_frames(n, p):
    return max(0, (n - (p["window_size"] - p["shift"])) // p["shift"])


_PATHS = {
    "ref": os.path.join(ROOT, "oracle", "_ref", "libref_mfcc.so"),
    "ref_f64": os.path.join(ROOT, "oracle", "_ref", "libref_mfcc_f64.so"),
    "ref_dlibm": os.path.join(ROOT, "oracle", "_ref", "libref_mfcc_dlibm.so"),
    "port": os.path.join(ROOT, "oracle", "liboracle_port.so"),
}


def available(kind):
    return os.path.exists(_PATHS[kind])


class RefLib:
    def __init__(self, kind="ref"):
        self.kind = kind
        self.prefix = "port" if kind == "port" else "ref"
        self.lib = C.CDLL(_PATHS[kind])
        L, px = self.lib, self.prefix

        def fn(name, res, args):
            f = getattr(L, f"{px}_{name}")
            f.restype, f.argtypes = res, args
            return f

        self.last_error = fn("last_error", C.c_char_p, [])
        self.mfcc_create = fn("mfcc_create", C.c_void_p,
                              [C.c_int] * 4 + [C.c_float] * 3 + [C.c_int, C.c_int, C.c_float] + [C.c_int] * 5)
        self.mfcc_destroy = fn("mfcc_destroy", None, [C.c_void_p])
        self.mfcc_set_window = fn("mfcc_set_window", None, [C.c_void_p, _f32p])
        self.mfcc_set_alpha = fn("mfcc_set_alpha", None, [C.c_void_p, C.c_float])
        self.mfcc_input_buffer_size = fn("mfcc_input_buffer_size", C.c_int, [C.c_void_p])
        self.mfcc_estimated_window_count = fn("mfcc_estimated_window_count", C.c_int, [C.c_void_p, C.c_int])
        self.mfcc_width = fn("mfcc_width", C.c_int, [C.c_void_p])
        self.mfcc_set_input = fn("mfcc_set_input", C.c_int, [C.c_void_p, _i16p, C.c_int])
        self.mfcc_flush = fn("mfcc_flush", C.c_int, [C.c_void_p])
        self.mfcc_apply = fn("mfcc_apply", C.c_int, [C.c_void_p])
        self.mfcc_get_output = fn("mfcc_get_output", C.c_int, [C.c_void_p, _f32p, C.c_int])
        self.segmenter_create = fn("segmenter_create", C.c_void_p, [C.c_int] * 4)
        self.segmenter_destroy = fn("segmenter_destroy", None, [C.c_void_p])
        self.segmenter_set_window = fn("segmenter_set_window", None, [C.c_void_p, _f32p])
        self.segmenter_set_input = fn("segmenter_set_input", C.c_int,
                                      [C.c_void_p, _i16p, _f32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)])
        self.segmenter_flush = fn("segmenter_flush", C.c_int,
                                  [C.c_void_p, _f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)])
        self.segmenter_remaining_samples = fn("segmenter_remaining_samples", C.c_int, [C.c_void_p])
        self.segmenter_samples = fn("segmenter_samples", C.c_int, [C.c_void_p])
        self.segmenter_is_flushed = fn("segmenter_is_flushed", C.c_int, [C.c_void_p])
        self.segmenter_was_flushed = fn("segmenter_was_flushed", C.c_int, [C.c_void_p])
        self.delta_apply = fn("delta_apply", None, [_f32p, _f32p, C.c_int, C.c_int, C.c_int])
        self.normalizer_create = fn("normalizer_create", C.c_void_p, [C.c_int, C.c_int])
        self.normalizer_destroy = fn("normalizer_destroy", None, [C.c_void_p])
        self.normalizer_normalize = fn("normalizer_normalize", None, [C.c_void_p, _f32p, C.c_int, C.c_int])
        self.make_window = fn("make_window", None, [_f32p, C.c_int])
        self.extract_c = fn("extract", C.c_int,
                            [C.POINTER(RefParams), _i16p, _i64p, C.c_int, _f32p, _i64p, C.c_int, C.c_int,
                             _i64p, C.POINTER(C.c_double)])

    # ------------------------------------------------------------------ helpers
    def window(self, n):
        w = np.empty(n, np.float32)
        self.make_window(w.ctypes.data_as(_f32p), n)
        return w

    def err(self):
        return (self.last_error() or b"").decode()

    def extract(self, p, utts, sample_limit=0, n_threads=1):
        """Run the reference driver loop over a list of int16 arrays.

        sample_limit=0: MfccCpu sized to each utterance (reference blocks it into <=2 set_input calls);
        sample_limit>=len: single set_input + flush (Q1 path), as the reference driver's default 10 M.
        Returns (list of [T,width] float32 arrays, seconds).
        """
        utts = [np.ascontiguousarray(u, np.int16) for u in utts]
        offs = np.zeros(len(utts) + 1, np.int64)
        offs[1:] = np.cumsum([len(u) for u in utts])
        pcm = np.concatenate(utts) if utts else np.zeros(0, np.int16)
        foffs = np.zeros(len(utts) + 1, np.int64)
        foffs[1:] = np.cumsum([total_frames(len(u), p) for u in utts])
        w = width_of(p)
        out = np.zeros((int(foffs[-1]), w), np.float32)
        got = np.zeros(len(utts), np.int64)
        secs = C.c_double(0)
        rp = RefParams(**p)
        rc = self.extract_c(C.byref(rp), pcm.ctypes.data_as(_i16p), offs.ctypes.data_as(_i64p), len(utts),
                            out.ctypes.data_as(_f32p), foffs.ctypes.data_as(_i64p), int(sample_limit),
                            int(n_threads), got.ctypes.data_as(_i64p), C.byref(secs))
        if rc != 0:
            raise RuntimeError(self.err())
        res = [out[foffs[i]:foffs[i] + got[i]] for i in range(len(utts))]
        return res, secs.value


class RefMfcc:
    """Thin object wrapper with the ParamBase verb set (parambase.h:23-32)."""

    def __init__(self, lib, input_buffer_size, p):
        self.L = lib
        self.p = p
        self.h = lib.mfcc_create(int(input_buffer_size), p["window_size"], p["shift"], p["num_banks"],
                                 p["sample_rate"], p["low_freq"], p["high_freq"], p["ceps_len"], p["want_c0"],
                                 p["lift_coef"], p["norm"], p["dyn"], p["delta_l1"], p["delta_l2"],
                                 p["norm_after_dyn"])
        if not self.h:
            raise RuntimeError(lib.err())
        self.set_alpha(p.get("alpha", 1.0))

    def close(self):
        if self.h:
            self.L.mfcc_destroy(self.h)
            self.h = None

    __del__ = close

    def set_window(self, w):
        w = np.ascontiguousarray(w, np.float32)
        self.L.mfcc_set_window(self.h, w.ctypes.data_as(_f32p))

    def set_alpha(self, a):
        self.L.mfcc_set_alpha(self.h, float(a))

    def get_input_buffer_size(self):
        return self.L.mfcc_input_buffer_size(self.h)

    def estimated_window_count(self, n):
        return self.L.mfcc_estimated_window_count(self.h, int(n))

    def get_output_data_width(self):
        return self.L.mfcc_width(self.h)

    def set_input(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        r = self.L.mfcc_set_input(self.h, pcm.ctypes.data_as(_i16p), len(pcm))
        if r < 0:
            raise RuntimeError(self.L.err())
        return r

    def flush(self):
        r = self.L.mfcc_flush(self.h)
        if r < 0:
            raise RuntimeError(self.L.err())
        return r

    def apply(self):
        if self.L.mfcc_apply(self.h) != 0:
            raise RuntimeError(self.L.err())

    def get_output_data(self, wc):
        out = np.zeros((wc, self.get_output_data_width()), np.float32)
        if wc and self.L.mfcc_get_output(self.h, out.ctypes.data_as(_f32p), wc) != 0:
            raise RuntimeError(self.L.err())
        return out


def read_pcm(path):
    """16-bit mono PCM from RIFF (44-byte header) or NIST SPHERE (1024-byte header) — SURVEY §2 #15."""
    raw = open(path, "rb").read()
    skip = 1024 if raw[:4] == b"NIST" else 44
    return np.frombuffer(raw[skip:], dtype="<i2").copy()

"""CPU tier: pins the oracle. (1) FFT shim vs the DFT definition; (2) C port == reference's own classes bit-for-bit;
(3) both == committed golden vectors; (4) float64 NumPy restatement within float noise; (5) known-answer stage tests;
(6) streaming invariants of the reference state machine."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from golden_io import load_golden, load_pcm


def test_fft_shim_matches_dft_definition(port):
    lib = port.lib
    lib.fftwf_alloc_real.restype = C.c_void_p
    lib.fftwf_alloc_complex.restype = C.c_void_p
    lib.fftwf_plan_many_dft_r2c.restype = C.c_void_p
    lib.fftwf_execute.argtypes = [C.c_void_p]
    lib.fftwf_destroy_plan.argtypes = [C.c_void_p]
    rng = np.random.default_rng(0)
    for n in (256, 512, 64, 1024):
        hm = 11  # not a multiple of the SIMD group
        x = rng.standard_normal((hm, n)).astype(np.float32)
        out = np.zeros((hm, n, 2), np.float32)
        nn = C.c_int(n)
        plan = lib.fftwf_plan_many_dft_r2c(1, C.byref(nn), hm, C.c_void_p(x.ctypes.data), None, 1, n,
                                           C.c_void_p(out.ctypes.data), None, 1, n, 0)
        assert plan
        lib.fftwf_execute(C.c_void_p(plan))
        lib.fftwf_destroy_plan(C.c_void_p(plan))
        got = out[:, :n // 2 + 1, 0] + 1j * out[:, :n // 2 + 1, 1]
        j = np.arange(n)
        k = np.arange(n // 2 + 1)
        F = np.exp(-2j * np.pi * np.outer(k, j) / n)          # the definition, float64
        want = x.astype(np.float64) @ F.T
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 5e-7, (n, err)


@pytest.mark.parametrize("case", sorted(load_golden().keys()))
def test_port_matches_golden(port, case):
    g = load_golden()[case]
    res, _ = port.extract(g["params"], [load_pcm()[g["utt"]]], sample_limit=g["sample_limit"])
    assert res[0].shape == (g["frames"], g["width"])
    np.testing.assert_array_equal(res[0][g["rows"]], g["feats"])   # bit-exact: same expression order, same libm


@pytest.mark.parametrize("case", ["c1_sample1", "c2_a1_single", "c2_a1_stream16k", "v_a1_cvn", "c5_a1_8k"])
def test_reference_build_matches_golden(case):
    if not ol.available("ref"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    g = load_golden()[case]
    res, _ = ol.RefLib("ref").extract(g["params"], [load_pcm()[g["utt"]]], sample_limit=g["sample_limit"])
    np.testing.assert_array_equal(res[0][g["rows"]], g["feats"])


def test_survey_anchor_values():
    """Appendix C of SURVEY.md: first rows produced by an independent probe build of the reference."""
    G = load_golden()
    np.testing.assert_allclose(G["c1_sample1_13nc0"]["feats"][0, :5],
                               [-7.502535, 0.747759, -2.914673, 0.202171, -0.076944], atol=2e-5)
    assert abs(G["c1_sample1"]["feats"][0, 12] - (-79.277710)) < 1e-4
    assert abs(G["c2_a1_single"]["feats"][0, 12] - (-17.123459)) < 1e-4
    assert abs(G["c3_a1_40mel_nonorm"]["feats"][0, 12] - (-90.482315)) < 1e-4
    for name, frames in (("c2_a0001_single", 711), ("c2_a1_single", 504), ("c2_a2_single", 1011),
                         ("c2_a3_single", 1517), ("c2_a4_single", 2023), ("c2_a5_single", 2529)):
        assert G[name]["frames"] == frames


@pytest.mark.parametrize("case,q1", [("c1_sample1", False), ("c2_a1_single", True), ("c2_a1_blocks", False),
                                      ("c3_a1_40mel", True), ("v_a1_cvn", True), ("v_a1_minmax", True),
                                      ("v_a1_delta_only", True), ("v_a1_l1_2_l2_1", True), ("v_a1_fbank", True),
                                      ("v_a1_alpha090", True), ("c5_a1_8k", True), ("v_a1_norm_before_dyn", True)])
def test_numpy_float64_restatement(case, q1):
    import sys, os
    sys.path.insert(0, os.path.join(ol.ROOT, "oracle"))
    import oracle_np
    g = load_golden()[case]
    want = oracle_np.mfcc(load_pcm()[g["utt"]], g["params"], q1=q1)
    err = np.abs(want[g["rows"]] - g["feats"]).max()
    # float32 pipeline vs float64 restatement: c0 ~ -80 carries ~1e-5 of rounding, CVN divides by small std
    assert err < (2e-3 if g["params"]["norm"] in (2, 3) else 3e-4), err


def test_delta_of_linear_ramp_is_slope(oracle):
    L, rows, dim = 3, 20, 5
    slope = np.arange(1, dim + 1, dtype=np.float32) * 0.25
    x = (np.arange(rows + 2 * L, dtype=np.float32)[:, None] * slope[None, :]).copy()
    out = np.zeros((rows, dim), np.float32)
    fp = C.POINTER(C.c_float)
    oracle.delta_apply(x.ctypes.data_as(fp), out.ctypes.data_as(fp), dim, rows, L)
    np.testing.assert_allclose(out, np.tile(slope, (rows, 1)), rtol=1e-6)


@pytest.mark.parametrize("norm", ["cmn", "cvn", "minmax"])
def test_normalizer_known_answers(oracle, norm):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((200, 13)) * 3 + 7).astype(np.float32)
    y = x.copy()
    h = oracle.normalizer_create(ol.NORM[norm], 13)
    oracle.normalizer_normalize(h, y.ctypes.data_as(C.POINTER(C.c_float)), 200, 0)
    assert np.abs(y.mean(0)).max() < 1e-5
    if norm == "cvn":
        np.testing.assert_allclose(y.std(0, ddof=1), 1.0, rtol=1e-5)
    if norm == "minmax":
        np.testing.assert_allclose(np.abs(y).max(0), 1.0, rtol=1e-5)
    # use_last_stats re-applies the SAME affine map to new rows
    z = x[:7].copy()
    oracle.normalizer_normalize(h, z.ctypes.data_as(C.POINTER(C.c_float)), 7, 1)
    np.testing.assert_array_equal(z, y[:7])
    oracle.normalizer_destroy(h)


def test_streamed_equals_blocks_without_norm(oracle):
    """any blocking of the same utterance yields identical rows when normalisation is off."""
    pcm = load_pcm()["a1"]
    p = ol.default_params(dyn="acc")
    a, _ = oracle.extract(p, [pcm], sample_limit=0)
    b, _ = oracle.extract(p, [pcm], sample_limit=16000)
    c, _ = oracle.extract(p, [pcm], sample_limit=5000)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[0], c[0])


def test_q1_single_block_flush_rows(oracle):
    """Q1: with ONE set_input, the flushed rows' statics repeat rows T-2D..T-D-1; deltas are unaffected."""
    pcm = load_pcm()["a1"]
    p = ol.default_params(dyn="acc")
    single, _ = oracle.extract(p, [pcm], sample_limit=1 << 20)
    blocks, _ = oracle.extract(p, [pcm], sample_limit=0)
    T, D, c = 504, 6, 13
    np.testing.assert_array_equal(single[0][:T - D], blocks[0][:T - D])
    np.testing.assert_array_equal(single[0][T - D:, c:], blocks[0][T - D:, c:])
    np.testing.assert_array_equal(single[0][T - D:, :c], blocks[0][T - 2 * D:T - D, :c])


def test_frame_count_rule(oracle):
    p = ol.default_params()
    m = ol.RefMfcc(oracle, 160000, p)
    for n, t in ((54682, 340), (160000, 998), (81000, 504), (400, 1), (399, 0)):
        assert m.estimated_window_count(n) == t
    assert m.get_input_buffer_size() == 998 * 160 + 240
    with pytest.raises(RuntimeError, match="buffer is too small"):
        m.set_input(np.zeros(160000, np.int16))
    m.close()

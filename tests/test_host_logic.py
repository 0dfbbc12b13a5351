"""CPU tier: host-side logic of the product (tables, frame-count rule, sharding, statistics finalize) and the
NumPy model of the in-kernel FFT decomposition. No GPU needed."""
import os
import sys

import numpy as np
import pytest

import afe_loader
import oracle_lib as ol

afe = afe_loader.load()
sys.path.insert(0, os.path.join(ol.ROOT, "oracle"))
sys.path.insert(0, os.path.join(ol.ROOT, "tools"))
import oracle_np  # noqa: E402
import fft_model  # noqa: E402


@pytest.mark.parametrize("nb,W,sr,hi,alpha", [(23, 400, 16000.0, 8000.0, 1.0), (40, 400, 16000.0, 8000.0, 1.0),
                                              (20, 200, 8000.0, 4000.0, 1.0), (40, 400, 16000.0, 8000.0, 0.9),
                                              (23, 400, 16000.0, 8000.0, 1.12), (15, 400, 16000.0, 8000.0, 1.0)])
def test_filter_tables_match_float64_restatement(nb, W, sr, hi, alpha):
    p = afe.make_params(window_size=W, num_banks=nb, sample_rate=sr, high_freq=hi)
    edges, filt = afe.build_filters(p, alpha)
    N2 = 1 << (W - 1).bit_length()
    _, e64, W64 = oracle_np.mel_tables(nb, N2, sr, 64.0, hi, alpha)
    np.testing.assert_array_equal(edges, e64)
    assert np.all(np.diff(edges) >= 0) and edges[-1] <= N2 // 2
    for b in range(nb):
        np.testing.assert_allclose(filt[b % 2, edges[b]:edges[b + 2]], W64[b, edges[b]:edges[b + 2]], atol=3e-5)


def test_survey_filter_edges():
    """cfgA/23 mel edges 2,5,8,...,205,229,256; 40 mel 2,4,5,7,...,240,256; cfgB/20 mel 2,4,7,...,116,128."""
    e23, _ = afe.build_filters(afe.make_params(num_banks=23))
    assert list(e23[:3]) == [2, 5, 8] and list(e23[-3:]) == [205, 229, 256]
    e40, _ = afe.build_filters(afe.make_params(num_banks=40))
    assert list(e40[:4]) == [2, 4, 5, 7] and list(e40[-2:]) == [240, 256] and len(set(e40)) == 42
    e20, _ = afe.build_filters(afe.make_params(window_size=200, shift=80, num_banks=20, sample_rate=8000.0, high_freq=4000.0))
    assert list(e20[:3]) == [2, 4, 7] and list(e20[-2:]) == [116, 128]


def test_dct_matrix_matches_float64_restatement():
    for nb, C, c0 in ((23, 12, 1), (40, 12, 1), (20, 13, 0)):
        m = afe.build_dct(afe.make_params(num_banks=nb, ceps_len=C, want_c0=c0))
        np.testing.assert_allclose(m, oracle_np.dct_matrix(nb, C, c0, 22.0), atol=3e-5)  # float cosf argument, lifter up to 12x
        if c0:
            np.testing.assert_allclose(m[:, C], np.sqrt(2.0 / nb), rtol=1e-7)  # c0 is the LAST column


def test_window_matches_reference_formula():
    w = afe.make_window(400)
    i = np.arange(400)
    np.testing.assert_allclose(w, (0.56 - 0.46 * np.cos(2 * np.pi * i / 400)) / 32768, rtol=2e-7)  # 0.56, divisor W (Q7)


def test_frame_count_rule_matches_float_form():
    """parambase.cpp:16-19 evaluates in float32; it must agree with integer floor-div at every tested size."""
    for W, S in ((400, 160), (200, 80)):
        for n in (399, 400, 54682, 81000, 160000, 16_000_000, 28_800_000, 28_799_880):
            got = afe.estimated_window_count(n, W, S)
            f32 = int(np.floor(np.float32(n - (W - S)) / np.float32(S)))
            assert got == f32
            if n < (1 << 24):
                assert got == max((n - (W - S)) // S, -1) or n < W - S
    assert afe.estimated_window_count(54682, 400, 160) == 340
    assert afe.estimated_window_count(28_800_000, 200, 80) == 359_998
    p = afe.make_params(dyn=2)
    assert afe.lib().afe_output_width(p) == 39
    assert afe.lib().afe_output_width(afe.make_params(ceps_len=0, dyn=1)) == 46


def test_fft_decomposition_model():
    rng = np.random.default_rng(0)
    for N2 in (512, 256):
        for _ in range(3):
            x = rng.standard_normal(N2)
            np.testing.assert_allclose(fft_model.rfft_model(x), np.fft.rfft(x), atol=1e-10)


def test_shard_utterances_balanced_and_contiguous():
    rng = np.random.default_rng(2)
    lens = rng.integers(16000, 400000, size=1000)
    for n_ranks in (1, 2, 4, 8):
        s = afe.shard_utterances(lens, n_ranks)
        assert s[0] == 0 and s[-1] == 1000 and np.all(np.diff(s) >= 0)
        loads = [lens[s[r]:s[r + 1]].sum() for r in range(n_ranks)]
        assert max(loads) - min(loads) <= 2 * lens.max()
    s = afe.shard_utterances(np.array([5, 5, 5]), 8)      # more ranks than utterances: empty shards allowed
    assert s[-1] == 3 and np.all(np.diff(s) >= 0)


@pytest.mark.parametrize("norm", [1, 2, 3])
def test_cmvn_finalize_host_formulas(norm):
    rng = np.random.default_rng(4)
    x = (rng.standard_normal((500, 7)) * 3 + 2).astype(np.float32).astype(np.float64)
    w = 7
    stats = np.concatenate([x.sum(0), (x * x).sum(0), [500.0], x.min(0), x.max(0)])
    mean, scale = afe.cmvn_finalize_host(norm, w, stats)
    np.testing.assert_allclose(mean, x.mean(0), rtol=1e-6)
    if norm == 2:
        np.testing.assert_allclose(scale, 1 / x.std(0, ddof=1), rtol=1e-6)   # unbiased, normalizercpu.cpp:48-49
    if norm == 3:
        np.testing.assert_allclose(scale, 1 / np.maximum(np.abs(x.min(0) - mean), np.abs(x.max(0) - mean)), rtol=1e-6)


def test_afe_extract_command_line_errors(tmp_path):
    """Host driver argument handling needs no GPU: usage / unknown option / unreadable list exit with code 2."""
    import subprocess
    exe = os.path.join(ol.ROOT, "asr-featext-opencl_b200", "afe_extract")
    if not os.path.exists(exe):
        pytest.skip("afe_extract not built")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "usage: afe_extract" in r.stderr
    r = subprocess.run([exe, "--no-such-option", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "unknown option" in r.stderr
    r = subprocess.run([exe, "--scp", str(tmp_path / "missing.scp")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "can't open list" in r.stderr
    r = subprocess.run([exe, "only_one_file.wav"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2


def test_mfcccuda_compiles_against_the_reference_headers():
    """Drop-in at source level: MfccCuda derives from the REFERENCE's own MfccBase (parambase.h:25: set_alpha is not virtual)
    when built with -DAFE_USE_REFERENCE_HEADERS, and is driven through a ParamBase* (tests/cpp/dropin_check.cpp).
    Only where the reference tree exists (the authoring container); the GPU tier runs the linked binary."""
    import subprocess
    if not os.path.isdir("/root/reference"):
        pytest.skip("no reference tree")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-DAFE_USE_REFERENCE_HEADERS", "-I/root/reference",
                        "-I" + os.path.join(ol.ROOT, "include"), "-I" + os.path.join(ol.ROOT, "asr-featext-opencl_b200", "host"),
                        os.path.join(ol.ROOT, "tests", "cpp", "dropin_check.cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    assert r.returncode == 0, r.stdout
    mirror = open(os.path.join(ol.ROOT, "asr-featext-opencl_b200", "host", "afe_stage_api.hpp")).read()
    assert "virtual void set_alpha" not in mirror and "void set_alpha(float alpha) { m_alpha = alpha; }" in mirror


def test_numpy_oracle_preemphasis_zero_is_identity():
    pcm = (np.random.default_rng(0).standard_normal(8000) * 3000).astype(np.int16)
    p = ol.default_params(dyn="acc")
    a, b = oracle_np.mfcc(pcm, p), oracle_np.mfcc(pcm, dict(p, preemphasis=0.0))
    np.testing.assert_array_equal(a, b)
    assert np.abs(oracle_np.mfcc(pcm, dict(p, preemphasis=0.97)) - a).max() > 0.1


def test_shard_stream_covers_every_frame_once_with_the_reference_halo():
    """afe_shard_stream (f4): contiguous frame ranges, each rank's sample range = its frames + D frames of context per side,
    i.e. the (W - S) + 2 D S carry-over of segmentercpu.cpp:69-73,90-92 around interior cuts."""
    for total, W, S, D, n in ((28_800_000, 200, 80, 6, 8), (81_000, 400, 160, 6, 3), (160_000, 400, 160, 0, 4), (54_682, 400, 160, 2, 2)):
        sh = afe.shard_stream(total, W, S, D, n)
        T = (total - (W - S)) // S
        assert sh["first"][0] == 0 and int(sh["first"][-1] + sh["count"][-1]) == T
        assert np.all(sh["first"][1:] == sh["first"][:-1] + sh["count"][:-1])            # contiguous, no overlap
        for r in range(n):
            c0 = max(0, sh["first"][r] - D)
            c1 = min(T, sh["first"][r] + sh["count"][r] + D)
            assert sh["sample_begin"][r] == c0 * S and sh["local_first"][r] == sh["first"][r] - c0
            assert sh["sample_count"][r] == (c1 - c0 - 1) * S + W
            assert sh["sample_begin"][r] + sh["sample_count"][r] <= total
            assert afe.estimated_window_count(int(sh["sample_count"][r]), W, S) == c1 - c0
        if n > 1:   # two neighbours share exactly (W - S) + 2 D S samples
            overlap = sh["sample_begin"][0] + sh["sample_count"][0] - sh["sample_begin"][1]
            assert overlap == (W - S) + 2 * D * S
    with pytest.raises(afe.AfeError, match="too short"):
        afe.shard_stream(4000, 400, 160, 6, 4)


def test_documents_point_at_files_that_exist():
    """DESIGN.md / INTEGRATION.md / README.md / profiles/README.md cite evidence and sources by path: every cited path that
    names a tracked area of this repo must exist (a stale citation is a claim without evidence)."""
    import re
    root = ol.ROOT
    areas = ("profiles/", "tools/", "tests/", "oracle/", "include/", "asr-featext-opencl_b200/")
    missing = []
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(root, doc)).read()
        for m in re.finditer(r"`([A-Za-z0-9_./\-]+)`", text):
            path = m.group(1)
            if "*" in path or "…" in path or "..." in path:
                continue
            cands = [path]
            if doc.startswith("profiles") and "/" not in path:
                if not re.search(r"\.(json|jsonl|txt|csv)$", path):
                    continue
                cands = [os.path.join("profiles", path)]          # the index names its files without the directory
            elif path.startswith(("csrc/", "host/")):
                cands = [os.path.join("asr-featext-opencl_b200", path)]
            if not cands[0].startswith(areas):
                continue
            if cands[0].startswith("oracle/_ref/") or cands[0].endswith((".so", "afe_extract", "afe_stream_bench", "tc_fft_proto")):
                continue                                           # built artefacts, not tracked
            if not os.path.exists(os.path.join(root, cands[0].rstrip("/"))):
                missing.append((doc, path))
    assert not missing, missing

"""GPU tier: the C++ host mirror (host/afe_stage_api.hpp + host/mfcccuda.hpp) driven by the reference's block loop in
host/afe_extract.cpp — text layout `| %f | %f | ... |` (ASR_OCL.cpp:252-260) incl. the ms/Hz timestamps (Q6)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from common import assert_close, ref_alpha_sweep
from golden_io import load_pcm

pytestmark = pytest.mark.gpu
EXE = os.path.join(ol.ROOT, "asr-featext-opencl_b200", "afe_extract")


def write_wav(path, pcm, sr=16000):
    data = np.ascontiguousarray(pcm, "<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16)
    hdr += b"data" + struct.pack("<I", len(data))
    assert len(hdr) == 44
    open(path, "wb").write(hdr + data)


def run(args):
    if not os.path.exists(EXE):
        pytest.skip("afe_extract not built")
    r = subprocess.run([EXE] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r


@pytest.mark.parametrize("limit", [10_000_000, 16000])
def test_text_output_matches_reference_layout_and_values(oracle, tmp_path, limit):
    pcm = load_pcm()["a1"]
    wav, out = str(tmp_path / "a1.wav"), str(tmp_path / "a1.txt")
    write_wav(wav, pcm)
    run(["--banks", "23", "--ceps", "12", "--c0", "1", "--norm", "1", "--dyn", "2", "--l1", "3", "--l2", "3",
         "--sample-limit", str(limit), wav, out])
    lines = open(out).read().splitlines()
    assert len(lines) == 504
    pat = re.compile(r"^\| -?\d+\.\d{6} \|( -?\d+\.\d{6} \|){39}$")
    assert all(pat.match(l) for l in lines), lines[0]
    vals = np.array([[float(x) for x in l.strip("| ").split(" | ")] for l in lines])
    # Q6: time = 0.5f*window_ms/sr + t*(shift_ms/sr) = 0.00078125 + t*0.000625 (ms divided by Hz)
    np.testing.assert_allclose(vals[:, 0], 0.5 * 25 / 16000 + np.arange(504) * (10 / 16000), atol=1e-6)
    p = ol.default_params(norm="cmn", dyn="acc")
    want = oracle.extract(p, [pcm], sample_limit=limit if limit < len(pcm) else 1 << 22)[0][0]
    got = vals[:, 1:].astype(np.float32)
    assert np.abs(got - want).max() < 5e-4 + 1e-6   # %f keeps 6 decimals
    assert_close(got, np.round(want.astype(np.float64), 6).astype(np.float32), p, "text")


def test_binary_output_and_nist_header(oracle, tmp_path):
    pcm = load_pcm()["sample1"]
    nist = str(tmp_path / "s.wav")
    open(nist, "wb").write(b"NIST_1A\n   1024\n".ljust(1024, b" ") + np.ascontiguousarray(pcm, "<i2").tobytes())  # SPHERE: 1024-byte header
    out = str(tmp_path / "s.bin")
    run(["--banks", "23", "--norm", "0", "--dyn", "0", "--text-output", "0", nist, out])
    got = np.fromfile(out, np.float32).reshape(-1, 13)
    want = oracle.extract(ol.default_params(), [pcm], sample_limit=1 << 22)[0][0]
    assert_close(got, want, ol.default_params(), "binary")


def test_driver_reports_reference_error_strings(tmp_path):
    wav = str(tmp_path / "short.wav")
    write_wav(wav, np.zeros(400 + 160 * 5, np.int16))
    r = subprocess.run([EXE, "--dyn", "2", wav, str(tmp_path / "o.txt")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "window count is too small" in r.stderr


def read_htk(path):
    raw = open(path, "rb").read()
    n, period, size, kind = struct.unpack(">iihh", raw[:12])
    return n, period, size, kind, np.frombuffer(raw[12:], ">f4").reshape(n, size // 4).astype(np.float32)


def test_batch_mode_scp_and_htk(oracle, tmp_path):
    """--batch: all files in ONE fused launch; --scp list; HTK parameter files. Rows must equal the
    per-file streaming driver within the tolerance and the oracle's single-block result (Q1 reproduced)."""
    names = ["a1", "a0001", "a3"]
    pcm = load_pcm()
    scp = tmp_path / "list.scp"
    with open(scp, "w") as f:
        for n in names:
            write_wav(str(tmp_path / f"{n}.wav"), pcm[n])
            f.write(f"{tmp_path / (n + '.wav')} {tmp_path / (n + '.htk')}\n")
    opts = ["--banks", "23", "--ceps", "12", "--c0", "1", "--norm", "1", "--dyn", "2", "--l1", "3", "--l2", "3"]
    r = run(opts + ["--batch", "1", "--htk", "1", "--scp", str(scp)])
    # one fused launch per run of equal tile counts (clustered normalisation): 2-tile a1 + a0001, then 3-tile a3
    m = re.search(r"batch: (\d+) files, \d+ frames, \d+ tile\(s\), (\d+) kernel launch", r.stderr)
    assert m and int(m.group(1)) == len(names) and 1 <= int(m.group(2)) <= 2, r.stderr
    p = ol.default_params(norm="cmn", dyn="acc")
    for n in names:
        rows, period, size, kind, got = read_htk(str(tmp_path / f"{n}.htk"))
        want = oracle.extract(p, [pcm[n]], sample_limit=1 << 22)[0][0]
        assert (rows, period, size) == (len(want), 100000, 4 * 39)
        assert kind == (6 | 0x2000 | 0x100 | 0x200 | 0x800)   # MFCC_0_D_A_Z
        assert_close(got, want, p, n + " batch htk")
        # streaming driver, same file, text output
        txt = str(tmp_path / f"{n}.txt")
        run(opts + [str(tmp_path / f"{n}.wav"), txt])
        vals = np.array([[float(x) for x in l.strip("| ").split(" | ")] for l in open(txt).read().splitlines()])
        assert np.abs(vals[:, 1:] - got).max() < 2e-4 + 1e-6


def test_batch_mode_text_equals_streaming_layout(tmp_path):
    pcm = load_pcm()["a1"]
    wav = str(tmp_path / "a1.wav")
    write_wav(wav, pcm)
    run(["--banks", "23", "--norm", "0", "--dyn", "0", "--batch", "1", wav, str(tmp_path / "b.txt")])
    run(["--banks", "23", "--norm", "0", "--dyn", "0", wav, str(tmp_path / "s.txt")])
    b, s = open(tmp_path / "b.txt").read().splitlines(), open(tmp_path / "s.txt").read().splitlines()
    assert len(b) == len(s) == 504
    assert [l.split(" | ")[0] for l in b] == [l.split(" | ")[0] for l in s]   # timestamps (Q6) byte for byte


def test_alpha_sweep_through_parambase_pointer(oracle, tmp_path):
    """VTLN sweep of the reference driver (ASR_OCL.cpp:198-218,234-243): one set_input, then set_alpha -> apply ->
    get_output_data per alpha through a ParamBase* whose set_alpha is NOT virtual (parambase.h:25); one output file per
    alpha, named like the reference names them. A sweep whose alphas did not reach the device would write identical files."""
    pcm = load_pcm()["sample1"]
    wav = str(tmp_path / "s.wav")
    write_wav(wav, pcm)
    run(["--banks", "23", "--norm", "1", "--dyn", "2", "--text-output", "0", "--alpha", "0.9", "--alpha-max", "1.1",
         "--alpha-step", "0.1", wav, str(tmp_path / "o.bin")])
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("o") and f.endswith(".bin"))
    assert files == ["o0.900000.bin", "o1.000000.bin", "o1.100000.bin"], files
    p = ol.default_params(norm="cmn", dyn="acc")
    alphas = [np.float32(0.9) + i * np.float32(0.1) for i in range(3)]   # float arithmetic of the driver's loop
    want = ref_alpha_sweep(oracle, pcm, p, 10_000_000, alphas)
    outs = []
    for k, f in enumerate(files):
        got = np.fromfile(str(tmp_path / f), np.float32).reshape(-1, 39)
        assert_close(got, want[k], p, f"alpha {alphas[k]}")
        outs.append(got)
    assert np.abs(outs[0] - outs[1]).max() > 0.1 and np.abs(outs[2] - outs[1]).max() > 0.1


DROPIN = os.path.join(ol.ROOT, "oracle", "_ref", "dropin_ref")


def test_dropin_with_the_reference_base_classes(oracle, tmp_path):
    """oracle/_ref/dropin_ref = tests/cpp/dropin_main.cpp compiled against the REFERENCE's parambase.h / mfccbase.h and
    linked with the reference's parambase.cpp / mfccbase.cpp (oracle/Makefile) + MfccCuda over libafe_cuda.so: the object
    drops into the reference's own class hierarchy, and the alpha set through the non-virtual ParamBase::set_alpha reaches
    the GPU (streamed in 16 000-sample blocks, CMN, delta + delta-delta)."""
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/dropin_ref not built (needs the reference tree at build time)")
    pcm = load_pcm()["a1"]
    raw, out = str(tmp_path / "a1.s16"), str(tmp_path / "a1.f32")
    np.ascontiguousarray(pcm, "<i2").tofile(raw)
    r = subprocess.run([DROPIN, raw, out, "16000", "0.9", "1.1", "0.1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out, np.float32).reshape(3, 504, 39)
    # the reference objects fed the same blocks and the same sweep (statistics of the flush block = the LAST alpha's)
    p = ol.default_params(norm="cmn", dyn="acc")
    alphas = [np.float32(0.9) + np.float32(0.1) * i for i in range(3)]
    want = ref_alpha_sweep(oracle, pcm, p, 16000, alphas)
    for k in range(3):
        assert_close(got[k], want[k], p, f"dropin alpha {alphas[k]}")
    assert np.abs(got[0] - got[1]).max() > 0.1

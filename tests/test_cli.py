"""The C front end processing/C/godual.c (record loop of godual_ranging.m:57-130 over the C ABI) and its
MAT-v5 writer.  On CPU it is linked against the test-only emulation library; on the GPU box against the product."""
import json
import os
import subprocess

import numpy as np
import pytest
import scipy.io

import parity_common as pc
from amaranth_twstft_b200 import synth
from amaranth_twstft_b200.api import Correlator
from oracle import correlator as oc

CDIR = os.path.join(pc.ROOT, "processing", "C")
N = 10000


def _build(libdir, libname, out):
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-o", out, os.path.join(CDIR, "godual.c"), f"-L{libdir}",
                           f"-l{libname}", f"-Wl,-rpath,{libdir}", "-lm", "-lpthread"])
    return out


def _run_cli(exe, tmp_path, lib, remote=0, nrec=3, gpus=None):
    bits, code = pc.make_code(N)
    kw = dict(ch1=dict(df=51234.5)) if remote else {}
    recs = [synth.record_pair(code, 700 + i, **kw)[0] for i in range(nrec)]
    data = tmp_path / "1671606372short.bin"
    with open(data, "wb") as f:
        for r in recs:
            f.write(r.tobytes())
        f.write(recs[0][:1000].tobytes())          # trailing partial record: ignored (godual_ranging.m:81,102)
    codef = tmp_path / "code.bin"
    codef.write_bytes(bits.tobytes())
    out = subprocess.run([exe, str(data), str(codef), str(remote), "2"] + ([str(gpus)] if gpus else []),
                         capture_output=True, text=True, check=True)
    rows = [ln.split("\t") for ln in out.stdout.replace("\r", "").strip().splitlines()]
    assert len(rows) == nrec and all(len(r) == (5 if remote else 9) for r in rows)
    assert f"{nrec} records in" in out.stderr and "records/s" in out.stderr          # file -> rows throughput line
    mat = scipy.io.loadmat(str(tmp_path / (("remote" if remote else "") + "1671606372short.mat")))
    win = (80000.0, 120000.0) if remote else (-20000.0, 20000.0)
    with Correlator(nsamples=N, window_hz=win, max_batch=2, lib=lib, fast_peak=True) as cor:   # like godual.c
        cor.set_code(bits)
        ref = cor.process(np.stack(recs))
    for ch in ((1,) if remote else (1, 2)):
        for var, field in (("indice", "indice"), ("correction", "correction"), ("SNR%dr", "SNRr"), ("SNR%di", "SNRi"),
                           ("df", "df"), ("puissance", "puissance"), ("puissance%dcode", "puissancecode"),
                           ("puissance%dnoise", "puissancenoise")):
            name = var % ch if "%d" in var else f"{var}{ch}"
            assert mat[name].shape == (1, nrec), name
            assert np.array_equal(mat[name][0], ref[field][:, ch - 1]), name
    if remote:
        assert "indice2" not in mat
    assert np.array_equal(mat["code"][0], code)
    for i, r in enumerate(rows):
        assert int(r[0]) == i + 1
        delay = (ref["indice"][i, 0] - 1 + ref["correction"][i, 0]) / 5e6 / 3
        assert abs(float(r[1]) - delay) < 1e-12 and abs(float(r[2]) - ref["df"][i, 0]) < 1e-3
    # and the stored values are the reference's: oracle check of the first record
    freq = oc.freq_axis_G(5e6, N)
    o = oc.record_G(recs[0], oc.window(freq, *win), freq, oc.fcode_G(code), code)
    assert int(mat["indice1"][0, 0]) == o[1]["indice"]


@pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
@pytest.mark.parametrize("remote", [0, 1])
def test_godual_cli_emulation(tmp_path, remote):
    exe = _build(os.path.dirname(pc.EMU_LIB), "twstft_emu", str(tmp_path / "godual_emu"))
    _run_cli(exe, tmp_path, pc.EMU_LIB, remote)


@pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
def test_godual_cli_sharded_with_overlapped_ingest(tmp_path):
    """Seven records, chunks of batch*gpus = 4 through two plans: the reader thread refills a slot while the other is
    processed, rows stay in file order, the short last chunk and the trailing partial record end the loop."""
    exe = _build(os.path.dirname(pc.EMU_LIB), "twstft_emu", str(tmp_path / "godual_emu"))
    _run_cli(exe, tmp_path, pc.EMU_LIB, 0, nrec=7, gpus=2)


@pytest.mark.gpu
def test_godual_cli_cuda(tmp_path):
    exe = _build(os.path.dirname(pc.CUDA_LIB), "twstft", str(tmp_path / "godual"))
    _run_cli(exe, tmp_path, pc.CUDA_LIB, 0)


# ---------------------------------------------------------------------------------------------------------------
# processing/C/godual_cpp.c: the C++ program's flavour (processing/CPP/main.cpp) -- whole-file estimator, Hamming-windowed
# code spectrum, <name>C.mat with [P x 1] variables
def _run_cpp_cli(libdir, libname, lib, tmp_path, nrec=25, remote=0):
    exe = str(tmp_path / "godual_cpp")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-o", exe, os.path.join(CDIR, "godual_cpp.c"), f"-L{libdir}",
                           f"-l{libname}", f"-Wl,-rpath,{libdir}", "-lm"])
    bits, code = pc.make_code(N)
    recs = [synth.record_pair(code, 900 + i, ch1=dict(df=2345.6, delay=3000.0 + 0.1 * i), ch2=dict(df=-321.0))[0]
            for i in range(nrec)]
    data = tmp_path / "1671606372short.bin"
    with open(data, "wb") as f:
        for r in recs:
            f.write(r.tobytes())
    codef = tmp_path / "code.bin"
    codef.write_bytes(bits.tobytes())
    out = subprocess.run([exe, str(data), str(codef), str(remote)], capture_output=True, text=True, check=True)
    lines = out.stdout.splitlines()
    assert lines[-1] == "No more data" and any(ln.startswith("df1=") for ln in lines)
    mat = scipy.io.loadmat(str(tmp_path / (("remote" if remote else "") + "1671606372shortC.mat")))
    raw = np.concatenate(recs)
    dfs = oc.df_CPP(raw, 2, remote=remote)
    fw = oc.fcode_CPP(code)
    for ch in ((1,) if remote else (1, 2)):
        for nm in ("correction", "SNR", "df", "puissance"):
            assert mat[f"{nm}{ch}"].shape == (nrec, 1), nm
        assert mat[f"xval{ch}"].dtype == np.complex128 and mat[f"xval{ch}m1"].shape == (nrec, 1)
        assert np.all(mat[f"df{ch}"][:, 0] == dfs[ch])
        for i in (0, nrec - 1):
            d = oc.unpack_interleaved2(recs[i])[ch - 1]
            o = oc.processing_CPP(d, dfs[ch], fw, code)
            assert abs(mat[f"correction{ch}"][i, 0] - o["correction"]) < 3e-4
            assert abs(mat[f"SNR{ch}"][i, 0] - o["SNR"]) < 0.01
            assert abs(mat[f"puissance{ch}code"][i, 0] - o["puissancecode"]) < 0.01
            assert abs(mat[f"xval{ch}"][i, 0] - o["xval"]) < 1e-4 * abs(o["xval"])
            assert abs(mat[f"xval{ch}p1"][i, 0] - o["xvalp1"]) < 1e-4 * abs(o["xval"])
    if remote:
        assert "correction2" not in mat


@pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
@pytest.mark.parametrize("remote", [0, 1])
def test_godual_cpp_cli_emulation(tmp_path, remote):
    _run_cpp_cli(os.path.dirname(pc.EMU_LIB), "twstft_emu", pc.EMU_LIB, tmp_path, remote=remote)


@pytest.mark.gpu
def test_godual_cpp_cli_cuda(tmp_path):
    _run_cpp_cli(os.path.dirname(pc.CUDA_LIB), "twstft", pc.CUDA_LIB, tmp_path, nrec=50)


# ---------------------------------------------------------------------------------------------------------------
# processing/C/claudio_aligned.c: the K-variant production loops (lo / re / ranging scripts) in C over the C ABI
def _build_k(libdir, libname, out):
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-o", out, os.path.join(CDIR, "claudio_aligned.c"), f"-L{libdir}",
                           f"-l{libname}", f"-Wl,-rpath,{libdir}", "-lm"])
    return out


def _run_k_cli(exe, tmp_path, lib, mode, fs=250000.0, n=N, nblocks=2.3, batch=8):
    from amaranth_twstft_b200.ktrack import KTracker, PREFIX
    bits, code = pc.make_code(n)
    df = {"lo": 1234.5, "re": 50321.5, "ranging": -2222.25}[mode]
    skip_s = 2.0
    raw = synth.stream_single(code, int(nblocks * 2 * fs), 9, delay=0.3 * n + 0.4, df=df, fs=fs, dc=(30.0, -12.0))
    data = tmp_path / "1716807811_1716808150.994543932_1.bin"
    data.write_bytes(raw.tobytes())
    codef = tmp_path / "n_code.bin"
    codef.write_bytes(bits.tobytes())
    out = subprocess.run([exe, mode, str(data), str(codef), "0", str(tmp_path), repr(fs), repr(skip_s), str(batch)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    mat = scipy.io.loadmat(str(tmp_path / (PREFIX[mode] + "1716807811_1716808150.994543932_1.mat")))
    with KTracker(bits, fs=fs, ls=2, max_batch=batch, lib=lib, mode=mode, skip_s=skip_s) as kt:
        ref = kt.process_file(raw)
    assert ref is not None and not ref["status"] and len(ref["moved"]) >= 1
    # same library underneath: every stored number is identical, in the order/shape of the reference's own files
    schema = json.load(open(os.path.join(pc.ROOT, "tests", "golden", "mat_schema.json")))
    names = list(next(v for k, v in schema.items() if "localclaudio" in k))
    assert [k for k in mat if not k.startswith("__")] == names
    for name in ("correction1", "df", "indice1", "SNR1i", "SNR1r", "puissance1", "xval1", "moved", "movedval", "code"):
        assert mat[name].shape == (1, len(ref[name])), (name, mat[name].shape)
        assert np.array_equal(mat[name][0], ref[name]), name
    assert mat["xval1"].dtype == np.complex128
    assert mat["puissancecode"][0, 0] == ref["puissancecode"] and mat["puissancenoise"][0, 0] == ref["puissancenoise"]
    lines = out.stdout.replace("\r", "").strip().splitlines()
    assert lines[0].startswith("n\tdt1\tdf1\tP1\tSNR1")
    rows = [ln.split("\t") for ln in lines[1:] if not ln.startswith("MOVED")]
    assert len(rows) == len(ref["indice1"]) and sum(ln.startswith("MOVED") for ln in lines) == len(ref["moved"])
    i = len(rows) - 1
    assert abs(float(rows[i][1]) - (ref["indice1"][i] - 1 + ref["correction1"][i]) / fs / 3) < 1e-12
    return ref


@pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
@pytest.mark.parametrize("mode", ["lo", "re", "ranging"])
def test_claudio_aligned_cli_emulation(tmp_path, mode):
    exe = _build_k(os.path.dirname(pc.EMU_LIB), "twstft_emu", str(tmp_path / "claudio_aligned_emu"))
    _run_k_cli(exe, tmp_path, pc.EMU_LIB, mode)


@pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
def test_claudio_aligned_cli_reports_missing_carrier(tmp_path):
    """re/ranging on noise: the script would loop for ever; the build exits with status 2."""
    exe = _build_k(os.path.dirname(pc.EMU_LIB), "twstft_emu", str(tmp_path / "claudio_aligned_emu"))
    rng = np.random.default_rng(5)
    raw = rng.normal(0, 3000, int(2.2 * 2 * 250000) * 2).round().astype(np.int16)
    (tmp_path / "noise_2.bin").write_bytes(raw.tobytes())
    (tmp_path / "n_code.bin").write_bytes(pc.make_code(N)[0].tobytes())
    out = subprocess.run([exe, "re", str(tmp_path / "noise_2.bin"), str(tmp_path / "n_code.bin"), "0", str(tmp_path),
                          "250000", "0.5", "8"], capture_output=True, text=True)
    assert out.returncode == 2 and "carrier not found" in out.stderr
    assert not (tmp_path / "remoteclaudionoise_2.mat").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lo", "re"])
def test_claudio_aligned_cli_cuda(tmp_path, mode):
    exe = _build_k(os.path.dirname(pc.CUDA_LIB), "twstft", str(tmp_path / "claudio_aligned"))
    _run_k_cli(exe, tmp_path, pc.CUDA_LIB, mode)

"""processing/C/twstft_mex.c (the drop-in for the Octave functions processing() and search_df()) compiled against
the stand-in MEX runtime of tests/mex_stub and RUN: mexFunction receives the arguments and the script-level globals
exactly as godual_ranging.m / claudio_aligned_code_*_separate.m would pass them; outputs against the oracle.
CPU: linked with the test-only host emulation build; GPU (-m gpu): linked with the CUDA product library."""
import os
import struct
import subprocess

import numpy as np
import pytest

import parity_common as pc
from amaranth_twstft_b200 import synth
from oracle import correlator as oc

ROOT = pc.ROOT
STUB = os.path.join(ROOT, "tests", "mex_stub")


def _host(kind):
    exe = os.path.join(STUB, "_build", f"mex_host_{kind}")
    lib = pc.EMU_LIB if kind == "emu" else pc.CUDA_LIB
    if not os.path.exists(lib):
        pytest.skip(f"{lib} not built")
    subprocess.run(["make", "-C", STUB, f"_build/mex_host_{kind}"], check=True, capture_output=True)
    return exe


def _blob(path, args, globs):
    with open(path, "wb") as f:
        for isg, items in ((0, args), (1, globs)):
            for name, v in items:
                v = np.atleast_1d(np.asarray(v))
                cplx = np.iscomplexobj(v)
                f.write(name.encode().ljust(32, b"\0"))
                f.write(struct.pack("<iiq", isg, int(cplx), v.size))
                f.write(np.ascontiguousarray(v.real, dtype=np.float64).tobytes())
                if cplx:
                    f.write(np.ascontiguousarray(v.imag, dtype=np.float64).tobytes())


def _call(exe, fname, nlhs, args, globs, tmp_path):
    blob = str(tmp_path / f"{fname}.blob")
    _blob(blob, args, globs)
    out = subprocess.run([exe, fname, str(nlhs), blob], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vals, cur = [], None
    for line in out.stdout.splitlines():
        if line.startswith("output"):
            cur = []
            vals.append(cur)
        else:
            re, im = (float(t) for t in line.split())
            cur.append(complex(re, im))
    return [v[0] if len(v) == 1 else np.array(v) for v in vals]


def _globals(code, freq, fs):
    n = len(code)
    return [("fs", fs), ("Nint", 1.0), ("code", code), ("freq", freq), ("temps", np.arange(n) / fs)]


def _check_processing(kind, n, tmp_path):
    exe = _host(kind)
    fs = 5e6
    bits, code = pc.make_code(n)
    # variant G: d=d-mean(d); [indice,correction,SNRr,SNRi,df,puissance,puissancecode,puissancenoise]=processing(d,k)
    d1, _ = oc.unpack_interleaved2(synth.record_pair(code, 61, fs=fs)[0])
    d = d1 - d1.mean()
    freq = oc.freq_axis_G(fs, n)
    k0 = oc.window(freq, -20000.0, 20000.0)
    o = _call(exe, "processing", 8, [("d", d), ("k", (k0 + 1).astype(float))], _globals(code, freq, fs), tmp_path)
    ref = oc.processing_G(d, k0, freq, oc.fcode_G(code), code, fs)
    assert int(o[0].real) == ref["indice"]
    assert abs(o[1].real - ref["correction"]) / 3 < pc.TOL_DELAY_SAMPLES
    assert abs(10 * np.log10(o[2].real + o[3].real) - 10 * np.log10(ref["SNRr"] + ref["SNRi"])) < pc.TOL_SNR_DB
    assert o[4].real == ref["df"]
    for got, nm in zip(o[5:8], ("puissance", "puissancecode", "puissancenoise")):
        assert abs(got.real / ref[nm] - 1) < 1e-3, nm
    # variant K: [xval,indice,correction,...]=processing(d,k,df)  and  processing(d,df)
    dfk = 1234.5
    raw, _ = synth.record_single(code, 62, delay=777.25, df=dfk, snr_db=-5.0, fs=fs)
    dk = oc.unpack_single(raw)
    dk = dk - dk.mean()
    refk = oc.processing_K(dk, dfk, oc.fcode_K(code), code, fs)
    fk = oc.freq_axis_K(fs, 1) if n == int(fs) else np.linspace(-fs / 2, fs / 2 - 1.0, n)
    for args in ([("d", dk), ("k", (k0 + 1).astype(float)), ("df", dfk)], [("d", dk), ("df", dfk)]):
        o = _call(exe, "processing", 8, args, _globals(code, fk, fs), tmp_path)
        assert int(o[1].real) == refk["indice"]
        assert abs(o[2].real - refk["correction"]) / 3 < pc.TOL_DELAY_SAMPLES
        assert abs(np.angle(o[0] * np.conj(refk["xval"]))) < pc.TOL_PHASE_RAD
        assert abs(10 * np.log10(o[3].real + o[4].real) - 10 * np.log10(refk["SNRr"] + refk["SNRi"])) < pc.TOL_SNR_DB


def _check_search_df(kind, tmp_path, fs, n, df):
    exe = _host(kind)
    bits, code = pc.make_code(n)
    raw = synth.stream_single(code, int(2 * fs), 12, delay=1234.25, df=df, fs=fs, dc=(10.0, 5.0))
    d = oc.unpack_single(raw)
    freq = oc.freq_axis_K(fs, 2)
    k0 = oc.window(freq, 92000.0, 108000.0)
    o = _call(exe, "search_df", 1, [("d", d), ("k", (k0 + 1).astype(float)), ("thr", 20.0)],
              [("fs", fs), ("Nint", 1.0), ("code", code), ("freq", freq)], tmp_path)
    kref, tried = oc.search_df(d, k0, 20.0, freq, oc.fcode_K(code), fs=fs)
    assert kref > 0 and int(o[0].real) == kref, (o, kref, tried)


def test_mex_processing_signatures_on_the_emulated_kernels(tmp_path):
    _check_processing("emu", 10000, tmp_path)


def test_mex_search_df_on_the_emulated_kernels(tmp_path):
    _check_search_df("emu", tmp_path, 250000.0, 10000, 50321.5)


def test_mex_reports_missing_globals(tmp_path):
    exe = _host("emu")
    blob = str(tmp_path / "x.blob")
    _blob(blob, [("d", np.zeros(10000, dtype=complex)), ("k", np.arange(1.0, 5.0))], [("fs", 5e6)])
    out = subprocess.run([exe, "processing", "8", blob], capture_output=True, text=True)
    assert out.returncode == 3 and "global variable" in out.stderr


@pytest.mark.gpu
def test_mex_processing_signatures_on_the_gpu(tmp_path):
    _check_processing("cuda", 200000, tmp_path)


@pytest.mark.gpu
def test_mex_search_df_on_the_gpu(tmp_path):
    _check_search_df("cuda", tmp_path, 5e6, 200000, 50321.5)

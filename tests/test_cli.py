"""The C front end processing/C/godual.c (record loop of godual_ranging.m:57-130 over the C ABI) and its
MAT-v5 writer.  On CPU it is linked against the test-only emulation library; on the GPU box against the product."""
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
                           f"-l{libname}", f"-Wl,-rpath,{libdir}", "-lm"])
    return out


def _run_cli(exe, tmp_path, lib, remote=0):
    bits, code = pc.make_code(N)
    kw = dict(ch1=dict(df=51234.5)) if remote else {}
    recs = [synth.record_pair(code, 700 + i, **kw)[0] for i in range(3)]
    data = tmp_path / "1671606372short.bin"
    with open(data, "wb") as f:
        for r in recs:
            f.write(r.tobytes())
        f.write(recs[0][:1000].tobytes())          # trailing partial record: ignored (godual_ranging.m:81,102)
    codef = tmp_path / "code.bin"
    codef.write_bytes(bits.tobytes())
    out = subprocess.run([exe, str(data), str(codef), str(remote), "2"], capture_output=True, text=True, check=True)
    rows = [ln.split("\t") for ln in out.stdout.replace("\r", "").strip().splitlines()]
    assert len(rows) == 3 and all(len(r) == (5 if remote else 9) for r in rows)
    mat = scipy.io.loadmat(str(tmp_path / (("remote" if remote else "") + "1671606372short.mat")))
    win = (80000.0, 120000.0) if remote else (-20000.0, 20000.0)
    with Correlator(nsamples=N, window_hz=win, max_batch=2, lib=lib) as cor:
        cor.set_code(bits)
        ref = cor.process(np.stack(recs))
    for ch in ((1,) if remote else (1, 2)):
        for var, field in (("indice", "indice"), ("correction", "correction"), ("SNR%dr", "SNRr"), ("SNR%di", "SNRi"),
                           ("df", "df"), ("puissance", "puissance"), ("puissance%dcode", "puissancecode"),
                           ("puissance%dnoise", "puissancenoise")):
            name = var % ch if "%d" in var else f"{var}{ch}"
            assert mat[name].shape == (1, 3), name
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


@pytest.mark.gpu
def test_godual_cli_cuda(tmp_path):
    exe = _build(os.path.dirname(pc.CUDA_LIB), "twstft", str(tmp_path / "godual"))
    _run_cli(exe, tmp_path, pc.CUDA_LIB, 0)

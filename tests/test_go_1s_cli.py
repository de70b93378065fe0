"""processing/C/go_1s.c -- the per-second reducer of acquisition/go_1s.m in C -- against its numpy oracle
(oracle/reduce.py) on synthetic K-variant result files in the three container flavours found in the field:
plain MAT-v5 (Octave `save -mat`, our C writer), whole-file gzip (the reference's *.mat.gz) and element-compressed."""
import gzip
import os
import subprocess

import numpy as np
import pytest
import scipy.io

from oracle import reduce as orx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CDIR = os.path.join(ROOT, "processing", "C")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("go1s") / "go_1s")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-o", out, os.path.join(CDIR, "go_1s.c"), "-lz", "-lm"])
    return out


def _kfile(rng, n, base_ns, slope_ns_s, lead=4, tail=3, gap_at=None, jump_at=None, fs=5e6):
    """A K-variant result record: weak windows at both ends, delay base + slope*t (+ noise), 25 windows/s."""
    t = np.arange(n) / 25.0
    delay_ns = base_ns + slope_ns_s * t + 0.05 * rng.standard_normal(n)
    if jump_at is not None:
        delay_ns[jump_at:] += 200.0                     # one lost sample
    samples = delay_ns * 1e-9 * fs
    indice1 = np.floor(samples)
    correction1 = (samples - indice1) * 3.0
    amp = np.full(n, 1.0) + 0.01 * rng.standard_normal(n)
    amp[:lead] = 0.05
    amp[n - tail:] = 0.05
    if gap_at is not None:
        amp[gap_at:gap_at + 2] = 0.05
    xval1 = amp * np.exp(1j * rng.uniform(0, 2 * np.pi, n))
    snr = 10 ** (-1.5 + 0.05 * rng.standard_normal(n))
    return dict(correction1=correction1[None], df=np.zeros((1, n // 50 + 1)), indice1=indice1[None],
                SNR1i=(0.4 * snr)[None], SNR1r=(0.6 * snr)[None], code=np.ones((1, 16)), puissance1=np.ones((1, n)),
                puissancecode=np.ones((1, 1)), puissancenoise=np.ones((1, 1)), xval1=xval1[None],
                moved=np.zeros((1, 0)), movedval=np.zeros((1, 0)))


def _flat(m):
    return {k: np.asarray(v).ravel() for k, v in m.items()}


def _write(tmp_path, mats, flavours):
    paths = []
    for i, (m, fl) in enumerate(zip(mats, flavours)):
        p = str(tmp_path / f"k{i}.mat")
        scipy.io.savemat(p, m, format="5", do_compression=(fl == "zelem"))
        if fl == "gz":
            with open(p, "rb") as f, gzip.open(p + ".gz", "wb") as g:
                g.write(f.read())
            os.remove(p)
            p += ".gz"
        paths.append(p)
    return paths


def _check(exe, tmp_path, mats, flavours, mjd=60457.4321):
    paths = _write(tmp_path, mats, flavours)
    out = str(tmp_path / "out.1s")
    r = subprocess.run([exe, *paths, repr(mjd), out], capture_output=True, text=True)
    ref = orx.session(*[_flat(m) for m in mats])
    if ref is None:
        assert r.returncode == 3, (r.returncode, r.stderr)
        return None, r
    assert r.returncode == 0, r.stderr
    rows = orx.one_second_rows(ref, mjd)
    lines = open(out).read().splitlines()
    assert lines[0] == "# MJD\t\tOPlocal\tOPremote\tLTFBlocal\tLTBBremote"
    assert len(lines) == 1 + len(rows) and len(rows) >= 3
    for ln, row in zip(lines[1:], rows):
        got = [float(x) for x in ln.split("\t")]
        assert ln.split("\t")[0] == "%f" % row[0]
        assert np.allclose(got[1:], row[1:], rtol=0, atol=2e-6), (got, row)       # "%f": 1e-6 ns resolution
    summary = dict(kv.split("=") for kv in r.stdout.strip().splitlines()[-1].split())
    res = ref["res"]
    assert int(summary["n"]) == len(res) and int(summary["nan"]) == int(np.isnan(res).sum())
    assert abs(float(summary["resmean"]) - np.nanmean(res)) < 1e-5
    assert abs(float(summary["resstd"]) - np.nanstd(res, ddof=1)) < 1e-5
    assert abs(float(summary["snrop"]) - ref["snrop"]) < 1e-3 and abs(float(summary["snrlt"]) - ref["snrlt"]) < 1e-3
    return ref, r


def test_plain_session(exe, tmp_path):
    rng = np.random.default_rng(1)
    mats = [_kfile(rng, 400, 1000.0, 0.3), _kfile(rng, 400, 2.6e5, 0.5), _kfile(rng, 400, 1200.0, -0.2),
            _kfile(rng, 400, 2.6e5 + 31.0, 0.4)]
    ref, r = _check(exe, tmp_path, mats, ["plain", "gz", "zelem", "plain"])
    assert len(ref["oplo"]) == 400 - 4 - 3 - 11                     # k(11:end-1) of the strong windows


def test_gaps_sample_loss_and_truncations(exe, tmp_path):
    rng = np.random.default_rng(2)
    mats = [_kfile(rng, 500, 1000.0, 0.3, jump_at=420), _kfile(rng, 500, 2.6e5, 0.5, gap_at=330),
            _kfile(rng, 500, 1200.0, -0.2, gap_at=460), _kfile(rng, 500, 2.6e5 + 31.0, 0.4, tail=60)]
    ref, r = _check(exe, tmp_path, mats, ["plain", "plain", "gz", "plain"])
    assert "Sample loss" in r.stdout and "truncated" in r.stdout
    assert len(ref["oplo"]) == len(ref["ltlo"]) == len(ref["opre"]) == len(ref["ltre"])


def test_outliers_become_nan_and_short_sessions_are_skipped(exe, tmp_path):
    rng = np.random.default_rng(3)
    mats = [_kfile(rng, 300, 1000.0, 0.3), _kfile(rng, 300, 2.6e5, 0.5), _kfile(rng, 300, 1200.0, -0.2),
            _kfile(rng, 300, 2.6e5 + 31.0, 0.4)]
    mats[1]["correction1"][0, 150] += 3 * 60.0 * 5e-3              # one window 60 ns off -> |res - median| > 5
    ref, _ = _check(exe, tmp_path, mats, ["plain"] * 4)
    assert np.isnan(ref["res"]).sum() == 1
    short = [_kfile(rng, 110, 1000.0, 0.3), _kfile(rng, 110, 2.6e5, 0.5), _kfile(rng, 110, 1200.0, -0.2),
             _kfile(rng, 110, 2.6e5, 0.4)]
    ref, r = _check(exe, tmp_path, short, ["plain"] * 4)
    assert ref is None and "too short" in r.stderr


def test_reads_the_c_writer(exe, tmp_path):
    """matv5.h (what claudio_aligned writes) -> matv5_read.h round trip through a tiny C program."""
    src = tmp_path / "rt.c"
    src.write_text('#include "matv5.h"\n#include "matv5_read.h"\n'
                   'int main(int c,char**v){double re[3]={1.5,-2,3e9},im[3]={0.25,0,-7};FILE*f=fopen(v[1],"wb");\n'
                   'mat5_open(f,"t");mat5_vector(f,"xval1",re,im,3);mat5_vector(f,"moved",re,NULL,0);fclose(f);\n'
                   'mat5_file m;if(mat5_load(v[1],&m))return 2;const mat5_var*x=mat5_find(&m,"xval1");\n'
                   'if(!x||x->n!=3||x->re[2]!=3e9||x->im[0]!=0.25||x->im[2]!=-7)return 3;\n'
                   'const mat5_var*e=mat5_find(&m,"moved");if(!e||e->n!=0)return 4;mat5_free(&m);return 0;}\n')
    exe2 = str(tmp_path / "rt")
    subprocess.check_call(["gcc", "-O2", "-std=c11", f"-I{CDIR}", "-o", exe2, str(src), "-lz", "-lm"])
    assert subprocess.run([exe2, str(tmp_path / "rt.mat")]).returncode == 0
    m = scipy.io.loadmat(str(tmp_path / "rt.mat"))
    assert m["xval1"][0, 2] == 3e9 - 7j

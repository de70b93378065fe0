"""processing/C/gofinal.c (host-only): the text reducers processing/Octave/gofinal_{op,ltfb}.m and tw.m in C, checked
against their numpy restatement (oracle/reduce.py) on synthetic result files in the Octave and the C++ flavours."""
import gzip
import json
import os
import subprocess
import time

import numpy as np
import pytest
import scipy.io

import parity_common as pc
from oracle import reduce as orr

CDIR = os.path.join(pc.ROOT, "processing", "C")
END = 1671606372


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gofinal") / "gofinal")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-o", out, os.path.join(CDIR, "gofinal.c"), "-lz", "-lm"])
    return out


def _files(rng, n, jump_at=None, cflavour=False):
    def one(two, base, slope):
        m = {}
        for c in ((1, 2) if two else (1,)):
            sol = base * (1.0 if c == 1 else 0.7) + slope * np.arange(n) + rng.normal(0, 0.02, n)
            if jump_at is not None and c == 1 and two:
                sol[jump_at:] += 200.0                      # 13 us step: the two-cluster branch
            snr = np.full(n, 1e-3)
            snr[:7] = 1e-6                                   # TX inactive at the beginning
            if cflavour:
                m[f"correction{c}"] = sol[:, None]
                m[f"SNR{c}"] = (10 * np.log10(snr))[:, None]
                m[f"df{c}"] = rng.normal(100 * c, 0.3, n)[:, None]
            else:
                m[f"indice{c}"] = np.floor(sol)[None, :]
                m[f"correction{c}"] = (sol - np.floor(sol))[None, :]
                m[f"SNR{c}r"] = (snr * 0.7)[None, :]
                m[f"SNR{c}i"] = (snr * 0.3)[None, :]
                m[f"df{c}"] = rng.normal(100 * c, 0.3, n)[None, :]
        if not cflavour:
            m["code"] = np.ones((1, 5000000))
        return m
    return one(True, 7.5e6, 0.3), one(False, 3.1e6, -0.2)


def _fmt(rows):
    out = []
    for r in rows:
        stamp = time.strftime("%Y %m %d %H %M %S", time.gmtime(r[0]))
        out.append("%s\t%.12f\t%.3f\t%.1f\t%.12f\t%.3f\t%.1f\t%.12f\t%.3f\t%.1f" % ((stamp,) + tuple(r[1:])))
    return out


def _run(exe, tmp_path, mode, lo, re, lname, rname, gz=False, **flav):
    for name, m in ((lname, lo), (rname, re)):
        scipy.io.savemat(str(tmp_path / name), m)
        if gz:
            raw = (tmp_path / name).read_bytes()
            (tmp_path / name).unlink()
            with gzip.open(str(tmp_path / (name + ".gz")), "wb") as f:
                f.write(raw)
    sfx = ".gz" if gz else ""
    out = subprocess.run([exe, mode, lname + sfx, rname + sfx, str(END), f"{mode}.txt"], cwd=str(tmp_path),
                         capture_output=True, text=True, check=True, env=dict(os.environ, TZ="UTC"))
    rows, stats = orr.gofinal_rows(lo, re, END, mode == "ltfb", **flav)
    text = (tmp_path / f"{mode}.txt").read_bytes().decode()
    head, body = text.split("\r\n", 1)
    assert head == "% y  m  d  h  m  s\tdelay\t\tdf1\tSNR1\tdf2\tSNR2\tdelayrem\t\tdf1rem\tSNR1rem"
    assert body.split("\n")[:-1] == _fmt(rows)
    tag = "LTFB" if mode == "ltfb" else "OP"
    std = float(out.stdout.split(f"std{tag}=")[1].split("\t")[0])
    moy = float(out.stdout.split(f"<{tag}>=")[1])
    assert abs(std - stats[0]) < 1e-5 * max(1.0, stats[0]) and abs(moy - stats[1]) < 1e-6
    return rows


@pytest.mark.parametrize("mode", ["op", "ltfb"])
def test_octave_flavour(exe, tmp_path, mode):
    lo, re = _files(np.random.default_rng(3), 200)
    rows = _run(exe, tmp_path, mode, lo, re, "1671606192.mat", "remote1671606192.mat", gz=(mode == "ltfb"))
    assert len(rows) == (200 if mode == "op" else 193)        # ltfb keeps SNR1 > max-10 dB only


def test_op_mode_ignores_a_C_in_the_path_and_dates_from_the_remote_file(exe, tmp_path):
    """gofinal_op.m has no C++-flavour branch (only gofinal_ltfb.m:35 does) and takes dateini from length(indice1) of
    the REMOTE file (gofinal_op.m:36-37): a directory named Claudio/ must not flip the flavour, and a remote file
    with more records than the local one shifts every epoch."""
    lo, re = _files(np.random.default_rng(8), 200)
    _, re_long = _files(np.random.default_rng(9), 230)
    sub = tmp_path / "Claudio"
    sub.mkdir()
    rows = _run(exe, sub, "op", lo, re_long, "1671606192.mat", "remote1671606192.mat")
    assert len(rows) == 200 and rows[0][0] == int(END - 230)


def test_cpp_flavour_files(exe, tmp_path):
    """A 'C' in the path selects correction = indice0+corr and SNR in dB (gofinal_ltfb.m:35-45)."""
    lo, re = _files(np.random.default_rng(4), 180, cflavour=True)
    _run(exe, tmp_path, "ltfb", lo, re, "1671606192C.mat", "remote1671606192C.mat", local_c=True, remote_c=True)


def test_two_cluster_selection_keeps_the_reference_quirk(exe, tmp_path):
    lo, re = _files(np.random.default_rng(5), 260, jump_at=120)
    rows = _run(exe, tmp_path, "ltfb", lo, re, "1671606192.mat", "remote1671606192.mat")
    assert 100 < len(rows) < 200


def test_two_way_combination(exe, tmp_path):
    rng = np.random.default_rng(6)

    def table(start, n, d10, d13, snr0):
        t = np.zeros((n, 15))
        for i in range(n):
            g = time.gmtime(start + i)
            t[i, :6] = (g.tm_year, g.tm_mon, g.tm_mday, g.tm_hour, g.tm_min, g.tm_sec)
        t[:, 8] = snr0
        t[:5, 8] = snr0 - 30
        t[-4:, 8] = snr0 - 30
        t[:, 9] = d10 + rng.normal(0, 2e-10, n)
        t[:, 12] = d13 + rng.normal(0, 2e-10, n)
        t[:, 6] = 0.5
        return t
    op = table(END, 170, 0.2500003, 0.7253001, -5.0)
    lt = table(END + 7, 170, 0.7747005, 0.2500001, -6.0)      # l10-l13 = 0.5247004; op wraps by +1 s to 0.5247002
    for name, t in (("a.b.c.txt", op), ("x_ltfb.txt", lt)):
        with open(tmp_path / name, "w") as f:
            f.write("% y  m  d  h  m  s\tdelay\r\n")
            for r in t:
                f.write("%d %02d %02d %02d %02d %02d\t" % tuple(r[:6]) + "\t".join("%.12f" % v for v in r[6:]) + "\n")
    out = subprocess.run([exe, "tw", "a.b.c.txt", "x_ltfb.txt"], cwd=str(tmp_path), capture_output=True, text=True,
                         check=True).stdout
    op_r, lt_r = (np.loadtxt(tmp_path / n, comments="%") for n in ("a.b.c.txt", "x_ltfb.txt"))
    d, std, mean, n = orr.two_way(op_r, lt_r)
    assert f"diff={d}" in out and d == 7
    assert abs(float(out.split("std=")[1].split()[0]) - std) < 1e-5
    assert abs(float(out.split("mean=")[1].split()[0]) - mean) < 1e-5
    assert int(out.split("\nn=")[1]) == n


# ---------------------------------------------------------------------------------------------------------------
# Pinned by the reference's OWN reducer outputs (tests/golden/ref_op.txt, ref_ltfb.txt: one of the 413 file pairs under
# experiments/230111_twstft_2M5/{OP,LTFB}; tools/make_golden.py also checked the row grammar of all 78 413 rows)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _parse(path):
    import calendar
    text = open(path, "rb").read().decode()
    head, body = text.split("\r\n", 1)
    rows = []
    for ln in body.split("\n"):
        if ln:
            f = ln.split("\t")
            rows.append((calendar.timegm(time.strptime(f[0], "%Y %m %d %H %M %S")),) + tuple(float(v) for v in f[1:]))
    return head, body.split("\n"), rows


@pytest.mark.parametrize("name", ["ref_op.txt", "ref_ltfb.txt"])
def test_row_format_reproduces_the_reference_files_byte_for_byte(name):
    """The rows gofinal writes are compared with _fmt() above; _fmt() of the values parsed from the reference's own
    files gives those files back, so the format (header, \\r\\n after the header only, %.12f/%.3f/%.1f) is the reference's."""
    head, lines, rows = _parse(os.path.join(GOLD, name))
    assert head == "% y  m  d  h  m  s\tdelay\t\tdf1\tSNR1\tdf2\tSNR2\tdelayrem\t\tdf1rem\tSNR1rem"
    assert _fmt(rows) == [ln for ln in lines if ln]
    census = json.load(open(os.path.join(GOLD, "ref_txt_census.json")))
    assert census["files"] >= 400 and census["rows_not_matching_the_grammar"] == 0


def test_two_way_combination_of_the_reference_files(exe, tmp_path):
    """tw.m on a file pair the reference shipped: gofinal tw parses them as they are and agrees with the restatement."""
    out = subprocess.run([exe, "tw", os.path.join(GOLD, "ref_op.txt"), os.path.join(GOLD, "ref_ltfb.txt")],
                         capture_output=True, text=True, check=True).stdout
    op_r, lt_r = (np.loadtxt(os.path.join(GOLD, n), comments="%") for n in ("ref_op.txt", "ref_ltfb.txt"))
    ref = orr.two_way(op_r, lt_r)
    assert f"diff={ref[0]}" in out
    if len(ref) > 1:
        assert abs(float(out.split("std=")[1].split()[0]) - ref[1]) < 1e-5
        assert abs(float(out.split("mean=")[1].split()[0]) - ref[2]) < 1e-5
        assert int(out.split("\nn=")[1]) == ref[3]

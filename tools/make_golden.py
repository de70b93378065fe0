#!/usr/bin/env python3
"""Generate tests/golden/* from the reference tree (run in the authoring container only).

Reads /root/reference (read-only) and writes small fixtures:
  prn_digests.json   sha256 + head bytes of the 19 PRN code files shipped by the reference
  parabola_kat.npz   15 050 (|xvalm1|,|xval|,|xvalp1|,correction) triples + df KAT from
                     experiments/220616_Besancon/*.mat.gz
  ref_py_*.json      outputs of the reference's own numpy correlator
                     experiments/221219_twoway/processing/godual_ranging.py (function
                     `processing`, executed UNMODIFIED; only the hard-coded file call on its
                     last line is dropped and, for the small cases, the module constant fs is set)
                     on seeded synthetic records from amaranth_twstft_b200.synth.
  mat_schema.json    variable names / shapes / dtypes of reference result .mat files.
  ref_op.txt, ref_ltfb.txt, ref_txt_census.json   one pair of the reference's own *op.txt / *ltfb.txt reducer outputs
                     (experiments/230111_twstft_2M5) and a census of the row grammar over all 414 shipped files
The GPU box has no /root/reference: tests only read the fixtures.
"""
import glob
import gzip
import hashlib
import io
import json
import os
import re
import sys

import numpy as np
import scipy.io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from amaranth_twstft_b200 import synth  # noqa: E402
from oracle import prn  # noqa: E402


def read_maybe_gz(path):
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if path.endswith(".gz") else raw


PRN_FILES = [  # (path, bitlen, taps_a, taps_b, noiselen)  -- bitlen of the two *25000* files is really 15
    ("experiments/221207_twoway_codes/codes/noiselen100000_bitlen17_taps09.bin.gz", 17, 9, None, 100000),
    ("experiments/221207_twoway_codes/codes/noiselen100000_bitlen17_taps15.bin.gz", 17, 15, None, 100000),
    ("experiments/221207_twoway_codes/codes/noiselen10000_bitlen14_taps43.bin.gz", 14, 43, None, 10000),
    ("experiments/221207_twoway_codes/codes/noiselen10000_bitlen14_taps57.bin.gz", 14, 57, None, 10000),
    ("experiments/221207_twoway_codes/codes/noiselen2500000_bitlen22_taps03.bin.gz", 22, 3, None, 2500000),
    ("experiments/221207_twoway_codes/codes/noiselen2500000_bitlen22_taps57.bin.gz", 22, 57, None, 2500000),
    ("experiments/221207_twoway_codes/codes/noiselen250000_bitlen18_taps39.bin.gz", 18, 39, None, 250000),
    ("experiments/221207_twoway_codes/codes/noiselen250000_bitlen18_taps63.bin.gz", 18, 63, None, 250000),
    ("experiments/221207_twoway_codes/codes/noiselen25000_bitlen17_taps03.bin.gz", 15, 3, None, 25000),
    ("experiments/221207_twoway_codes/codes/noiselen25000_bitlen17_taps17.bin.gz", 15, 17, None, 25000),
    ("experiments/221207_twoway_codes/codes/noiselen500000_bitlen19_taps39.bin.gz", 19, 39, None, 500000),
    ("experiments/221207_twoway_codes/codes/noiselen500000_bitlen19_taps63.bin.gz", 19, 63, None, 500000),
    ("experiments/221207_twoway_codes/codes/noiselen5000_bitlen13_taps27.bin.gz", 13, 27, None, 5000),
    ("experiments/221207_twoway_codes/codes/noiselen5000_bitlen13_taps39.bin.gz", 13, 39, None, 5000),
    ("experiments/220706_TWSTFT/LTFB_prn57bpsk22bits.bin", 22, 57, None, None),
    ("experiments/220706_TWSTFT/OP_prn22bpskcode0.bin", 22, 3, None, None),
    ("experiments/231001_DLL_PLL/0.bin", 17, 9, None, None),
    ("experiments/231001_DLL_PLL/1.bin", 17, 15, None, None),
    ("Doc/EuGRD_tutorial/slides_jmf/prn10bpsk.bin", 10, 9, None, None),
]


def make_prn():
    out = []
    for path, bitlen, ta, tb, nl in PRN_FILES:
        data = read_maybe_gz(os.path.join(REF, path))
        nl = nl or len(data)
        mine = prn.prn_file_bytes(bitlen, nl, ta, tb).tobytes()
        assert mine == data, f"oracle PRN differs from reference file {path}"
        out.append(dict(file=path, bitlen=bitlen, taps_a=ta, taps_b=tb, noiselen=nl,
                        sha256=hashlib.sha256(data).hexdigest(), head=list(data[:64]),
                        ones=int(np.frombuffer(data, dtype=np.uint8).sum())))
        print("PRN ok", path, nl)
    json.dump(out, open(os.path.join(OUT, "prn_digests.json"), "w"), indent=1)


def make_parabola():
    am1, a0, ap1, corr, dfs = [], [], [], [], []
    for f in sorted(glob.glob(os.path.join(REF, "experiments/220616_Besancon/*.mat.gz"))):
        m = scipy.io.loadmat(io.BytesIO(read_maybe_gz(f)))
        for ch in ("1", "2"):
            if "xval" + ch not in m:
                continue
            am1.append(np.abs(m["xval" + ch + "m1"]).ravel())
            a0.append(np.abs(m["xval" + ch]).ravel())
            ap1.append(np.abs(m["xval" + ch + "p1"]).ravel())
            corr.append(m["correction" + ch].ravel())
        if "df" in m:
            dfs.append(np.asarray(m["df"], dtype=np.float64).ravel())
    am1, a0, ap1, corr = map(np.concatenate, (am1, a0, ap1, corr))
    dfs = np.concatenate(dfs)
    print("parabola triples", len(corr), "df samples", len(dfs))
    np.savez_compressed(os.path.join(OUT, "parabola_kat.npz"), am1=am1, a0=a0, ap1=ap1,
                        correction=corr, df=dfs)


def load_reference_processing():
    src = open(os.path.join(REF, "experiments/221219_twoway/processing/godual_ranging.py")).read()
    lines = src.rstrip("\n").split("\n")
    assert lines[-1].startswith("ranging("), lines[-1]
    ns = {}
    exec(compile("\n".join(lines[:-1]), "godual_ranging.py", "exec"), ns)
    return ns


def make_ref_py():
    ns = load_reference_processing()
    cases = [  # name, bitlen, taps, noiselen, fs, seeds
        ("n5e6", 22, 3, 2500000, 5e6, [20261017, 20261018]),
        ("n2e5", 17, 9, 100000, 2e5, [20261017 + i for i in range(4)]),
        ("n1e4", 13, 27, 5000, 1e4, [20261017 + i for i in range(6)]),
    ]
    for name, bitlen, taps, noiselen, fs, seeds in cases:
        bits = prn.lfsr_bits(bitlen, taps, noiselen)
        code = np.repeat(np.array(list(bits.tobytes())), 2) * 2 - 1      # as godual_ranging.py:74-76
        fcode = np.conj(np.fft.fft(code))
        n = len(code)
        ns["fs"] = fs
        freq = np.linspace(-fs / 2, fs / 2, num=n, dtype=float)          # :78
        frange = 8000 * fs / 5e6
        k = np.nonzero((freq < 2 * frange) & (freq > -2 * frange))[0]     # :79 with foffset=0
        temps = np.array(range(0, n)) / fs
        recs = []
        for seed in seeds:
            scale = fs / 5e6
            # carrier offset kept inside the reference python's +/-8 kHz (scaled) window
            ov = dict(df=float(np.random.default_rng(seed + 999).uniform(-7000, 7000) * scale))
            raw, (p1, p2) = synth.record_pair(code.astype(np.float64), seed, fs=fs, ch1=ov)
            outs = {}
            for ch, sl in ((1, (0, 1)), (2, (2, 3))):
                d = np.array(raw[sl[0]::4], dtype=complex)
                d.imag = np.array(raw[sl[1]::4])
                d -= np.mean(d)                                           # :100-103
                r = ns["processing"](d, k, freq, temps, fcode, code)
                names = ("indice", "correction", "SNRr", "SNRi", "df", "puissance",
                         "puissancecode", "puissancenoise")
                outs[str(ch)] = {nm: (int(v) if nm == "indice" else float(v)) for nm, v in zip(names, r)}
            recs.append(dict(seed=seed, ch1_override=ov, truth=[p1, p2], sha256=hashlib.sha256(raw.tobytes()).hexdigest(),
                             out=outs))
            print(name, seed, outs["1"]["indice"], outs["2"]["indice"])
        json.dump(dict(case=name, bitlen=bitlen, taps=taps, noiselen=noiselen, fs=fs,
                       window_hz=2 * frange, generator="amaranth_twstft_b200.synth.record_pair",
                       note="indice is 0-based (numpy argmax); SNR uses np.var (ddof=0); fine frequency ON",
                       records=recs),
                  open(os.path.join(OUT, f"ref_py_{name}.json"), "w"), indent=1)


def make_mat_schema():
    out = {}
    for f in sorted(glob.glob(os.path.join(REF, "experiments/240527/op/*.mat.gz")))[:3] + \
            sorted(glob.glob(os.path.join(REF, "experiments/240527/ltfb/*.mat.gz")))[:3]:
        m = scipy.io.loadmat(io.BytesIO(read_maybe_gz(f)))
        out[os.path.relpath(f, REF)] = {k: dict(shape=list(v.shape), dtype=str(v.dtype))
                                        for k, v in m.items() if not k.startswith("__")}
    json.dump(out, open(os.path.join(OUT, "mat_schema.json"), "w"), indent=1)


def make_text_fixtures():
    """One pair of the reference's OWN reducer outputs (gofinal_op.m / gofinal_ltfb.m rows, 414 such files ship under
    experiments/230111_twstft_2M5/{OP,LTFB}): pins the row format and feeds the `tw` combination."""
    for src, dst in (("experiments/230111_twstft_2M5/OP/1673442361.1673442552.59955op.txt.gz", "ref_op.txt"),
                     ("experiments/230111_twstft_2M5/LTFB/1673442372ltfb.txt.gz", "ref_ltfb.txt")):
        open(os.path.join(OUT, dst), "wb").write(read_maybe_gz(os.path.join(REF, src)))
    # every shipped file has the same header and row grammar: record that as a count
    rowre = re.compile(rb"^\d{4} \d\d \d\d \d\d \d\d \d\d(\t-?\d+\.\d{12}\t-?\d+\.\d{3}\t-?(\d+\.\d|Inf|NaN)){3}$")
    n_files = n_rows = bad = 0
    for path in sorted(glob.glob(os.path.join(REF, "experiments/230111_twstft_2M5/*/*[pb].txt.gz"))):
        lines = read_maybe_gz(path).split(b"\n")
        n_files += 1
        assert lines[0].rstrip(b"\r") == b"% y  m  d  h  m  s\tdelay\t\tdf1\tSNR1\tdf2\tSNR2\tdelayrem\t\tdf1rem\tSNR1rem", path
        for ln in lines[1:]:
            if ln:
                n_rows += 1
                bad += rowre.match(ln) is None
    json.dump({"files": n_files, "rows": n_rows, "rows_not_matching_the_grammar": bad},
              open(os.path.join(OUT, "ref_txt_census.json"), "w"), indent=1)
    print("text fixtures:", n_files, "files,", n_rows, "rows,", bad, "outside the row grammar")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1:] or ["prn", "parabola", "mat", "refpy", "text"]
    if "prn" in what:
        make_prn()
    if "parabola" in what:
        make_parabola()
    if "mat" in what:
        make_mat_schema()
    if "refpy" in what:
        make_ref_py()
    if "text" in what:
        make_text_fixtures()

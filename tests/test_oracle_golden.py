"""Pin the oracle: PRN files, parabola KAT, df-grid KAT and the reference's own numpy correlator."""
import hashlib
import json
import os

import numpy as np
import pytest

from amaranth_twstft_b200 import synth
from oracle import correlator as C
from oracle import prn

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("entry", json.load(open(os.path.join(G, "prn_digests.json"))),
                         ids=lambda e: os.path.basename(e["file"]))
def test_prn_files_byte_exact(entry):
    """common.py:59-73 output == every PRN file shipped by the reference (sha256)."""
    data = prn.prn_file_bytes(entry["bitlen"], entry["noiselen"], entry["taps_a"], entry["taps_b"])
    assert list(data[:64]) == entry["head"]
    assert int(data.sum()) == entry["ones"]
    assert hashlib.sha256(data.tobytes()).hexdigest() == entry["sha256"]


def test_prn_python_loop_matches_c():
    assert (prn.lfsr_bits_py(14, 43, 3000) == prn.lfsr_bits(14, 43, 3000)).all()
    assert (prn.lfsr_bits_py(22, 57, 3000, seed=5) == prn.lfsr_bits(22, 57, 3000, seed=5)).all()


def test_parabola_kat():
    """godual.m:41-50 stored correction == closed form of godual_ranging.m:33 on 15 050 triples."""
    z = np.load(os.path.join(G, "parabola_kat.npz"))
    c = C.parabola(z["am1"], z["a0"], z["ap1"])
    assert len(c) == 15050
    assert np.max(np.abs(c - z["correction"])) < 5e-15


def test_df_grid_kat():
    """df = linspace(-fs/2,fs/2,N)(idx)/2 : the 843.75016875 Hz quirk stored in the Besancon files."""
    z = np.load(os.path.join(G, "parabola_kat.npz"))
    freq = C.freq_axis_G(5e6, 5000000)
    grid = freq / 2
    df = z["df"]
    # every stored df must sit exactly on the quirky grid
    idx = np.rint((df * 2 + 2.5e6) * (5000000 - 1) / 5e6).astype(np.int64)
    assert np.max(np.abs(grid[idx] - df)) < 1e-9
    assert np.any(np.abs(df - 843.7501687498) < 1e-9)
    assert abs(grid[2501688 - 1] - 843.7501687498) < 1e-9      # Octave index 2501688


def _run_case(name):
    g = json.load(open(os.path.join(G, f"ref_py_{name}.json")))
    fs = g["fs"]
    bits = prn.lfsr_bits(g["bitlen"], g["taps"], g["noiselen"])
    code = C.code_bpsk(bits)
    n = len(code)
    freq = C.freq_axis_G(fs, n)
    k = C.window(freq, -g["window_hz"], g["window_hz"])
    fcode = C.fcode_G(code)
    for rec in g["records"]:
        raw, _ = synth.record_pair(code, rec["seed"], fs=fs, ch1=rec["ch1_override"])
        assert hashlib.sha256(raw.tobytes()).hexdigest() == rec["sha256"], "synthetic generator drifted"
        res = C.record_G(raw, k, freq, fcode, code, fs=fs, fine_freq=True, ddof=0)
        for ch in (1, 2):
            ref, o = rec["out"][str(ch)], res[ch]
            assert o["p0"] == ref["indice"]
            assert abs(o["correction"] - ref["correction"]) < 1e-9
            assert abs(o["df"] - ref["df"]) < 1e-9 * max(1.0, abs(ref["df"]))
            for nm in ("SNRr", "SNRi", "puissance", "puissancecode", "puissancenoise"):
                assert abs(o[nm] - ref[nm]) <= 1e-9 * abs(ref[nm]) + 1e-300, nm


def test_reference_python_small():
    _run_case("n1e4")


def test_reference_python_200k():
    _run_case("n2e5")


@pytest.mark.slow
def test_reference_python_full_size():
    """N=5e6 two-channel records, the shape BASELINE.json's metric is quoted on."""
    _run_case("n5e6")

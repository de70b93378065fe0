"""Shared parity checks: the library under test (CUDA product build on the GPU box, or the test-only
host emulation build on CPU) against the oracle, with the tolerances of BASELINE.json's north_star:
peak index bit-exact, fine delay 1e-4 sample, phase 1e-4 rad, SNR 0.01 dB."""
import os

import numpy as np

from amaranth_twstft_b200 import synth
from amaranth_twstft_b200.api import Correlator
from oracle import correlator as oc
from oracle import prn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libtwstft_emu.so")
CUDA_LIB = os.path.join(ROOT, "amaranth_twstft_b200", "libtwstft.so")

TOL_DELAY_SAMPLES = 1e-4
TOL_PHASE_RAD = 1e-4
TOL_SNR_DB = 0.01

CODES = {  # window length -> (bitlen, taps, noiselen): all from experiments/221207_twoway_codes/codes
    10000: (13, 27, 5000),
    200000: (17, 9, 100000),
    250000: (17, 9, 125000),
    1000000: (19, 39, 500000),
    5000000: (22, 3, 2500000),
}


def snr_db(r):
    return 10 * np.log10(r["SNRr"] + r["SNRi"])


def assert_row_matches(r, o, where="", delay_units=3.0):
    """r: one row of the library result, o: oracle dict (processing_G / processing_K)."""
    assert int(r["indice"]) == o["indice"], f"{where}: peak index {int(r['indice'])} != {o['indice']}"
    assert abs(r["correction"] - o["correction"]) / delay_units < TOL_DELAY_SAMPLES, \
        f"{where}: correction {r['correction']} vs {o['correction']}"
    if "s" in o:
        assert int(r["carrier_bin"]) == o["s"], f"{where}: carrier bin {r['carrier_bin']} vs {o['s']}"
    assert abs(r["df"] - o["df"]) <= 1e-9 * max(1.0, abs(o["df"])), f"{where}: df {r['df']} vs {o['df']}"
    xv = complex(r["xval_re"], r["xval_im"])
    assert abs(np.angle(xv * np.conj(o["xval"]))) < TOL_PHASE_RAD, f"{where}: phase"
    assert abs(abs(xv) / abs(o["xval"]) - 1) < 1e-4, f"{where}: |xval|"
    for nm in ("xvalm1", "xvalp1"):
        v = complex(r[nm + "_re"], r[nm + "_im"])
        assert abs(v - o[nm]) < 1e-4 * abs(o["xval"]), f"{where}: {nm}"
    assert abs(snr_db(r) - 10 * np.log10(o["SNRr"] + o["SNRi"])) < TOL_SNR_DB, f"{where}: SNR"
    tot = o["SNRr"] + o["SNRi"]
    assert abs(r["SNRr"] - o["SNRr"]) < 2e-3 * tot and abs(r["SNRi"] - o["SNRi"]) < 2e-3 * tot, f"{where}: SNRr/i"
    for nm in ("puissance", "puissancecode", "puissancenoise"):
        assert abs(r[nm] / o[nm] - 1) < 1e-3, f"{where}: {nm} {r[nm]} vs {o[nm]}"


def make_code(n):
    bitlen, taps, nl = CODES[n]
    bits = prn.lfsr_bits(bitlen, taps, nl)
    return bits, oc.code_bpsk(bits)


def run_variant_g(lib, n, seeds, fs=5e6, window=(-20000.0, 20000.0), max_batch=2, overrides=None):
    """Two-channel interleaved records through twstft_process vs godual_ranging.m semantics."""
    bits, code = make_code(n)
    recs = [synth.record_pair(code, s, fs=fs, **(overrides or {}))[0] for s in seeds]
    with Correlator(nsamples=n, fs=fs, window_hz=window, max_batch=max_batch, lib=lib) as cor:
        cor.set_code(bits)
        res = cor.process(np.stack(recs))
    freq = oc.freq_axis_G(fs, n)
    k = oc.window(freq, window[0], window[1])
    fcode = oc.fcode_G(code)
    for i, raw in enumerate(recs):
        ref = oc.record_G(raw, k, freq, fcode, code, fs=fs)
        for ch in (1, 2):
            assert_row_matches(res[i, ch - 1], ref[ch], where=f"N={n} seed={seeds[i]} ch{ch}")
    return res

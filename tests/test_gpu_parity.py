"""GPU parity tests proper: the CUDA library through its C ABI against the oracle."""
import json
import os

import numpy as np
import pytest

import parity_common as pc
from amaranth_twstft_b200 import synth
from amaranth_twstft_b200.api import Correlator
from oracle import correlator as oc
from oracle import prn

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_backend_is_cuda():
    with Correlator(nsamples=10000, max_batch=1, lib=pc.CUDA_LIB) as cor:
        assert cor.backend == "cuda"


@pytest.mark.parametrize("entry", json.load(open(os.path.join(G, "prn_digests.json"))),
                         ids=lambda e: os.path.basename(e["file"]))
def test_gpu_lfsr_matches_reference_files(entry):
    import ctypes
    import hashlib
    from amaranth_twstft_b200 import _lib
    lib = _lib.load(pc.CUDA_LIB)
    n = entry["noiselen"]
    out = np.zeros(n * (2 if entry["taps_b"] else 1), dtype=np.uint8)
    rc = lib.twstft_prn_generate(0, entry["bitlen"], entry["taps_a"], entry["taps_b"] or 0, 1, n, out.ctypes.data)
    assert rc == 0
    assert hashlib.sha256(out.tobytes()).hexdigest() == entry["sha256"]


@pytest.mark.parametrize("n", [10000, 200000, 250000, 1000000])
def test_variant_g_small(n):
    pc.run_variant_g(pc.CUDA_LIB, n, seeds=[11, 12, 13])


@pytest.mark.parametrize("n,fast", [(20000, False), (50000, False), (500000, False), (500000, True), (50000, True)])
def test_other_shipped_code_lengths(n, fast):
    """N = 2 x {10 000, 25 000, 250 000} chips with the matching shipped LFSR codes
    (experiments/221207_twoway_codes/codes/noiselen{10000,25000,250000}_*): 800-point rows (radices 20.2.20)."""
    pc.run_variant_g(pc.CUDA_LIB, n, seeds=[21, 22], fast=fast)


def test_variant_g_full_size():
    """configs[1]: two-channel 1-s records at 5 MS/s, 2.5 Mchip/s 22-bit code."""
    pc.run_variant_g(pc.CUDA_LIB, 5000000, seeds=[20261017, 20261018], max_batch=2)


@pytest.mark.parametrize("name", ["n1e4", "n2e5", "n5e6"])
def test_against_reference_python_outputs(name):
    """fine_freq=1, ddof=0: CUDA results against the committed outputs of the reference's own
    experiments/221219_twoway/processing/godual_ranging.py on the same seeded records (n5e6 = configs[1] shape)."""
    pc.run_ref_py_golden(pc.CUDA_LIB, name)


@pytest.mark.parametrize("n", [200000, 250000, 1000000, 5000000])
def test_fast_peak_path_matches_the_oracle(n):
    """One inverse transform + exact local lags + certificate (k_fast.cuh): same rows as the oracle, all certified."""
    pc.run_variant_g(pc.CUDA_LIB, n, seeds=(11, 12) if n >= 1000000 else (11, 12, 13), fast=True)


def test_fast_peak_path_with_fine_frequency_and_qpsk():
    """Fast peak path behind the fine-frequency front end (golden outputs of the reference's own python at N = 5e6)
    and with the complex QPSK replica / x^4 estimator."""
    pc.run_ref_py_golden(pc.CUDA_LIB, "n5e6", fast=True)
    pc.run_qpsk(pc.CUDA_LIB, fast=True)


@pytest.mark.parametrize("n,delay", [(200000, 1234.56), (200000, 0.4), (5000000, 2500000.3)])
def test_fast_peak_path_variant_k_windows(n, delay):
    pc.run_variant_k(pc.CUDA_LIB, n, delay, fast=True)


def test_fast_peak_path_degenerate_records():
    pc.run_fast_peak_degenerate(pc.CUDA_LIB, 250000)


def test_fast_peak_path_falls_back_without_a_peak():
    pc.run_fast_peak_fallback(pc.CUDA_LIB, 250000)


def test_full_size_circular_shift_property():
    """Size-independent property at the BASELINE size (N = 5e6).  With a derotation frequency of a whole number of
    cycles per window the whole chain is circularly shift-equivariant: rotating the record by d samples moves the
    peak by 3*d on the interpolated grid (mod 3N) and leaves fine delay, |xval| and SNR unchanged; the estimator's
    magnitude spectrum, hence its arg-max, does not move (fp32 transforms: small tolerances)."""
    n = 5000000
    bits, code = pc.make_code(n)
    raw, _ = synth.record_pair(code, 4242, ch1=dict(df=4321.0), ch2=dict(df=0.0))
    d = 1234567
    rolled = np.roll(raw.reshape(-1, 4), d, axis=0).reshape(-1)
    both = np.concatenate([raw, rolled])
    with Correlator(nsamples=n, max_batch=2, lib=pc.CUDA_LIB) as cor:
        cor.set_code(bits)
        df, bins = cor.estimate_carrier(both)
        res = cor.process_windows(both, [0, n], [[4321.0, 0.0]] * 2)
    assert np.array_equal(bins[0], bins[1]) and np.array_equal(df[0], df[1])
    for ch in (0, 1):
        a, b = res[0, ch], res[1, ch]
        assert (int(a["indice"]) - 1 + 3 * d) % (3 * n) == int(b["indice"]) - 1
        assert abs(a["correction"] - b["correction"]) < 3e-4
        assert abs(pc.snr_db(a) - pc.snr_db(b)) < pc.TOL_SNR_DB
        assert abs(abs(complex(a["xval_re"], a["xval_im"])) / abs(complex(b["xval_re"], b["xval_im"])) - 1) < 1e-4


@pytest.mark.parametrize("n", [200000, 5000000])
def test_cpp_program_flavour(n):
    """processing/CPP/main.cpp:224-361 semantics (Hamming-windowed code spectrum, given df, var/n)."""
    pc.run_variant_cpp(pc.CUDA_LIB, n, seeds=(5, 6) if n < 5000000 else (5,))


@pytest.mark.parametrize("m,s,foffset,remote", [(10000, 7, 0.0, 0), (200000, 6, 0.0, 0), (200000, 3, 2500.0, 1)])
def test_cpp_whole_file_estimator(m, s, foffset, remote):
    """GoRanging::df (processing/CPP/main.cpp:363-450): L = s*m decimated points, s sub-transforms on the plan."""
    pc.run_decimated_estimator(pc.CUDA_LIB, m=m, s=s, foffset=foffset, remote=remote)


def test_cpp_whole_file_estimator_raw_mean_with_offset():
    """Large DC term and foffset != 0: the raw mean is subtracted after the derotation (main.cpp:381-420)."""
    pc.run_decimated_estimator(pc.CUDA_LIB, m=200000, s=3, foffset=2500.0, remote=0, dc=900.0 - 650.0j, check_carrier=False)


def test_qpsk_x4_estimator_and_complex_replica():
    """configs[2]: QPSK, x^4 estimator, 100 kchip/s code (N = 200 000)."""
    pc.run_qpsk(pc.CUDA_LIB)


@pytest.mark.parametrize("n,delay", [(10000, 0.0), (10000, 1.0 / 3.0), (10000, 9999.6), (200000, 4321.7), (200000, 199999.9)])
def test_variant_k_windows(n, delay):
    """K variant (fcode=fft(code), P=fcode.*conj(ffty), wrap quirks) through twstft_process_windows."""
    pc.run_variant_k(pc.CUDA_LIB, n, delay)


def test_remote_window_and_negative_offset():
    pc.run_variant_g(pc.CUDA_LIB, 250000, seeds=[5], window=(80000.0, 120000.0),
                     overrides=dict(ch1=dict(df=51234.5), ch2=dict(df=49000.25)))
    pc.run_variant_g(pc.CUDA_LIB, 250000, seeds=[6], window=(-120000.0, -80000.0),
                     overrides=dict(ch1=dict(df=-51234.5), ch2=dict(df=-43000.0)))


def test_host_and_device_entry_points_agree():
    """twstft_process (host records, staged copies) == twstft_process_device (resident records), 5 records
    through max_batch 4 (chunked staging + a ragged last batch)."""
    n = 250000
    bits, code = pc.make_code(n)
    recs = np.stack([synth.record_pair(code, s)[0] for s in range(50, 55)])
    with Correlator(nsamples=n, max_batch=4, lib=pc.CUDA_LIB) as cor:
        cor.set_code(bits)
        a = cor.process(recs)
        dev = cor.to_device(recs)
        b = cor.process_device(dev, len(recs))
        cor.device_free(dev)
    for nm in a.dtype.names:
        assert np.array_equal(a[nm], b[nm]), nm


def test_doppler_bank_small():
    pc.run_doppler_bank(pc.CUDA_LIB)


def test_doppler_bank_64_bins_full_size():
    """configs[3]: 1-s record at -10 dB correlated over 64 Doppler bins in +/-2 kHz."""
    res, dfs = pc.run_doppler_bank(pc.CUDA_LIB, n=5000000, nbins=64, span=2000.0, true_df=437.0, seed=20261017,
                                   check_every=7)
    assert abs(dfs[int(np.argmax(res["metric"][:, 0]))] - 437.0) <= (dfs[1] - dfs[0])


def test_search_df_finds_the_carrier():
    pc.run_search_df(pc.CUDA_LIB)


def test_block2_estimator_production_size():
    """2-s blocks at 5 MS/s: 10^7-point estimator spectrum through the fused radix-2 split, df bit-exact."""
    pc.run_block2_estimator(pc.CUDA_LIB, n=5000000, fs=5e6, seeds=(3,))


def test_block2_estimator_small():
    pc.run_block2_estimator(pc.CUDA_LIB)


def test_k_tracking_loop_small():
    pc.run_ktrack(pc.CUDA_LIB)


def test_k_tracking_loop_production_shape():
    """claudio_aligned_code_lo_separate.m configuration: fs = 5 MS/s, 100 kchip/s code (N = 200 000, 25 windows
    per second), 2-s blocks; 2.2 blocks of a continuous recording with the code 61 234.5 samples late."""
    got = pc.run_ktrack(pc.CUDA_LIB, fs=5e6, n=200000, nblocks=2.2, delay=61234.5, df=-3210.25, seed=20261017,
                        max_batch=32)
    assert len(got["df"]) == 2 and len(got["indice1"]) >= 98


def test_k_re_tracking_loop_small():
    pc.run_ktrack_re(pc.CUDA_LIB)


def test_k_re_tracking_loop_production_shape():
    """claudio_aligned_code_re_separate.m configuration: fs = 5 MS/s, 100 kchip/s code (N = 200 000), 2-s blocks,
    remote window 92..108 kHz: search_df over the 10^7-point block spectrum, seven-bin estimator over
    10^7 + len(dold) samples, 2.2 blocks with the code 61 234.5 samples late."""
    got = pc.run_ktrack_re(pc.CUDA_LIB, fs=5e6, n=200000, nblocks=2.2, delay=61234.5, df=50321.5, seed=20261018,
                           max_batch=32, skip_s=0.1)
    assert len(got["df"]) == 2 and len(got["indice1"]) >= 98


def test_estimate_bins_any_length():
    pc.run_estimate_bins(pc.CUDA_LIB)
    pc.run_estimate_bins(pc.CUDA_LIB, fs=5e6, extra=123457)


def test_search_df_on_two_second_block():
    pc.run_search_df_block(pc.CUDA_LIB)
    pc.run_search_df_block(pc.CUDA_LIB, fs=5e6, n=200000)


@pytest.mark.parametrize("n", [200000, 5000000])
def test_octave_function_seam_processing(n):
    """processing(d,k) / processing(d,k,df) with the complex double column the Octave functions receive
    (godual_ranging.m:12, claudio_aligned_code_lo_separate.m:24) through twstft_processing_cf64{,_df}."""
    pc.run_octave_seam(pc.CUDA_LIB, n, seeds=(51,) if n > 1000000 else (51, 52))


def test_octave_function_seam_search_df():
    pc.run_octave_seam_search_df(pc.CUDA_LIB)
    pc.run_octave_seam_search_df(pc.CUDA_LIB, fs=5e6, n=200000, df=50321.5, window=(92000.0, 108000.0))


def test_fast_peak_governor_and_batched_fallback():
    pc.run_fast_peak_governor(pc.CUDA_LIB, 250000)


@pytest.mark.parametrize("kind,n", [("G", 5000000), ("K", 200000)])
def test_snr_sweep_minus20_to_minus5_db(kind, n):
    """Field range of input SNR (experiments/221127_SNR/README.md:56-62), 16 seeds per cell, fast peak path on:
    peak index and carrier bin exact, fine delay / phase / SNR inside the north-star tolerances in every cell."""
    stats = pc.run_snr_sweep(pc.CUDA_LIB, n, kind=kind)
    assert len(stats) == 4 and all(c["fast_seen"] + 0 >= 0 for c in stats)


def test_step_halving_acquisition():
    """experiments/231001_DLL_PLL/rxcomplex.cpp:522-568 on top of the bank kernels, N = 2e5 and the 1-s window."""
    pc.run_acquire(pc.CUDA_LIB, 200000)
    pc.run_acquire(pc.CUDA_LIB, 5000000, true_df=-912.6, fstep=1000.0)


def test_variant_g_peak_at_array_edge_sets_status():
    pc.run_edge_peak(pc.CUDA_LIB, 250000)


@pytest.mark.parametrize("n", [250000, 5000000])
def test_records_without_a_peak_match_the_oracle(n):
    pc.run_no_peak_vs_oracle(pc.CUDA_LIB, n)

"""CPU checks of the CUDA kernels' arithmetic: the kernel phase functions compiled for the host
(tests/emu, test-only) are run through the same C ABI and compared with the oracle.  These are the
'host logic' tests; the parity tests proper run on the GPU (test_gpu_parity.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import parity_common as pc
from amaranth_twstft_b200 import synth
from amaranth_twstft_b200.api import Correlator
from oracle import correlator as oc
from oracle import prn

G = os.path.join(os.path.dirname(__file__), "golden")
needs_emu = pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
pytestmark = needs_emu


@pytest.mark.parametrize("n", [10000, 200000, 250000])
def test_variant_g(n):
    pc.run_variant_g(pc.EMU_LIB, n, seeds=[11, 12, 13])


@pytest.mark.parametrize("n,fast", [(20000, False), (50000, False), (50000, True), (20000, True)])
def test_other_shipped_code_lengths(n, fast):
    """N = 2 x {10 000, 25 000} chips (experiments/221207_twoway_codes/codes): the 800-point row transform
    (radices 20.2.20) and the 5.25 column transform with the 400-point rows."""
    pc.run_variant_g(pc.EMU_LIB, n, seeds=[21, 22], fast=fast)


@pytest.mark.parametrize("name", ["n1e4", "n2e5"])
def test_fine_frequency_against_reference_python_outputs(name):
    """A10 + the whole G path against tests/golden/ref_py_*.json (outputs of the reference's own numpy code)."""
    pc.run_ref_py_golden(pc.EMU_LIB, name)


@pytest.mark.parametrize("n", [10000, 200000])
def test_cpp_program_flavour(n):
    pc.run_variant_cpp(pc.EMU_LIB, n)


@pytest.mark.parametrize("s,foffset,remote", [(3, 0.0, 0), (8, 0.0, 0), (5, 1500.0, 1)])
def test_cpp_whole_file_estimator(s, foffset, remote):
    pc.run_decimated_estimator(pc.EMU_LIB, m=10000, s=s, foffset=foffset, remote=remote)


def test_cpp_whole_file_estimator_raw_mean_with_offset():
    """main.cpp:381-420 subtracts the mean of the RAW samples from the derotated ones: with a large DC term and
    foffset != 0 that leaves a strong tone at -foffset which the estimator must see exactly as the reference does."""
    pc.run_decimated_estimator(pc.EMU_LIB, m=10000, s=4, foffset=1500.0, remote=0, dc=900.0 - 650.0j, check_carrier=False)


@pytest.mark.parametrize("n", [10000, 200000, 250000])
def test_fast_peak_path_matches_the_oracle(n):
    pc.run_variant_g(pc.EMU_LIB, n, seeds=(11, 12, 13), fast=True)


def test_fast_peak_path_with_fine_frequency_and_qpsk():
    """The fast peak path behind the other front ends: fine-frequency refinement (golden outputs of the reference's
    python) and the complex QPSK replica with the x^4 estimator."""
    pc.run_ref_py_golden(pc.EMU_LIB, "n2e5", fast=True)
    pc.run_qpsk(pc.EMU_LIB, fast=True)


@pytest.mark.parametrize("delay", [1234.56, 3.2, 199998.7])
def test_fast_peak_path_variant_k_windows(delay):
    """processing(d,k,df) of the K scripts through the fast peak path, including peaks at the array ends (fallback)."""
    pc.run_variant_k(pc.EMU_LIB, 200000, delay, fast=True)


def test_fast_peak_path_degenerate_records():
    pc.run_fast_peak_degenerate(pc.EMU_LIB)


def test_fast_peak_path_falls_back_without_a_peak():
    pc.run_fast_peak_fallback(pc.EMU_LIB, 10000)


def test_variant_g_1e6():
    pc.run_variant_g(pc.EMU_LIB, 1000000, seeds=[11])


@pytest.mark.slow
def test_variant_g_full_size_record():
    pc.run_variant_g(pc.EMU_LIB, 5000000, seeds=[20261017], max_batch=1)


@pytest.mark.parametrize("entry", [e for e in json.load(open(os.path.join(G, "prn_digests.json")))
                                   if e["noiselen"] <= 500000], ids=lambda e: os.path.basename(e["file"]))
def test_lfsr_kernel_matches_reference_files(entry):
    from amaranth_twstft_b200 import _lib
    lib = _lib.load(pc.EMU_LIB)
    n = entry["noiselen"]
    out = np.zeros(n, dtype=np.uint8)
    assert lib.twstft_prn_generate(0, entry["bitlen"], entry["taps_a"], 0, 1, n, out.ctypes.data) == 0
    assert hashlib.sha256(out.tobytes()).hexdigest() == entry["sha256"]


def test_lfsr_kernel_qpsk_interleave_and_seed():
    from amaranth_twstft_b200 import _lib
    lib = _lib.load(pc.EMU_LIB)
    out = np.zeros(2 * 1000, dtype=np.uint8)
    assert lib.twstft_prn_generate(0, 17, 9, 15, 5, 1000, out.ctypes.data) == 0
    assert (out == prn.prn_file_bytes(17, 1000, 9, 15, seed=5)).all()


@pytest.mark.parametrize("axis,lo,hi,n", [("G", -20000, 20000, 5000000), ("G", 80000, 120000, 5000000),
                                          ("G", -120000, -80000, 5000000), ("K", 100000, 120000, 200000),
                                          ("K", -8000, 8000, 200000), ("G", -20000, 20000, 10000)])
def test_window_bins_equal_octave_find(axis, lo, hi, n):
    fs = 5e6
    freq = oc.freq_axis_G(fs, n) if axis == "G" else np.linspace(-fs / 2, fs / 2 - 1.0, n)
    k = oc.window(freq, lo, hi)
    with Correlator(nsamples=n, fs=fs, window_hz=(lo, hi), freq_axis=axis, max_batch=1, lib=pc.EMU_LIB) as cor:
        assert cor.window_bins() == (k[0], k[-1])


def test_remote_window_and_negative_offset():
    """remote channel: carrier in the 80-120 kHz window of the x^2 spectrum (godual_ranging.m:82-90)."""
    pc.run_variant_g(pc.EMU_LIB, 250000, seeds=[5], window=(80000.0, 120000.0),
                     overrides=dict(ch1=dict(df=51234.5), ch2=dict(df=49000.25)))
    pc.run_variant_g(pc.EMU_LIB, 250000, seeds=[6], window=(-120000.0, -80000.0),
                     overrides=dict(ch1=dict(df=-51234.5), ch2=dict(df=-43000.0)))


def test_batching_is_transparent():
    """7 records through max_batch=3 == the same records one by one."""
    bits, code = pc.make_code(10000)
    recs = np.stack([synth.record_pair(code, s)[0] for s in range(30, 37)])
    with Correlator(nsamples=10000, max_batch=3, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        a = cor.process(recs)
    with Correlator(nsamples=10000, max_batch=1, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        b = np.concatenate([cor.process(r) for r in recs])
    for nm in a.dtype.names:
        assert np.array_equal(a[nm], b[nm]), nm


@pytest.mark.parametrize("delay", [0.0, 0.2, 1.0 / 3.0, 17.4, 4321.7, 9999.6])
def test_variant_k_windows(delay):
    """processing(d,k,df) of claudio_aligned_code_lo_separate.m:24-77 incl. the neighbour-wrap quirks."""
    pc.run_variant_k(pc.EMU_LIB, 10000, delay)


def test_variant_k_window_offsets_in_long_buffer():
    """Windows at arbitrary sample offsets of a longer single-channel buffer (dindex tracking, A11)."""
    n = 10000
    bits, code = pc.make_code(n)
    parts = [synth.record_single(code, 90 + i, delay=21.3, df=-777.0)[0] for i in range(3)]
    raw = np.concatenate(parts)
    offs = [0, 5, n - 3, 2 * n]
    with Correlator(nsamples=n, variant="K", layout="single", max_batch=3, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        res = cor.process_windows(raw, offs, [-777.0] * len(offs))
    d = oc.unpack_single(raw)
    for i, o in enumerate(offs):
        w = d[o:o + n]
        ref = oc.processing_K(w - w.mean(), -777.0, oc.fcode_K(code), code)
        pc.assert_row_matches(res[i, 0], ref, where=f"K offset={o}")


def test_variant_g_peak_at_array_edge_sets_status():
    pc.run_edge_peak(pc.EMU_LIB, 10000)


def test_records_without_a_peak_match_the_oracle():
    pc.run_no_peak_vs_oracle(pc.EMU_LIB, 10000)


def test_estimate_carrier_only():
    n = 250000
    bits, code = pc.make_code(n)
    recs = np.stack([synth.record_pair(code, s)[0] for s in (1, 2)])
    with Correlator(nsamples=n, max_batch=2, lib=pc.EMU_LIB) as cor:
        df, bins = cor.estimate_carrier(recs)
    freq = oc.freq_axis_G(5e6, n)
    k = oc.window(freq, -20000, 20000)
    for i in range(2):
        d1, d2 = oc.unpack_interleaved2(recs[i])
        for ch, d in enumerate((d1, d2)):
            s, f = oc.carrier_estimate(d - d.mean(), k, freq)
            assert bins[i, ch] == s and df[i, ch] == f


def test_qpsk_x4_estimator_and_complex_replica():
    pc.run_qpsk(pc.EMU_LIB)


def test_doppler_bank_matches_search_df_body():
    res, dfs = pc.run_doppler_bank(pc.EMU_LIB)
    assert abs(dfs[int(np.argmax(res["metric"][:, 0]))] - 437.0) <= (dfs[1] - dfs[0])


def test_search_df_finds_the_carrier():
    pc.run_search_df(pc.EMU_LIB)


def test_block2_estimator_equals_2n_point_fft():
    pc.run_block2_estimator(pc.EMU_LIB)
    pc.run_block2_estimator(pc.EMU_LIB, window=(100000.0, 120000.0), fs=250000.0 * 2)   # remote window, K-lo OP=0


def test_k_tracking_loop_moved_and_dold_carry():
    got = pc.run_ktrack(pc.EMU_LIB)
    assert len(got["df"]) == 2 and len(got["indice1"]) >= 98      # 2 blocks of 50 windows, one lost to alignment


def test_k_result_file_has_the_reference_schema(tmp_path):
    """save_mat writes the variables of lo_separate.m:166 with the shapes found in the reference's own result files
    (experiments/240527/op/localclaudio*.mat.gz, fixture tests/golden/mat_schema.json)."""
    import scipy.io
    from amaranth_twstft_b200.ktrack import save_mat
    got = pc.run_ktrack(pc.EMU_LIB, nblocks=1.1)
    path = str(tmp_path / "localclaudio_test.mat")
    save_mat(path, got)
    m = scipy.io.loadmat(path)
    schema = json.load(open(os.path.join(G, "mat_schema.json")))
    ref = next(v for k, v in schema.items() if "localclaudio" in k)
    for name, meta in ref.items():
        assert name in m, name
        assert m[name].ndim == 2 and m[name].shape[0] == 1, (name, m[name].shape)
        assert str(m[name].dtype) == meta["dtype"], (name, m[name].dtype, meta["dtype"])
    assert m["indice1"].shape == m["correction1"].shape == m["xval1"].shape == m["SNR1r"].shape
    assert m["code"].shape == (1, 10000)


def test_k_re_tracking_loop_search_and_fractional_index():
    """claudio_aligned_code_re_separate.m: search_df on the block at skip_s, seven-bin estimator on [dold ; block],
    un-floored indice1 -> fractional dindex and round() slices."""
    got = pc.run_ktrack_re(pc.EMU_LIB)
    assert len(got["df"]) == 2 and got["indice1"][1] != np.floor(got["indice1"][1])


def test_k_ranging_tracking_loop():
    pc.run_ktrack_re(pc.EMU_LIB, mode="ranging", df=1234.5, delay=777.7, skip_s=0.0)


def test_k_re_returns_none_without_carrier():
    from amaranth_twstft_b200.ktrack import KTracker
    from oracle import ktrack as okt
    bits, code = pc.make_code(10000)
    raw = np.random.default_rng(3).normal(0, 3000, int(2.1 * 2 * 250000) * 2).round().astype(np.int16)
    with KTracker(bits, fs=250000.0, ls=2, max_batch=8, lib=pc.EMU_LIB, mode="re", skip_s=0.5) as kt:
        assert kt.process_file(raw) is None
    assert okt.re_separate(raw, code, fs=250000.0, skip_s=0.5) is None


def test_estimate_bins_any_length():
    pc.run_estimate_bins(pc.EMU_LIB)
    pc.run_estimate_bins(pc.EMU_LIB, power=4, extra=333)


def test_search_df_on_two_second_block():
    pc.run_search_df_block(pc.EMU_LIB)
    pc.run_search_df_block(pc.EMU_LIB, df=-3456.25, window=(-8000.0, 8000.0))      # ranging window


@pytest.mark.parametrize("n", [10000, 250000])
def test_octave_function_seam_processing(n):
    pc.run_octave_seam(pc.EMU_LIB, n)


def test_octave_function_seam_search_df():
    pc.run_octave_seam_search_df(pc.EMU_LIB)


def test_fast_peak_governor_and_batched_fallback():
    pc.run_fast_peak_governor(pc.EMU_LIB, 10000)


@pytest.mark.parametrize("kind", ["G", "K"])
def test_snr_sweep_small(kind):
    """The sweep harness of the GPU suite at N = 1e4 (40 dB of processing gain: -10 dB and up only)."""
    pc.run_snr_sweep(pc.EMU_LIB, 10000, kind=kind, snrs=(-10.0, -3.0), seeds_per_cell=3, procs=2)


def test_step_halving_acquisition():
    pc.run_acquire(pc.EMU_LIB, 10000)

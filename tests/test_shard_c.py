"""The C host sharder (twstft_shard_*: one host thread + plan per GPU), the cycled day-of-data driver and the
device-side synthetic generator, on the test-only host emulation (two "GPUs" = two plans of the emulator)."""
import os

import numpy as np
import pytest

import parity_common as pc
from amaranth_twstft_b200 import synth
from amaranth_twstft_b200.api import Correlator, Sharder
from oracle import correlator as oc

pytestmark = pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
N = 10000


def _records(code, seeds):
    return np.concatenate([synth.record_pair(code, s)[0] for s in seeds])


def test_sharded_rows_equal_single_plan_rows_at_their_record_index():
    bits, code = pc.make_code(N)
    raw = _records(code, range(100, 107))                     # 7 records over 2 and 3 plans: ragged ranges
    with Correlator(nsamples=N, max_batch=2, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        one = cor.process(raw)
    for g in (2, 3):
        with Sharder(g, devices=[0] * g, nsamples=N, max_batch=2, lib=pc.EMU_LIB) as sh:
            sh.set_code(bits)
            assert np.array_equal(sh.process(raw), one)
    with Sharder(3, devices=[0] * 3, nsamples=N, max_batch=2, lib=pc.EMU_LIB) as sh:   # fewer records than GPUs
        sh.set_code(bits)
        assert np.array_equal(sh.process(raw[:4 * N * 2]), one[:2])


def test_cycled_run_draws_record_r_from_pool_slot_r_mod_pool():
    bits, code = pc.make_code(N)
    pool = _records(code, range(200, 205))                    # pool of 5, batch 2: chunks split at the pool end
    with Correlator(nsamples=N, max_batch=2, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        rows = cor.process(pool)
        cyc = cor.process_cycled(pool, 5, first_record=3, nrecords=9)
        assert np.array_equal(cyc, rows[[(3 + i) % 5 for i in range(9)]])
    with Sharder(2, devices=[0, 0], nsamples=N, max_batch=2, lib=pc.EMU_LIB) as sh:
        sh.set_code(bits)
        for resident in (True, False):
            out, wall, dev = sh.process_cycled(pool, 13, resident=resident)
            assert np.array_equal(out, rows[[i % 5 for i in range(13)]])
            assert wall >= 0 and len(dev) == 2


def test_device_synthetic_records_carry_the_drawn_parameters():
    bits, code = pc.make_code(N)
    with Correlator(nsamples=N, max_batch=3, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        ptr = cor.synth_records(777, 3, snr1_db=3.0, snr2_db=6.0)
        raw = cor.from_device(ptr, 3 * N * 2 * 4)
        res = cor.process_device(ptr, 3)
        cor.device_free(ptr)
    assert raw.std() > 1500 and abs(raw[0::4].astype(float).mean() - 11.3) < 60
    freq = oc.freq_axis_G(5e6, N)
    k = oc.window(freq, -20000.0, 20000.0)
    for r in range(3):
        ref = oc.record_G(raw[r * 4 * N:(r + 1) * 4 * N], k, freq, oc.fcode_G(code), code)
        for ch in (1, 2):
            pc.assert_row_matches(res[r, ch - 1], ref[ch], where=f"synthetic record {r} ch{ch}")
        d2 = (res[r, 1]["indice"] - 1) / 3.0                  # loopback delay ~0.7253 of the window
        assert abs(d2 - 0.7253 * N) < 3
        assert 0.24 * N < (res[r, 0]["indice"] - 1) / 3.0 < 1.0 * N

"""Host-side mirror of the reference's per-second correlator interface over the C ABI.

Reference seams mirrored here (paths relative to the reference root):
  Correlator.process          record loop of processing/Octave/godual_ranging.m:75-102
  Correlator.process_windows  processing(d,k,df) of acquisition/claudio_aligned_code_lo_separate.m:24-77
  Correlator.estimate_carrier godual_ranging.m:14-15
  Correlator.doppler_bank     per-bin body of search_df over a grid of offsets (cross-ambiguity search),
                              acquisition/claudio_aligned_code_re_separate.m:33-42
  Correlator.search_df        k=search_df(d,k,df_threshold), claudio_aligned_code_re_separate.m:26-46
  Correlator.gen_code/set_code  amaranth_twstft/common.py:59-73 / godual_ranging.m:62-66
Results come back as a numpy structured array whose field names are the reference's .mat
variable names (indice, correction, SNRr, SNRi, df, puissance, puissancecode, puissancenoise, xval*).
"""
import ctypes as C

import numpy as np

from . import _lib

RESULT_DTYPE = np.dtype([(n, np.float64) for n in _lib.RESULT_FIELDS])
BANK_DTYPE = np.dtype([(n, np.float64) for n in _lib.BANK_FIELDS])


class TwstftError(RuntimeError):
    pass


class Correlator:
    """One plan = one GPU.  Not re-entrant; use one Correlator per host thread / GPU."""

    def __init__(self, nsamples=5000000, fs=5e6, variant="G", modulation="bpsk", layout="interleaved2",
                 window_hz=(-20000.0, 20000.0), freq_axis=None, nint=1, samples_per_chip=2, ddof=1,
                 est_sub_mean=None, max_batch=4, device=0, lib=None, fine_freq=False, code_window=False, fast_peak=False,
                 _plan=None):
        self.lib = _lib.load(lib)
        cfg = _lib.Config()
        self.lib.twstft_default_config(C.byref(cfg))
        cfg.fs = fs
        cfg.nsamples = int(nsamples)
        cfg.samples_per_chip = samples_per_chip
        cfg.variant = _lib.VARIANT_K if str(variant).upper().startswith("K") else _lib.VARIANT_G
        cfg.modulation = _lib.QPSK if str(modulation).lower() == "qpsk" else _lib.BPSK
        cfg.nint = nint
        cfg.layout = _lib.LAYOUT_SINGLE if layout in ("single", 1) else _lib.LAYOUT_INTERLEAVED2
        if freq_axis is None:
            freq_axis = "K" if cfg.variant == _lib.VARIANT_K else "G"
        cfg.freq_axis = _lib.AXIS_K if freq_axis == "K" else _lib.AXIS_G
        cfg.win_lo_hz, cfg.win_hi_hz = float(window_hz[0]), float(window_hz[1])
        cfg.est_sub_mean = int(cfg.variant == _lib.VARIANT_G if est_sub_mean is None else est_sub_mean)
        cfg.ddof = ddof
        cfg.max_batch = max_batch
        cfg.device_id = device
        cfg.fine_freq = int(bool(fine_freq))
        cfg.code_window = int(bool(code_window))     # 1: Hamming (processing/CPP/main.cpp:717-719)
        cfg.fast_peak = int(bool(fast_peak))
        self.cfg = cfg
        self.nch = 2 if cfg.layout == _lib.LAYOUT_INTERLEAVED2 else 1
        self.N = int(nsamples)
        self._owned = _plan is None
        if _plan is not None:            # a plan owned by a Sharder
            self._plan = _plan
            return
        self._plan = C.c_void_p()
        rc = self.lib.twstft_plan_create(C.byref(cfg), C.byref(self._plan))
        if rc != 0:
            msg = self.lib.twstft_last_error(self._plan).decode() if self._plan else "plan creation failed"
            self.close()
            raise TwstftError(f"twstft_plan_create failed ({rc}): {msg}")

    # ------------------------------------------------------------------ lifetime
    def close(self):
        for ptr in getattr(self, "_pinned", []):
            self.lib.twstft_host_free(ptr)
        self._pinned = []
        if getattr(self, "_plan", None):
            if getattr(self, "_owned", True):
                self.lib.twstft_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise TwstftError(f"{what} failed ({rc}): {self.lib.twstft_last_error(self._plan).decode()}")

    @property
    def backend(self):
        return self.lib.twstft_backend().decode()

    @property
    def kernel_launches(self):
        return int(self.lib.twstft_kernel_launches(self._plan))

    @property
    def fast_counters(self):
        """(records through the fast peak path, records re-run through the full path)"""
        a, b = C.c_int64(), C.c_int64()
        self.lib.twstft_fast_counters(self._plan, C.byref(a), C.byref(b))
        return a.value, b.value

    @property
    def fast_suspended(self):
        """Records that went through the full path because the fast path's fall-back rate made it a loss."""
        return int(self.lib.twstft_fast_suspended(self._plan))

    def window_bins(self):
        lo, hi = C.c_int64(), C.c_int64()
        self.lib.twstft_window_bins(self._plan, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    # ------------------------------------------------------------------ PRN replica
    def set_code(self, chips, qpsk=False):
        chips = np.ascontiguousarray(chips, dtype=np.uint8)
        self._check(self.lib.twstft_set_code(self._plan, chips.ctypes.data, chips.size, int(qpsk)), "twstft_set_code")

    def gen_code(self, bitlen, taps_a, taps_b=0, seed=1, noiselen=None):
        noiselen = int(noiselen if noiselen is not None else self.N // self.cfg.samples_per_chip)
        out = np.empty(noiselen * (2 if taps_b else 1), dtype=np.uint8)
        self._check(self.lib.twstft_gen_code(self._plan, bitlen, taps_a, taps_b, seed, noiselen, out.ctypes.data),
                    "twstft_gen_code")
        return out

    # ------------------------------------------------------------------ hot path
    def _records(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        per = 2 * self.nch * self.N
        if iq.size % per:
            raise ValueError(f"input holds {iq.size} int16, not a multiple of one record ({per})")
        return iq, iq.size // per

    def process(self, iq):
        """iq: int16 host array of whole records ([I1 Q1 I2 Q2]... or [I Q]...).  Returns a structured
        array of shape (nrecords, nch)."""
        iq, nrec = self._records(iq)
        out = np.zeros((nrec, self.nch), dtype=RESULT_DTYPE)
        self._check(self.lib.twstft_process(self._plan, iq.ctypes.data, nrec, out.ctypes.data), "twstft_process")
        return out

    def process_device(self, dev_ptr, nrecords):
        out = np.zeros((nrecords, self.nch), dtype=RESULT_DTYPE)
        self._check(self.lib.twstft_process_device(self._plan, dev_ptr, nrecords, out.ctypes.data),
                    "twstft_process_device")
        return out

    def process_windows(self, iq, offsets, df, on_device=False, total_instants=None):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        df = np.ascontiguousarray(np.broadcast_to(np.asarray(df, dtype=np.float64).reshape(len(offsets), -1),
                                                  (len(offsets), self.nch)))
        out = np.zeros((len(offsets), self.nch), dtype=RESULT_DTYPE)
        if on_device:
            ptr, tot = iq, int(total_instants)
        else:
            iq = np.ascontiguousarray(iq, dtype=np.int16)
            ptr, tot = iq.ctypes.data, iq.size // (2 * self.nch)
        self._check(self.lib.twstft_process_windows(self._plan, ptr, int(on_device), tot, offsets.ctypes.data,
                                                    df.ctypes.data, len(offsets), out.ctypes.data),
                    "twstft_process_windows")
        return out

    # ------------------------------------------------------------------ the Octave function seam (complex double windows)
    def _windows_cf64(self, d):
        d = np.ascontiguousarray(d, dtype=np.complex128).reshape(-1)
        if d.size % self.N:
            raise ValueError(f"input holds {d.size} complex samples, not a multiple of one window ({self.N})")
        return d, d.size // self.N

    def processing(self, d, k=None, df=None):
        """The reference's `processing` function on complex double window(s) whose mean the caller has removed.
          processing(d, k)        variant G, godual_ranging.m:12: k = 1-based indices into the fftshift'ed spectrum
                                  (a contiguous range, as k=find(...) gives; None: the plan's window)
          processing(d, df=df)    variant K, claudio_aligned_code_lo_separate.m:24 / ..._re_separate.m:48
        Returns one row per window."""
        d, nwin = self._windows_cf64(d)
        out = np.zeros(nwin, dtype=RESULT_DTYPE)
        if df is not None:
            dfa = np.ascontiguousarray(np.broadcast_to(np.asarray(df, dtype=np.float64).reshape(-1), (nwin,)))
            self._check(self.lib.twstft_processing_cf64_df(self._plan, d.ctypes.data, nwin, dfa.ctypes.data,
                                                           out.ctypes.data), "twstft_processing_cf64_df")
            return out
        k1, k2 = (0, 0) if k is None else (int(k[0]), int(k[-1]))
        if k is not None and k2 - k1 + 1 != len(k):
            raise ValueError("k must be a contiguous index range")
        self._check(self.lib.twstft_processing_cf64(self._plan, d.ctypes.data, nwin, k1, k2, out.ctypes.data),
                    "twstft_processing_cf64")
        return out

    def search_df_cf64(self, cor, d, k=None, df_threshold=20.0, ls=2):
        """k=search_df(d,k,df_threshold) on the complex double block the re/ranging scripts hand over
        (claudio_aligned_code_re_separate.m:26,150).  Returns (kbon, df, ncand)."""
        d = np.ascontiguousarray(d, dtype=np.complex128).reshape(-1)
        if d.size != self.N * ls:
            raise ValueError("search_df takes exactly one block of ls*N samples")
        kbon = np.zeros(1, dtype=np.int64)
        df = np.zeros(1)
        ncand = np.zeros(1, dtype=np.int32)
        k1, k2 = (0, 0) if k is None else (int(k[0]), int(k[-1]))
        self._check(self.lib.twstft_search_df_cf64(self._plan, cor._plan, d.ctypes.data, int(ls), k1, k2,
                                                   float(df_threshold), kbon.ctypes.data, df.ctypes.data,
                                                   ncand.ctypes.data), "twstft_search_df_cf64")
        return int(kbon[0]), float(df[0]), int(ncand[0])

    def estimate_decimated(self, iq, dec=25, foffset=0.0, remote=0):
        """GoRanging::df (processing/CPP/main.cpp:363-450): whole-file estimator on one instant out of `dec`.
        Returns (df[nch], pos[nch]); entries of channels not looked at (remote) stay 0 / -1."""
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        df = np.zeros(2)
        pos = np.full(2, -1, dtype=np.int64)
        self._check(self.lib.twstft_estimate_decimated(self._plan, iq.ctypes.data, iq.size // (2 * self.nch), int(dec),
                                                       float(foffset), int(remote), df.ctypes.data, pos.ctypes.data),
                    "twstft_estimate_decimated")
        return df[:self.nch], pos[:self.nch]

    def estimate_carrier(self, iq, ls=1):
        """Carrier offset of every block of ls*N sample instants (ls=2: the 2-s blocks of the K scripts)."""
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        per = 2 * self.nch * self.N * ls
        if iq.size % per:
            raise ValueError(f"input holds {iq.size} int16, not a multiple of one block ({per})")
        nrec = iq.size // per
        df = np.zeros((nrec, self.nch))
        bins = np.zeros((nrec, self.nch), dtype=np.int64)
        self._check(self.lib.twstft_estimate_carrier_block(self._plan, iq.ctypes.data, 0, nrec, int(ls), df.ctypes.data,
                                                           bins.ctypes.data), "twstft_estimate_carrier_block")
        return df, bins

    def doppler_bank(self, iq, dfs, offset=0, sub_mean=True):
        """Correlate the window of N sample instants at `offset` of the host buffer `iq` at 1x for every carrier
        offset in dfs; returns a structured array (nbins, nch) with b0, prnsig, metric, status."""
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        dfs = np.ascontiguousarray(dfs, dtype=np.float64)
        out = np.zeros((len(dfs), self.nch), dtype=BANK_DTYPE)
        self._check(self.lib.twstft_doppler_bank(self._plan, iq.ctypes.data, 0, iq.size // (2 * self.nch), int(offset),
                                                 dfs.ctypes.data, len(dfs), int(sub_mean), out.ctypes.data),
                    "twstft_doppler_bank")
        return out

    def acquire(self, iq, fc_init=0.0, frange=2000.0, fstep=250.0, min_step=1.0, offset=0, sub_mean=True, channel=0):
        """Step-halving Doppler / code-phase acquisition (experiments/231001_DLL_PLL/rxcomplex.cpp:522-568) on the window
        at `offset`: returns (fc, bank row of the best bin, bins evaluated)."""
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        fc = C.c_double()
        nb = C.c_int32()
        row = np.zeros(1, dtype=BANK_DTYPE)
        self._check(self.lib.twstft_acquire(self._plan, iq.ctypes.data, 0, iq.size // (2 * self.nch), int(offset),
                                            float(fc_init), float(frange), float(fstep), float(min_step), int(sub_mean),
                                            int(channel), C.byref(fc), row.ctypes.data, C.byref(nb)), "twstft_acquire")
        return fc.value, row[0], nb.value

    def search_df(self, iq, df_threshold=20.0, offset=0):
        """k=search_df(d,k,df_threshold): returns (kbon[nch] 1-based shifted-axis index or 0, df[nch], ncand[nch])."""
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        kbon = np.zeros(self.nch, dtype=np.int64)
        df = np.zeros(self.nch)
        ncand = np.zeros(self.nch, dtype=np.int32)
        self._check(self.lib.twstft_search_df(self._plan, iq.ctypes.data, 0, iq.size // (2 * self.nch), int(offset),
                                              float(df_threshold), kbon.ctypes.data, df.ctypes.data, ncand.ctypes.data),
                    "twstft_search_df")
        return kbon, df, ncand

    def search_df_block(self, cor, iq, ls=2, df_threshold=20.0, on_device=False):
        """search_df as the re/ranging scripts call it: spectrum of the whole block of ls*N sample instants (this plan),
        candidates tested on the first cor.N instants against cor's code.  Returns (kbon, df, ncand) per channel."""
        kbon = np.zeros(self.nch, dtype=np.int64)
        df = np.zeros(self.nch)
        ncand = np.zeros(self.nch, dtype=np.int32)
        if on_device:
            ptr = iq
        else:
            iq = np.ascontiguousarray(iq, dtype=np.int16)
            if iq.size != 2 * self.nch * self.N * ls:
                raise ValueError("search_df_block takes exactly one block of ls*N sample instants")
            ptr = iq.ctypes.data
        self._check(self.lib.twstft_search_df_block(self._plan, cor._plan, ptr, int(on_device), int(ls),
                                                    float(df_threshold), kbon.ctypes.data, df.ctypes.data,
                                                    ncand.ctypes.data), "twstft_search_df_block")
        return kbon, df, ncand

    def estimate_bins(self, iq, s_first, nbins, on_device=False, ninstants=None):
        """|fft(x.^p)| at the 0-based fftshift'ed positions s_first..s_first+nbins-1 of a buffer of any length."""
        mag = np.zeros((self.nch, nbins))
        if on_device:
            ptr, n = iq, int(ninstants)
        else:
            iq = np.ascontiguousarray(iq, dtype=np.int16)
            ptr, n = iq.ctypes.data, iq.size // (2 * self.nch)
        self._check(self.lib.twstft_estimate_bins(self._plan, ptr, int(on_device), n, int(s_first), int(nbins),
                                                  mag.ctypes.data), "twstft_estimate_bins")
        return mag

    # ------------------------------------------------------------------ timing (CUDA events on the plan's stream)
    def synchronize(self):
        self._check(self.lib.twstft_synchronize(self._plan), "twstft_synchronize")

    def timer_start(self):
        self._check(self.lib.twstft_timer_start(self._plan), "twstft_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.lib.twstft_timer_stop(self._plan, C.byref(ms)), "twstft_timer_stop")
        return float(ms.value)

    def profile(self, on):
        self.lib.twstft_profile_enable(self._plan, int(on))

    def profile_read(self):
        """{kernel name: (total ms, launches)} accumulated while profiling was enabled."""
        out, slot = {}, 0
        name = C.create_string_buffer(64)
        ms, cnt = C.c_double(), C.c_int64()
        while self.lib.twstft_profile_read(self._plan, slot, name, 64, C.byref(ms), C.byref(cnt)) == 0:
            out[name.value.decode()] = (ms.value, cnt.value)
            slot += 1
        return out

    def pinned_empty(self, nbytes):
        """uint8 numpy view over pinned host memory (twstft_host_alloc); keep the Correlator alive."""
        ptr = self.lib.twstft_host_alloc(nbytes)
        if not ptr:
            raise TwstftError("twstft_host_alloc failed")
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=np.uint8)
        self._pinned = getattr(self, "_pinned", []) + [ptr]
        return arr

    # ------------------------------------------------------------------ device memory helpers
    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        ptr = C.c_void_p()
        self._check(self.lib.twstft_device_alloc(self._plan, arr.nbytes, C.byref(ptr)), "twstft_device_alloc")
        self._check(self.lib.twstft_copy_to_device(self._plan, ptr, arr.ctypes.data, arr.nbytes), "twstft_copy_to_device")
        return ptr

    def device_free(self, ptr):
        self.lib.twstft_device_free(self._plan, ptr)

    def from_device(self, ptr, nbytes, dtype=np.int16):
        out = np.zeros(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self._check(self.lib.twstft_copy_to_host(self._plan, out.ctypes.data, ptr, out.nbytes), "twstft_copy_to_host")
        return out

    # ------------------------------------------------------------------ synthetic inputs on the device (bench tooling)
    def synth_records(self, seed0, nrecords, snr1_db=-5.0, snr2_db=0.6):
        """nrecords synthetic records in the plan's layout, generated in device memory (k_synth.cuh); returns the
        device pointer (free with device_free)."""
        ptr = C.c_void_p()
        nbytes = nrecords * self.N * self.nch * 4
        self._check(self.lib.twstft_device_alloc(self._plan, nbytes, C.byref(ptr)), "twstft_device_alloc")
        self._check(self.lib.twstft_synth_records(self._plan, int(seed0), nrecords, float(snr1_db), float(snr2_db), ptr),
                    "twstft_synth_records")
        return ptr

    def synth_stream(self, seed, ninstants, delay=0.0, df=0.0, snr_db=-5.0):
        ptr = C.c_void_p()
        self._check(self.lib.twstft_device_alloc(self._plan, ninstants * 4, C.byref(ptr)), "twstft_device_alloc")
        self._check(self.lib.twstft_synth_stream(self._plan, int(seed), int(ninstants), float(delay), float(df),
                                                 float(snr_db), ptr), "twstft_synth_stream")
        return ptr

    def process_cycled(self, pool, pool_records, first_record, nrecords, on_device=False):
        """nrecords records first_record.. drawn cyclically from a pool of pool_records records (host array or device
        pointer): the day-of-data driver."""
        out = np.zeros((nrecords, self.nch), dtype=RESULT_DTYPE)
        ptr = pool if on_device else np.ascontiguousarray(pool, dtype=np.int16).ctypes.data
        self._check(self.lib.twstft_process_cycled(self._plan, ptr, int(on_device), pool_records, first_record, nrecords,
                                                   out.ctypes.data), "twstft_process_cycled")
        return out

    def debug_fetch(self, what, lane, count, dtype):
        out = np.zeros(count, dtype=dtype)
        self._check(self.lib.twstft_debug_fetch(self._plan, what, lane, out.ctypes.data, out.nbytes), "twstft_debug_fetch")
        return out


def delay_samples(res):
    """Printed delay of godual_ranging.m:96-98 in samples: (indice-1+correction)/3."""
    return (res["indice"] - 1.0 + res["correction"]) / 3.0


class Sharder:
    """All GPUs of one box behind one object: one host thread + one plan per GPU inside the C library
    (twstft_shard_*), GPU g of G takes records [g*R/G, (g+1)*R/G), rows land at their record index.
    Mirrors the record loop of godual_ranging.m:75-102 / processing/CPP/main.cpp:452-519 over many GPUs."""

    def __init__(self, ngpus, devices=None, lib=None, **cfg_kw):
        probe = Correlator.__new__(Correlator)          # reuse the keyword -> twstft_config translation
        Correlator.__init__(probe, lib=lib, _plan=C.c_void_p(0), **cfg_kw)
        self.lib, self.cfg, self.nch, self.N = probe.lib, probe.cfg, probe.nch, probe.N
        self._shard = C.c_void_p()
        dev = (C.c_int * ngpus)(*devices) if devices is not None else None
        rc = self.lib.twstft_shard_create(C.byref(self.cfg), dev, ngpus, C.byref(self._shard))
        if rc != 0:
            msg = self.lib.twstft_shard_last_error(self._shard).decode() if self._shard else "shard creation failed"
            self.close()
            raise TwstftError(f"twstft_shard_create failed ({rc}): {msg}")
        self.ngpus = ngpus

    def close(self):
        if getattr(self, "_shard", None):
            self.lib.twstft_shard_destroy(self._shard)
            self._shard = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise TwstftError(f"{what} failed ({rc}): {self.lib.twstft_shard_last_error(self._shard).decode()}")

    def plan(self, g):
        """Correlator view of GPU g's plan (owned by the shard)."""
        kw = dict(nsamples=self.N, lib=None)
        c = Correlator.__new__(Correlator)
        c.lib, c.cfg, c.nch, c.N = self.lib, self.cfg, self.nch, self.N
        c._plan, c._owned, c._pinned = C.c_void_p(self.lib.twstft_shard_plan(self._shard, g)), False, []
        return c

    def set_code(self, chips, qpsk=False):
        chips = np.ascontiguousarray(chips, dtype=np.uint8)
        self._check(self.lib.twstft_shard_set_code(self._shard, chips.ctypes.data, chips.size, int(qpsk)),
                    "twstft_shard_set_code")

    def process(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        per = 2 * self.nch * self.N
        nrec = iq.size // per
        out = np.zeros((nrec, self.nch), dtype=RESULT_DTYPE)
        self._check(self.lib.twstft_shard_process(self._shard, iq.ctypes.data, nrec, out.ctypes.data), "twstft_shard_process")
        return out

    def process_cycled(self, pool, nrecords, resident=True):
        """nrecords records drawn cyclically from the host pool; returns (rows, wall seconds, per-GPU device seconds)."""
        pool = np.ascontiguousarray(pool, dtype=np.int16)
        per = 2 * self.nch * self.N
        out = np.zeros((nrecords, self.nch), dtype=RESULT_DTYPE)
        wall = C.c_double()
        dev = (C.c_double * self.ngpus)()
        self._check(self.lib.twstft_shard_process_cycled(self._shard, pool.ctypes.data, pool.size // per, nrecords,
                                                         int(bool(resident)), out.ctypes.data, C.byref(wall), dev),
                    "twstft_shard_process_cycled")
        return out, wall.value, list(dev)

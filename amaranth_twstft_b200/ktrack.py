"""K-variant front end: the windowing / tracking loops of the production scripts over the C ABI.

mode "lo" mirrors acquisition/claudio_aligned_code_lo_separate.m:114-164 (one X310 channel per file, 2-s blocks,
100 kchip/s code => N = 200 000 samples per window, 25 windows per second):

    per block   df   <- carrier estimator on the raw 2-s block            (:126-129)  twstft_estimate_carrier_block
                d    <- [dold ; block]                                    (:127)
    per window  processing(d(dindex:dindex+N-1)-mean, k, df)              (:131-133)  twstft_process_windows
                indice1 = floor(i/3); MOVED test and re-slice             (:134-146)
                dindex += N until the window would leave d                (:152-153)
    tail        dold <- d(dindex:end)                                     (:155-159)

mode "re" mirrors acquisition/claudio_aligned_code_re_separate.m:126-203 (remote; "ranging" is the same script
with the |f| < 8 kHz window, claudio_aligned_code_ranging_separate.m):

    start       fseek 30 s; kbon <- search_df(block)                      (:127,148-151) twstft_search_df_block
                then re-open the file and start from its first block      (:152-154)
    per block   d <- [dold ; block]; df <- max of |fft(d.^2)| over kbon-3:kbon+3, a transform of
                len(dold)+fs*ls points read on the fs*ls-point freq axis  (:165-167)     twstft_estimate_bins
    per window  processing(dpart, df); indice1 = i/3 NOT floored, so dindex turns fractional after a MOVED
                re-slice and every slice is taken at round(dindex)        (:170-185)

The loops are sequential (every window start depends on the previous peak) but moves are rare, so windows are
submitted to the GPU in speculative batches at the offsets they will have if nothing moves; the results are
scanned in order and the batch is cut at the first MOVED window, which is redone at its corrected offset.
Host control code only: every number in the output comes from the CUDA library.
Quirks kept: after a MOVED re-slice `indice1(p)` holds the raw interpolated-grid index of the second
processing() call (the scripts do not divide it again, lo :145 / re :185); `puissancecode/puissancenoise` keep the
last window only.  A re-slice that would leave the block is an Octave index error in the scripts; here the
block is abandoned at that window and `status` says so.  While the carrier is not found the re script re-opens
the file on every iteration and never ends; here the search continues with blocks 1, 2, ... and
process_file returns None when no block holds the carrier.

processing/C/claudio_aligned.c is the same front end in C (the production binary); this module is the Python
mirror the parity tests drive.

The C program processing/C/claudio_aligned.c is the authoritative implementation of this loop; this Python mirror over the same
C ABI exists so that the parity tests can drive the loop step by step.
"""
import numpy as np

from .api import Correlator

WINDOWS = {            # (mode, OP) -> estimator window in Hz on freq=linspace(-fs/2,fs/2-1,fs*ls)
    ("lo", 0): (-20000.0, 20000.0), ("lo", 1): (-20000.0, 20000.0),            # lo_separate.m:104-105
    ("lo-remote", 1): (-120000.0, -80000.0), ("lo-remote", 0): (100000.0, 120000.0),   # lo_separate.m:107-111
    ("re", 1): (-108000.0, -92000.0), ("re", 0): (92000.0, 108000.0),          # re_separate.m:136-140
    ("ranging", 0): (-8000.0, 8000.0), ("ranging", 1): (-8000.0, 8000.0),      # ranging_separate.m:135
}
PREFIX = {"lo": "localclaudio", "lo-remote": "remoteclaudio", "re": "remoteclaudio", "ranging": "rangingclaudio"}


def _oround(x):
    """Octave round(): half away from zero (window starts are positive)."""
    return int(np.floor(x + 0.5))


class KTracker:
    def __init__(self, chips, fs=5e6, ls=2, window_hz=None, samples_per_chip=2, max_batch=32,
                 lib=None, device=0, mode="lo", op=0, df_threshold=20.0, skip_s=30.0):
        if mode not in PREFIX:
            raise ValueError(f"mode {mode!r}: expected one of {sorted(PREFIX)}")
        chips = np.ascontiguousarray(chips, dtype=np.uint8)
        self.fs, self.ls = float(fs), int(ls)
        self.mode, self.search = mode, mode in ("re", "ranging")
        self.df_threshold, self.skip_s = float(df_threshold), float(skip_s)
        if window_hz is None:
            window_hz = WINDOWS[(mode, int(op))]
        self.N = int(len(chips) * samples_per_chip)
        self.code = 2.0 * np.repeat(chips.astype(np.float64), samples_per_chip) - 1.0      # :93-95
        # block estimator: freq=linspace(-fs/2,fs/2-fs/fs,fs*ls), raw samples (:103,126)
        self.est = Correlator(nsamples=int(fs), fs=fs, variant="K", layout="single", freq_axis="K",
                              window_hz=window_hz, est_sub_mean=0, max_batch=1, lib=lib, device=device)
        self.cor = Correlator(nsamples=self.N, fs=fs, variant="K", layout="single", max_batch=max_batch,
                              samples_per_chip=samples_per_chip, lib=lib, device=device, fast_peak=False)   # like claudio_aligned.c (peak parked at the window start)
        self.cor.set_code(chips)
        self.max_batch = max_batch
        nax = int(fs) * self.ls
        self.freq = lambda s: (fs / 2 - 1.0) if s == nax - 1 else -fs / 2 + s * ((fs / 2 - 1.0 + fs / 2) / (nax - 1))

    def close(self):
        self.est.close()
        self.cor.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def find_carrier(self, raw):
        """re/ranging: kbon (1-based index on the shifted block axis) of the first block where search_df succeeds,
        tried in the order 30-s block, block 1, block 2, ...; 0 if none."""
        blk = int(self.fs) * self.ls * 2
        nblocks = raw.size // blk
        skip = int(self.skip_s * self.fs) * 2                          # fseek(f,30*fs*2*2) bytes (:127)
        starts = ([skip] if skip + blk <= raw.size else []) + [b * blk for b in range(1, nblocks)]
        for s0 in starts:
            kbon = int(self.est.search_df_block(self.cor, raw[s0:s0 + blk], ls=self.ls,
                                                df_threshold=self.df_threshold)[0][0])
            if kbon:
                return kbon
        return 0

    def process_file(self, raw):
        """raw: int16 [I Q] samples of one channel file.  Returns a dict with the variables the scripts save
        (lo :166 / re :204): xval1, indice1, correction1, SNR1r, SNR1i, puissance1, df, moved, movedval,
        puissancecode, puissancenoise, code (+ status, kbon); None when re/ranging finds no carrier."""
        raw = np.ascontiguousarray(raw, dtype=np.int16)
        N, blk = self.N, int(self.fs) * self.ls * 2                    # int16 per block (:118)
        cols = {k: [] for k in ("xval1", "indice1", "correction1", "SNR1r", "SNR1i", "puissance1")}
        dfs, moved, movedval, status = [], [], [], []
        pcode = pnoise = np.nan
        dold = np.zeros(0, dtype=np.int16)
        kbon = 0
        if self.search:
            kbon = self.find_carrier(raw)
            if not kbon:
                return None

        def keep(row, indice):
            nonlocal pcode, pnoise
            cols["xval1"].append(complex(row["xval_re"], row["xval_im"]))
            cols["indice1"].append(float(indice))
            for k, f in (("correction1", "correction"), ("SNR1r", "SNRr"), ("SNR1i", "SNRi"), ("puissance1", "puissance")):
                cols[k].append(float(row[f]))
            pcode, pnoise = float(row["puissancecode"]), float(row["puissancenoise"])

        for b0 in range(0, raw.size - blk + 1, blk):
            block = raw[b0:b0 + blk]
            d = np.concatenate([dold, block])
            L = d.size // 2
            dev = self.cor.to_device(d)
            try:
                if self.search:      # max over kbon-3:kbon+3 of the spectrum of [dold ; block] (re :166-167)
                    mag = self.est.estimate_bins(dev, kbon - 4, 7, on_device=True, ninstants=L)[0]
                    df = float(self.freq(int(np.argmax(mag)) + kbon - 4) / 2)
                else:                # lo: raw 2-s block before dold is prepended (:126,129)
                    df = float(self.est.estimate_carrier(block, ls=self.ls)[0][0, 0])
                dfs.append(df)
                dindex, ok = 1.0, True                                 # 1-based, as in the script
                while ok:
                    nwin = min(self.max_batch, int((L - N + 1 - dindex) // N) + 1)
                    if nwin < 1:
                        nwin = 1                                       # do ... until: the body runs at least once
                    offs = [_oround(dindex + t * N) - 1 for t in range(nwin)]
                    while len(offs) > 1 and offs[-1] + N > L:          # round() pushed the last start past the block
                        offs.pop()
                    nwin = len(offs)
                    if offs[0] + N > L:
                        status.append("block shorter than one window")
                        break
                    res = self.cor.process_windows(dev, offs, [df] * nwin, on_device=True, total_instants=L)
                    for t in range(nwin):
                        row = res[t, 0]
                        ind = row["indice"] / 3.0                      # re :173
                        if not self.search:
                            ind = np.floor(ind)                        # lo :134
                        snr = 10 * np.log10(row["SNRi"] + row["SNRr"])
                        cur = dindex + t * N
                        if snr > -30 and ((43 < ind < N / 2) or (N / 2 < ind < N - 2)):     # :135-136
                            moved.append(len(cols["indice1"]) + 1)
                            movedval.append(ind + 1)
                            if cur - ind + 1 < 0:
                                cur += N
                            cur = cur - ind + 21                       # :142 (230802)
                            if _oround(cur) < 1 or _oround(cur) + N - 1 > L:
                                status.append(f"window {len(cols['indice1']) + 1}: re-slice leaves the block")
                                ok = False
                                break
                            r2 = self.cor.process_windows(dev, [_oround(cur) - 1], [df], on_device=True,
                                                          total_instants=L)[0, 0]
                            keep(r2, r2["indice"])                     # :145, not divided again
                            dindex = cur + N
                            break
                        keep(row, ind)
                        if t == nwin - 1:
                            dindex = cur + N
                    if dindex + N - 1 > L:                             # :153
                        break
            finally:
                self.cor.device_free(dev)
            dold = d[2 * (_oround(dindex) - 1):].copy() if dindex < L else np.zeros(0, dtype=np.int16)   # :155-159

        out = {k: np.asarray(v) for k, v in cols.items()}
        out.update(df=np.asarray(dfs), moved=np.asarray(moved, dtype=np.float64),
                   movedval=np.asarray(movedval, dtype=np.float64), puissancecode=pcode, puissancenoise=pnoise,
                   code=self.code, status=status, kbon=kbon)
        return out


def save_mat(path, out):
    """`save -mat nomout corr* df indic* SNR* code puissan* xval* moved*` (:166): MAT-v5, 1xP row vectors."""
    import scipy.io
    names = ("correction1", "df", "indice1", "SNR1r", "SNR1i", "code", "puissance1", "puissancecode",
             "puissancenoise", "xval1", "moved", "movedval")
    scipy.io.savemat(path, {k: np.atleast_2d(np.asarray(out[k])) for k in names}, format="5", oned_as="row")

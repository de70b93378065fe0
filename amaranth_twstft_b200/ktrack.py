"""K-variant front end: the windowing / tracking loop of the production scripts over the C ABI.

Mirrors acquisition/claudio_aligned_code_lo_separate.m:114-164 (one X310 channel per file, 2-s blocks,
100 kchip/s code => N = 200 000 samples per window, 25 windows per second):

    per block   df   <- carrier estimator on the raw 2-s block            (:126-129)  twstft_estimate_carrier_block
                d    <- [dold ; block]                                    (:127)
    per window  processing(d(dindex:dindex+N-1)-mean, k, df)              (:131-133)  twstft_process_windows
                indice1 = floor(i/3); MOVED test and re-slice             (:134-146)
                dindex += N until the window would leave d                (:152-153)
    tail        dold <- d(dindex:end)                                     (:155-159)

The loop is sequential (every window start depends on the previous peak) but moves are rare, so windows are
submitted to the GPU in speculative batches at the offsets they will have if nothing moves; the results are
scanned in order and the batch is cut at the first MOVED window, which is redone at its corrected offset.
Host control code only: every number in the output comes from the CUDA library.
Quirks kept: after a MOVED re-slice `indice1(p)` holds the raw interpolated-grid index of the second
processing() call (the script does not floor it again, :145); `puissancecode/puissancenoise` keep the last
window only (:133).  A re-slice that would leave the block is an Octave index error in the script; here the
block is abandoned at that window and `status` says so.
"""
import numpy as np

from .api import Correlator


class KTracker:
    def __init__(self, chips, fs=5e6, ls=2, window_hz=(-20000.0, 20000.0), samples_per_chip=2, max_batch=32,
                 lib=None, device=0):
        chips = np.ascontiguousarray(chips, dtype=np.uint8)
        self.fs, self.ls = float(fs), int(ls)
        self.N = int(len(chips) * samples_per_chip)
        self.code = 2.0 * np.repeat(chips.astype(np.float64), samples_per_chip) - 1.0      # :93-95
        # block estimator: freq=linspace(-fs/2,fs/2-fs/fs,fs*ls), raw samples (:103,126)
        self.est = Correlator(nsamples=int(fs), fs=fs, variant="K", layout="single", freq_axis="K",
                              window_hz=window_hz, est_sub_mean=0, max_batch=1, lib=lib, device=device)
        self.cor = Correlator(nsamples=self.N, fs=fs, variant="K", layout="single", max_batch=max_batch,
                              samples_per_chip=samples_per_chip, lib=lib, device=device)
        self.cor.set_code(chips)
        self.max_batch = max_batch

    def close(self):
        self.est.close()
        self.cor.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def process_file(self, raw):
        """raw: int16 [I Q] samples of one channel file.  Returns a dict with the variables the script saves
        (:166): xval1, indice1, correction1, SNR1r, SNR1i, puissance1, df, moved, movedval, puissancecode,
        puissancenoise, code (+ status)."""
        raw = np.ascontiguousarray(raw, dtype=np.int16)
        N, blk = self.N, int(self.fs) * self.ls * 2                    # int16 per block (:118)
        cols = {k: [] for k in ("xval1", "indice1", "correction1", "SNR1r", "SNR1i", "puissance1")}
        dfs, moved, movedval, status = [], [], [], []
        pcode = pnoise = np.nan
        dold = np.zeros(0, dtype=np.int16)

        def keep(row, indice):
            nonlocal pcode, pnoise
            cols["xval1"].append(complex(row["xval_re"], row["xval_im"]))
            cols["indice1"].append(float(indice))
            for k, f in (("correction1", "correction"), ("SNR1r", "SNRr"), ("SNR1i", "SNRi"), ("puissance1", "puissance")):
                cols[k].append(float(row[f]))
            pcode, pnoise = float(row["puissancecode"]), float(row["puissancenoise"])

        for b0 in range(0, raw.size - blk + 1, blk):
            block = raw[b0:b0 + blk]
            df = float(self.est.estimate_carrier(block, ls=self.ls)[0][0, 0])
            dfs.append(df)
            d = np.concatenate([dold, block])
            L = d.size // 2
            dev = self.cor.to_device(d)
            try:
                dindex, ok = 1, True                                   # 1-based, as in the script
                while ok:
                    nwin = min(self.max_batch, (L - N + 1 - dindex) // N + 1)
                    if nwin < 1:
                        nwin = 1                                       # do ... until: the body runs at least once
                    offs = [dindex - 1 + t * N for t in range(nwin)]
                    if offs[0] + N > L:
                        status.append("block shorter than one window")
                        break
                    res = self.cor.process_windows(dev, offs, [df] * nwin, on_device=True, total_instants=L)
                    for t in range(nwin):
                        row = res[t, 0]
                        ind = np.floor(row["indice"] / 3.0)            # :134
                        snr = 10 * np.log10(row["SNRi"] + row["SNRr"])
                        cur = offs[t] + 1
                        if snr > -30 and ((43 < ind < N / 2) or (N / 2 < ind < N - 2)):     # :135-136
                            moved.append(len(cols["indice1"]) + 1)
                            movedval.append(ind + 1)
                            if cur - ind + 1 < 0:
                                cur += N
                            cur = int(cur - ind + 21)                  # :142 (230802)
                            if cur < 1 or cur + N - 1 > L:
                                status.append(f"window {len(cols['indice1']) + 1}: re-slice leaves the block")
                                ok = False
                                break
                            r2 = self.cor.process_windows(dev, [cur - 1], [df], on_device=True, total_instants=L)[0, 0]
                            keep(r2, r2["indice"])                     # :145, not floored again
                            dindex = cur + N
                            break
                        keep(row, ind)
                        dindex = cur + N
                    if dindex + N - 1 > L:                             # :153
                        break
            finally:
                self.cor.device_free(dev)
            dold = d[2 * (dindex - 1):].copy() if dindex < L else np.zeros(0, dtype=np.int16)   # :155-159

        out = {k: np.asarray(v) for k, v in cols.items()}
        out.update(df=np.asarray(dfs), moved=np.asarray(moved, dtype=np.float64),
                   movedval=np.asarray(movedval, dtype=np.float64), puissancecode=pcode, puissancenoise=pnoise,
                   code=self.code, status=status)
        return out


def save_mat(path, out):
    """`save -mat nomout corr* df indic* SNR* code puissan* xval* moved*` (:166): MAT-v5, 1xP row vectors."""
    import scipy.io
    names = ("correction1", "df", "indice1", "SNR1r", "SNR1i", "code", "puissance1", "puissancecode",
             "puissancenoise", "xval1", "moved", "movedval")
    scipy.io.savemat(path, {k: np.atleast_2d(np.asarray(out[k])) for k in names}, format="5", oned_as="row")

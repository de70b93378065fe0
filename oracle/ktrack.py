"""ORACLE (test infrastructure): numpy fp64 restatement of the K-variant block/window loop,
acquisition/claudio_aligned_code_lo_separate.m:114-164, statement by statement (no batching, no
speculation).  Only tests/ may import it."""
import numpy as np

from . import correlator as oc


def lo_separate(raw_int16, code, fs=5e6, ls=2, window_hz=(-20000.0, 20000.0)):
    n = len(code)
    fcode = oc.fcode_K(code)                                  # :96
    freq = oc.freq_axis_K(fs, ls)                             # :103
    k = oc.window(freq, window_hz[0], window_hz[1])           # :105-113
    blk = int(fs) * 2 * ls
    out = {key: [] for key in ("xval1", "indice1", "correction1", "SNR1r", "SNR1i", "puissance1")}
    dfs, moved, movedval = [], [], []
    pcode = pnoise = np.nan
    dold = np.zeros(0, dtype=np.complex128)

    def run(dpart, df):
        return oc.processing_K(dpart - np.mean(dpart), df, fcode, code, fs=fs)

    def keep(r, indice):
        nonlocal pcode, pnoise
        out["xval1"].append(r["xval"]); out["indice1"].append(float(indice))
        out["correction1"].append(r["correction"]); out["SNR1r"].append(r["SNRr"]); out["SNR1i"].append(r["SNRi"])
        out["puissance1"].append(r["puissance"])
        pcode, pnoise = r["puissancecode"], r["puissancenoise"]

    for b0 in range(0, len(raw_int16) - blk + 1, blk):        # :117-119 whole blocks only
        dblk = oc.unpack_single(raw_int16[b0:b0 + blk])
        s, df = oc.carrier_estimate(dblk, k, freq)            # :126,129 raw block, before dold is prepended
        dfs.append(df)
        d = np.concatenate([dold, dblk])                      # :127
        dindex = 1
        while True:                                           # do ... until (:130-153)
            r = run(d[dindex - 1:dindex - 1 + n], df)
            ind = np.floor(r["indice"] / 3)                   # :134
            if 10 * np.log10(r["SNRi"] + r["SNRr"]) > -30 and ((43 < ind < n / 2) or (n / 2 < ind < n - 2)):
                moved.append(len(out["indice1"]) + 1)
                movedval.append(ind + 1)
                if dindex - ind + 1 < 0:
                    dindex += n
                dindex = int(dindex - ind + 21)
                r = run(d[dindex - 1:dindex - 1 + n], df)
                ind = r["indice"]                             # :145 overwrites indice1(p) with the raw index
            keep(r, ind)
            dindex += n
            if dindex + n - 1 > len(d):
                break
        dold = d[dindex - 1:] if dindex < len(d) else np.zeros(0, dtype=np.complex128)
    res = {key: np.asarray(v) for key, v in out.items()}
    res.update(df=np.asarray(dfs), moved=np.asarray(moved, dtype=np.float64),
               movedval=np.asarray(movedval, dtype=np.float64), puissancecode=pcode, puissancenoise=pnoise)
    return res

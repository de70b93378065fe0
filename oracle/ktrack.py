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


def _oround(x):
    """Octave round(): half away from zero (window starts are positive)."""
    return int(np.floor(x + 0.5))


def re_separate(raw_int16, code, fs=5e6, ls=2, window_hz=(92000.0, 108000.0), df_threshold=20.0, skip_s=30.0):
    """acquisition/claudio_aligned_code_re_separate.m:126-203 (= ..._ranging_separate.m with |f|<8 kHz), statement
    by statement: fseek 30 s in (:127), search_df on the block read there (:148-151), re-open and start again from the
    first block (:152-154), then per block  d=[dold;block]  (:165), carrier = max over kbon-3:kbon+3 of the spectrum of
    THAT (variable-length) d read on the fs*ls-point freq axis (:166-167), windows with indice1=i/3 un-floored (:173)
    so that dindex becomes fractional and every slice uses round(dindex) (:170,184).
    Deviation shared with the build: while the carrier is not found the script re-opens the file on
    every iteration and so tests the 30-s block, then block 1 for ever; here the search goes on with blocks 2, 3, ...
    and returns None if no block holds the carrier."""
    n = len(code)
    fcode = oc.fcode_K(code)
    freq = oc.freq_axis_K(fs, ls)
    k = oc.window(freq, window_hz[0], window_hz[1])
    blk = int(fs) * 2 * ls
    nblocks = len(raw_int16) // blk
    skip = int(skip_s * fs) * 2                                # fseek(f,30*fs*2*2) bytes = 30*fs*2 int16
    kbon = 0
    starts = ([skip] if skip + blk <= len(raw_int16) else []) + [b * blk for b in range(1, nblocks)]
    for s0 in starts:
        kbon, _ = oc.search_df(oc.unpack_single(raw_int16[s0:s0 + blk]), k, df_threshold, freq, fcode, fs=fs)
        if kbon:
            break
    if not kbon:
        return None
    out = {key: [] for key in ("xval1", "indice1", "correction1", "SNR1r", "SNR1i", "puissance1")}
    dfs, moved, movedval = [], [], []
    pcode = pnoise = np.nan
    dold = np.zeros(0, dtype=np.complex128)

    def run(d, dindex, df):
        a = _oround(dindex)
        dpart = d[a - 1:a - 1 + n]
        return oc.processing_K(dpart - np.mean(dpart), df, fcode, code, fs=fs)

    for b0 in range(0, nblocks * blk, blk):
        d = np.concatenate([dold, oc.unpack_single(raw_int16[b0:b0 + blk])])           # :165
        d2 = np.fft.fftshift(np.abs(np.fft.fft(d * d)))                                 # :166
        sel = d2[kbon - 4:kbon + 3]                                                     # d2(kbon-3:kbon+3)
        idx = int(np.argmax(sel)) + 1 + kbon - 3 - 1                                    # 1-based (:167)
        df = freq[idx - 1] / 2
        dfs.append(df)
        dindex = 1.0
        while True:
            r = run(d, dindex, df)
            ind = r["indice"] / 3                                                       # :173 not floored
            if 10 * np.log10(r["SNRi"] + r["SNRr"]) > -30 and ((43 < ind < n / 2) or (n / 2 < ind < n - 2)):
                moved.append(len(out["indice1"]) + 1)
                movedval.append(ind + 1)
                if dindex - ind + 1 < 0:
                    dindex += n
                dindex = dindex - ind + 21                                              # :183
                r = run(d, dindex, df)
                ind = r["indice"]                                                       # :185 raw index kept
            out["xval1"].append(r["xval"]); out["indice1"].append(float(ind))
            out["correction1"].append(r["correction"]); out["SNR1r"].append(r["SNRr"]); out["SNR1i"].append(r["SNRi"])
            out["puissance1"].append(r["puissance"])
            pcode, pnoise = r["puissancecode"], r["puissancenoise"]
            dindex += n
            if dindex + n - 1 > len(d):
                break
        dold = d[_oround(dindex) - 1:] if dindex < len(d) else np.zeros(0, dtype=np.complex128)
    res = {key: np.asarray(v) for key, v in out.items()}
    res.update(df=np.asarray(dfs), moved=np.asarray(moved, dtype=np.float64),
               movedval=np.asarray(movedval, dtype=np.float64), puissancecode=pcode, puissancenoise=pnoise, kbon=kbon)
    return res

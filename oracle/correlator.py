"""ORACLE (test infrastructure): numpy fp64 restatement of the per-second PRN correlator.

Follows, line by line where it matters:
  variant G : processing/Octave/godual_ranging.m:12-49 (processing) and :57-102 (record loop)
  variant K : acquisition/claudio_aligned_code_lo_separate.m:24-77, ..._re_separate.m:26-101
  python ref: experiments/221219_twoway/processing/godual_ranging.py:17-68 (G + fine frequency,
              np.var i.e. ddof=0) -- used to PIN this oracle (tools/make_golden.py)
  QPSK      : experiments/220822_qpsk_vs_bpsk/goqpsk.m:5-16 (replica), Doc/07_Demodulation.md:60-63 (x^4)
Octave semantics honoured: 1-based `indice`, first-max `max`, var = sum|x-mean|^2/(n-1),
ifft scale 1/n, inclusive linspace.  All indices RETURNED here are the reference's own
(1-based where the reference stores 1-based values); `p0` fields are 0-based helpers.
"""
import numpy as np

try:  # scipy's pocketfft is the same algorithm as numpy's, but lets the baseline use threads
    import scipy.fft as _fft
except Exception:  # pragma: no cover
    _fft = np.fft

FS = 5e6
NINT = 1


# --------------------------------------------------------------------------- replica (A1)
def code_bpsk(bits, rep=2):
    """godual_ranging.m:62-65: int8 0/1 -> sample-and-hold x rep -> +/-1 (float64)."""
    c = np.repeat(np.asarray(bits, dtype=np.float64), rep)
    return 2.0 * c - 1.0


def code_qpsk(bits_interleaved, rep=2):
    """goqpsk.m:9-13: codec=(codei+j*codeq)-mean, sample-and-hold x rep (complex128)."""
    b = np.asarray(bits_interleaved, dtype=np.float64)
    codec = b[0::2] + 1j * b[1::2]
    codec = codec - np.mean(codec)
    return np.repeat(codec, rep)


def fcode_G(code):
    """godual_ranging.m:66 / goqpsk.m:16: conj(fft(code)) (non-conjugating transpose)."""
    return np.conj(_fft.fft(code))


def fcode_K(code):
    """claudio_aligned_code_lo_separate.m:96: fft(code), NOT conjugated."""
    return _fft.fft(code)


# --------------------------------------------------------------------------- frequency axes (A3)
def freq_axis_G(fs, n):
    """godual_ranging.m:73: linspace(-fs/2,fs/2,N) (inclusive; the +0.25 Hz quirk)."""
    return np.linspace(-fs / 2, fs / 2, int(n))


def freq_axis_K(fs, ls):
    """claudio_aligned_code_lo_separate.m:103: linspace(-fs/2,fs/2-fs/fs,fs*ls)."""
    return np.linspace(-fs / 2, fs / 2 - 1.0, int(fs * ls))


def window(freq, lo_hz, hi_hz):
    """k=find((freq<hi)&(freq>lo)) -- strict inequalities; returns 0-based indices."""
    return np.nonzero((freq < hi_hz) & (freq > lo_hz))[0]


def carrier_estimate(d, k, freq, power=2):
    """godual_ranging.m:14-15: argmax of fftshift(abs(fft(d.^p))) inside window k.

    Returns (s, df): s = 0-based index into the fftshift'ed spectrum, df = freq[s]/power.
    power=4 is the QPSK rule of Doc/07_Demodulation.md:60-63."""
    x = d * d
    if power == 4:
        x = x * x
    spec = np.fft.fftshift(np.abs(_fft.fft(x)))
    s = int(np.argmax(spec[k])) + int(k[0])
    return s, freq[s] / power


# --------------------------------------------------------------------------- helpers
def _pad3(spec):
    """godual_ranging.m:26-28,38-40: bins 0..N/2-1 stay, N/2..N-1 go to the top of 3N."""
    n = len(spec)
    out = np.zeros(3 * n, dtype=np.complex128)
    out[: n // 2] = spec[: n // 2]
    out[3 * n - (n - n // 2):] = spec[n // 2:]
    return out


def _var(x, ddof):
    m = np.mean(x)
    return float(np.sum(np.abs(x - m) ** 2) / (len(x) - ddof))


def parabola(am1, a0, ap1):
    """godual_ranging.m:33: vertex of the parabola through |x(-1)|,|x(0)|,|x(+1)|."""
    return (am1 - ap1) / (am1 + ap1 - 2.0 * a0) / 2.0


def fine_frequency(y, fs):
    """godual_ranging.py:26-28 (the 221228 addition, commented out in processing/Octave)."""
    a = np.polyfit(np.arange(1, fs // 3, 10) / fs,
                   np.convolve(np.angle(y[0:int(fs // 3):10]), np.ones(100) / 100)[49:-50], 1)
    return a[0] / 2 / np.pi


# --------------------------------------------------------------------------- variant G (A3,A5-A9)
def processing_G(d, k, freq, fcode, code, fs=FS, power=2, fine_freq=False, ddof=1, df=None):
    """[indice,correction,SNRr,SNRi,df,puissance,puissancecode,puissancenoise]=processing(d,k)
    godual_ranging.m:12-49.  `d` has its mean already removed by the caller (:80,94).
    fine_freq/ddof=0 reproduce godual_ranging.py.  `df` not None skips the estimator."""
    n = len(d)
    temps = np.arange(n) / fs
    out = {}
    if df is None:
        s, df = carrier_estimate(d, k, freq, power)
        out["s"] = s
    y = d * np.exp(-1j * 2 * np.pi * df * temps)
    if fine_freq:
        dfl = fine_frequency(y, fs)
        y = y * np.exp(-1j * 2 * np.pi * dfl * temps)
        df = df + dfl
    ffty = _fft.fft(y)
    prnmap = _fft.ifft(_pad3(ffty * fcode))
    a = np.abs(prnmap)
    p0 = int(np.argmax(a))
    m3 = 3 * n
    xval, xvalm1, xvalp1 = prnmap[p0], prnmap[(p0 - 1) % m3], prnmap[(p0 + 1) % m3]
    correction = parabola(abs(xvalm1), abs(xval), abs(xvalp1))
    yint = _fft.ifft(_pad3(ffty))
    codetmp = np.repeat(code, 3)
    # yincode=[yint(indice-1:end);yint(1:indice-2)].*codetmp  -> rotate left by indice-2 = p0-1
    # Complex (QPSK) replica: despread with conj(code).  The reference only ever does this with a real
    # code (goqpsk.m:39-45 stops at the peak), for which conj() changes nothing.
    yincode = np.roll(yint, -(p0 - 1)) * np.conj(codetmp)
    v = _var(yincode, ddof)
    mr, mi = np.mean(yincode.real), np.mean(yincode.imag)
    out.update(indice=p0 + 1, p0=p0, correction=float(correction), xval=complex(xval),
               xvalm1=complex(xvalm1), xvalp1=complex(xvalp1),
               SNRr=float(mr * mr / v), SNRi=float(mi * mi / v), df=float(df),
               puissance=_var(y, ddof), puissancecode=float(mr * mr + mi * mi),
               puissancenoise=v)
    return out


# --------------------------------------------------------------------------- C++ program (processing/CPP/main.cpp)
def hamming_sigpack(n):
    """SigPack 1.2.7 hamming(N) (processing/CPP/main.cpp:16,717): 0.54-0.46*cos(2*pi*i/(N-1)).  Third-party code,
    not under /root/reference: parity with SigPack itself is UNPINNED."""
    i = np.arange(int(n), dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * i / (n - 1.0))


def fcode_CPP(code):
    """GoRanging::fill_fcode, main.cpp:658-725: conj(fft(code)) times hamming over the BIN index (:717-719)."""
    return np.conj(_fft.fft(code)) * hamming_sigpack(len(code))


def processing_CPP(d, df, fcode_win, code, fs=FS):
    """GoRanging::_process_method, main.cpp:224-361, for one channel of one record.  `d` raw complex samples (the
    mean is removed here, :248), `df` the file-wide offset (_foffset1/2, :270-275).  FFTW's backward transform is
    unnormalised, so prnmap/yint are 3N times Octave's; indices are 0-based; var divides by n (:350).
    Returns the values the C++ pushes into its vectors (:309-357)."""
    n = len(d)
    m3 = 3 * n
    d = d - np.mean(d)
    y = d * np.exp(-1j * 2 * np.pi * df * (np.arange(n) / fs))
    ffty = _fft.fft(y)
    puissance = float(np.sum(np.abs(ffty) ** 2))                       # :295
    prnmap = _fft.ifft(_pad3(ffty * fcode_win)) * m3                    # :289-301, unnormalised
    p0 = int(np.argmax(np.abs(prnmap)))                                 # :305
    xval, xvalm1, xvalp1 = prnmap[p0], prnmap[(p0 - 1) % m3], prnmap[(p0 + 1) % m3]
    a0, am1, ap1 = abs(xval), abs(xvalm1), abs(xvalp1)
    corr = (am1 - ap1) / 2.0 / (am1 + ap1 - 2.0 * a0)                   # :309
    yint = _fft.ifft(_pad3(ffty)) * m3                                  # :319-327
    yincode = np.roll(yint, -(p0 - 1)) * np.repeat(code, 3)             # std::rotate by indice1-1, :332-338
    snr1 = np.mean(yincode)
    s = float(np.mean(np.abs(yincode - snr1) ** 2))                     # :342-350 (divides by n)
    snr_r, snr_i = snr1.real ** 2 / s, snr1.imag ** 2 / s
    return dict(p0=p0, correction=float(p0 + corr), xval=complex(xval), xvalm1=complex(xvalm1),
                xvalp1=complex(xvalp1), df=float(df), puissance=puissance,
                puissancecode=float(10 * np.log10(abs(snr1) ** 2)), SNR=float(10 * np.log10(snr_r + snr_i)))


def df_CPP(raw_file, nch_total, fs=FS, dec=25, foffset=0.0, remote=0):
    """GoRanging::df, main.cpp:363-450: one sample instant out of `dec` over the WHOLE file (no filter), derotated
    by `foffset`, mean removed, squared, one FFT per channel; channel 1 looks at |f| < 16 kHz of the +-fs/2/dec
    axis (kmin..kmax-1), channel 2 at every bin; df = freq[pos]/2 + foffset.  raw_file: int16 [I1 Q1 I2 Q2]..."""
    r = np.asarray(raw_file, dtype=np.float64)
    ninst = len(r) // 4
    L = (len(r) * 2) // (2 * 4 * dec)                                   # ftell / (sizeof(int16)*4*N), :377
    idx = np.arange(L) * dec
    idx = idx[idx < ninst]
    t = np.arange(len(idx)) * (dec / fs)
    lo = np.exp(-1j * 2 * np.pi * foffset * t)
    d1 = r[4 * idx] + 1j * r[4 * idx + 1]
    m1 = np.mean(d1)                                                    # _mean1 accumulates the RAW samples (:387)
    x1 = d1 * lo
    L = len(x1)
    freq = np.linspace(-fs / 2 / dec, fs / 2 / dec, L)
    frange = 8000.0
    kmin, kmax = 0, 0
    for i in range(L) if L < 4096 else ():                              # literal loop of :393-398 for small L
        if freq[i] < 2 * frange:
            kmax = i
        if freq[i] <= -2 * frange:
            kmin = i
    if L >= 4096:
        w = np.nonzero(freq < 2 * frange)[0]
        kmax = int(w[-1]) if len(w) else 0
        w = np.nonzero(freq <= -2 * frange)[0]
        kmin = int(w[-1]) if len(w) else 0
    x1 = x1 - m1                                                        # ... and is subtracted AFTER the derotation (:417-420)
    out = np.fft.fftshift(np.abs(_fft.fft(x1 * x1))) if L % 2 == 0 else None
    if out is None:                                                     # memcpy halves of :413-414 for odd L
        f = np.abs(_fft.fft(x1 * x1))
        h = L // 2
        out = np.concatenate([f[h:h + h], f[:h], np.zeros(L - 2 * h)])      # the resize() zero stays
    pos1 = int(np.argmax(out[kmin:kmax])) + kmin
    res = {1: freq[pos1] / 2.0 + foffset, "pos1": pos1}
    if remote == 0:
        d2 = r[4 * idx + 2] + 1j * r[4 * idx + 3]
        x2 = d2 * lo - np.mean(d2)                                      # _mean2: raw mean, :392,437-440
        f = np.abs(_fft.fft(x2 * x2))
        h = L // 2
        out2 = np.concatenate([f[h:h + h], f[:h], np.zeros(L - 2 * h)])
        res["pos2"] = int(np.argmax(out2))
        res[2] = freq[res["pos2"]] / 2.0 + foffset
    return res


# --------------------------------------------------------------------------- variant K (A5-A9)
def processing_K(d, df, fcode, code, fs=FS, ddof=1):
    """[xval,indice,correction,SNRr,SNRi,puissance,puissancecode,puissancenoise]=processing(d,k,df)
    claudio_aligned_code_lo_separate.m:24-77 with length(d)==length(code) (cm=1).
    `fcode` is fft(code) (un-conjugated, :96); product fcode.*conj(ffty) (:34)."""
    n = len(d)
    temps = np.arange(n) / fs
    y = d * np.exp(-1j * 2 * np.pi * df * temps)
    ffty = _fft.fft(y)
    prnmap = _fft.ifft(_pad3(fcode * np.conj(ffty)))
    yint = _fft.ifft(_pad3(ffty))
    codetmp = np.repeat(code, 3)
    m3 = 3 * n
    i0 = int(np.argmax(np.abs(prnmap)))
    i1 = i0 + 1                                   # Octave index
    xval = prnmap[i0]
    xvalm1 = prnmap[i0 - 1] if (i1 - 1) >= 1 else prnmap[m3 - 1]      # :46-50
    xvalp1 = prnmap[i0 + 1] if (i1 + 1) < m3 else prnmap[0]           # :51-55 (strict <)
    correction = (abs(xvalm1) - abs(xvalp1)) / (abs(xvalm1) + abs(xvalp1) - 2 * abs(xval)) / 2
    if i1 > 2:                                    # :65-69 rotate the CODE left by i-2
        yincode = np.roll(codetmp, -(i1 - 2)) * yint
    else:
        yincode = codetmp * yint
    v = _var(yincode, ddof)
    mr, mi = np.mean(yincode.real), np.mean(yincode.imag)
    return dict(indice=i1, p0=i0, correction=float(correction), xval=complex(xval),
                xvalm1=complex(xvalm1), xvalp1=complex(xvalp1),
                SNRr=float(mr * mr / v), SNRi=float(mi * mi / v), df=float(df),
                puissance=_var(y, ddof), puissancecode=float(mr * mr + mi * mi),
                puissancenoise=v)


# --------------------------------------------------------------------------- search_df (A4) / Doppler bank (A12)
def xcorr_metric_1x(d, df, fcode, fs=FS):
    """Body of search_df's loop, claudio_aligned_code_re_separate.m:33-41:
    derotate first N samples, prnmap=abs(ifft(fcode.*conj(fft(y)))), [s,b]=max, zero b-5:b+5
    (no wrap), var of the magnitude array; returns (b0 0-based, prnsig, snr=prnsig^2/var)."""
    n = len(fcode)
    temps = np.arange(n) / fs
    y = d[:n] * np.exp(-1j * 2 * np.pi * df * temps)
    prnmap = np.abs(_fft.ifft(fcode * np.conj(_fft.fft(y))))
    b = int(np.argmax(prnmap))
    prnsig = float(prnmap[b])
    lo, hi = b - 5, b + 5                          # Octave b-5:b+5 (1-based), errors if out of range
    if lo < 0 or hi > n - 1:
        raise IndexError("search_df: peak within 5 samples of the edge (Octave index error)")
    prnmap[lo:hi + 1] = 0
    prnvar = float(np.var(prnmap, ddof=1))
    return b, prnsig, prnsig * prnsig / prnvar


def search_df(d, k, df_threshold, freq, fcode, fs=FS, power=2):
    """k=search_df(d,k,df_threshold), claudio_aligned_code_re_separate.m:26-46.
    Returns kbon as a 1-BASED index into the shifted spectrum (0 = not found),
    plus the list of (candidate 1-based index, snr) for diagnostics."""
    x = d * d
    if power == 4:
        x = x * x
    d2 = np.fft.fftshift(np.abs(_fft.fft(x)))
    sel = d2[k]
    ktmp = np.nonzero(sel > np.median(sel) * df_threshold)[0] + int(k[0])   # 0-based
    kbon, tried = 0, []
    if 0 < len(ktmp) < 100:
        for kk in ktmp:
            dftmp = freq[kk] / power
            _, _, snr = xcorr_metric_1x(d, dftmp, fcode, fs)
            tried.append((int(kk) + 1, snr))
            if snr > 100:
                kbon = int(kk) + 1
    return kbon, tried


def doppler_bank(d, dfs, fcode, fs=FS):
    """Config 4: per-bin body of search_df over a uniform grid of offsets.
    Returns arrays (b0, prnsig, metric) per bin, with the +/-5 exclusion WRAPPED
    (a grid cannot rely on the Octave index error)."""
    n = len(fcode)
    temps = np.arange(n) / fs
    b0 = np.zeros(len(dfs), dtype=np.int64)
    sig = np.zeros(len(dfs))
    met = np.zeros(len(dfs))
    for i, df in enumerate(dfs):
        y = d[:n] * np.exp(-1j * 2 * np.pi * df * temps)
        prnmap = np.abs(_fft.ifft(fcode * np.conj(_fft.fft(y))))
        b = int(np.argmax(prnmap))
        s = float(prnmap[b])
        prnmap[(np.arange(b - 5, b + 6)) % n] = 0
        b0[i], sig[i], met[i] = b, s, s * s / float(np.var(prnmap, ddof=1))
    return b0, sig, met


def acquire_halving(d, fcode, fc_init, frange, fstep, min_step=1.0, fs=FS):
    """Step-halving acquisition of experiments/231001_DLL_PLL/rxcomplex.cpp:522-568 with the per-bin body of
    search_df (claudio_aligned_code_re_separate.m:33-37) as the bin metric: grid fc-frange..fc+frange accumulated as
    fcc += fstep, best bin replaced only by a strictly larger peak, then fstep /= 2, frange = fstep, until
    fstep < min_step.  PARITY UNPINNED by the reference (its cross_spectrum band-limits with experiment-specific
    filter edges; only the search loop is restated).  Returns (fc, b0, prnsig, bins evaluated)."""
    n = len(fcode)
    temps = np.arange(n) / fs
    fc, pk, b0, nb = fc_init, 0.0, 0, 0
    while True:
        fcc, fhigh = fc - frange, fc + frange
        grid = []
        while fcc <= fhigh:
            grid.append(fcc)
            fcc += fstep
        for f in grid:
            prnmap = np.abs(_fft.ifft(fcode * np.conj(_fft.fft(d[:n] * np.exp(-1j * 2 * np.pi * f * temps)))))
            b = int(np.argmax(prnmap))
            if prnmap[b] > pk:
                pk, fc, b0 = float(prnmap[b]), f, b
        nb += len(grid)
        fstep = fstep / 2.0
        frange = fstep
        if fstep < min_step:
            break
    return fc, b0, pk, nb


# --------------------------------------------------------------------------- IQ reader (A2)
def unpack_interleaved2(raw_int16):
    """godual_ranging.m:76-79: [I1 Q1 I2 Q2]... -> (d1, d2) complex128, means NOT removed."""
    r = np.asarray(raw_int16).astype(np.float64)
    return r[0::4] + 1j * r[1::4], r[2::4] + 1j * r[3::4]


def unpack_single(raw_int16):
    """claudio_aligned_code_lo_separate.m:121: [I Q]... -> complex128."""
    r = np.asarray(raw_int16).astype(np.float64)
    return r[0::2] + 1j * r[1::2]


def record_G(raw_record, k, freq, fcode, code, fs=FS, channels=(1, 2), **kw):
    """One iteration of godual_ranging.m:75-102 on one 4N-int16 record."""
    d1, d2 = unpack_interleaved2(raw_record)
    res = {}
    if 1 in channels:
        res[1] = processing_G(d1 - np.mean(d1), k, freq, fcode, code, fs, **kw)
    if 2 in channels:
        res[2] = processing_G(d2 - np.mean(d2), k, freq, fcode, code, fs, **kw)
    return res

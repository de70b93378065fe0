"""Synthetic PRN-modulated noisy IQ records.

Host-side test/bench tooling: builds int16 records in the two on-disk layouts of the
reference (processing/Python/b210_file.py:60-65 interleaved two-channel; acquisition/
rx_multi_samples.cpp:214-215 one channel per file).  Not on the hot path.
Signal recipe follows experiments/221127_SNR/simu_snr.m (code + AWGN at a chosen SNR).
"""
import numpy as np

SEED0 = 20261017


def modulate(code, delay, df, phase, snr_db, fs=5e6, std_lsb=3000.0, rng=None):
    """One channel: circularly delayed `code` (real +/-1 or complex QPSK replica, already at the
    sample rate) * carrier + complex AWGN; returns complex128 before quantisation.

    delay may be fractional (band-limited spectral shift).  snr_db is signal/noise power in the
    full band fs.  Total I std is scaled to ~std_lsb."""
    rng = rng or np.random.default_rng(SEED0)
    n = len(code)
    di = int(np.floor(delay))
    frac = float(delay) - di
    s = np.roll(np.asarray(code, dtype=np.complex128), di)
    if frac != 0.0:
        kk = np.fft.fftfreq(n, 1.0 / n)
        s = np.fft.ifft(np.fft.fft(s) * np.exp(-2j * np.pi * kk * frac / n))
        if np.isrealobj(code):
            s = s.real.astype(np.complex128)
    s = s / np.sqrt(np.mean(np.abs(s) ** 2))
    t = np.arange(n) / fs
    s = s * np.exp(1j * (2 * np.pi * df * t + phase))
    ps = 10.0 ** (snr_db / 10.0)
    noise = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.0)
    x = np.sqrt(ps) * s + noise
    # I std ~ std_lsb: total complex power = 1+ps -> per-component variance (1+ps)/2
    return x * (std_lsb / np.sqrt((1.0 + ps) / 2.0))


def quantize(x):
    """round + saturate to int16, returns (I, Q)."""
    i = np.clip(np.rint(x.real), -32768, 32767).astype(np.int16)
    q = np.clip(np.rint(x.imag), -32768, 32767).astype(np.int16)
    return i, q


def record_single(code, seed, delay=None, df=None, phase=None, snr_db=-5.0, fs=5e6, dc=(0.0, 0.0)):
    """C1: one channel [I Q] int16[2N].  Unspecified parameters are drawn from the seed."""
    rng = np.random.default_rng(seed)
    n = len(code)
    delay = rng.uniform(0, n) if delay is None else delay
    df = rng.uniform(-9000, 9000) if df is None else df
    phase = rng.uniform(-np.pi, np.pi) if phase is None else phase
    x = modulate(code, delay, df, phase, snr_db, fs, rng=rng) + (dc[0] + 1j * dc[1])
    i, q = quantize(x)
    out = np.empty(2 * n, dtype=np.int16)
    out[0::2], out[1::2] = i, q
    return out, dict(delay=delay, df=df, phase=phase, snr_db=snr_db)


def record_pair(code, seed, fs=5e6, ch1=None, ch2=None):
    """C2: two-channel interleaved [I1 Q1 I2 Q2] int16[4N]; ch1 = measurement (downlink),
    ch2 = loopback (SNR +0.6 dB, df 0, delay ~0.7253 s) unless overridden."""
    rng = np.random.default_rng(seed)
    n = len(code)
    p1 = dict(delay=rng.uniform(0.25, 0.99) * n, df=rng.uniform(-9000, 9000),
              phase=rng.uniform(-np.pi, np.pi), snr_db=-5.0)
    p2 = dict(delay=0.7253 * n + rng.uniform(-1, 1), df=0.0,
              phase=rng.uniform(-np.pi, np.pi), snr_db=0.6)
    p1.update(ch1 or {})
    p2.update(ch2 or {})
    x1 = modulate(code, fs=fs, rng=rng, **p1) + (11.3 - 7.9j)
    x2 = modulate(code, fs=fs, rng=rng, **p2) + (-5.2 + 3.4j)
    out = np.empty(4 * n, dtype=np.int16)
    out[0::4], out[1::4] = quantize(x1)
    out[2::4], out[3::4] = quantize(x2)
    return out, (p1, p2)


def stream_single(code, nsamples, seed, delay=0.0, df=0.0, phase=0.3, snr_db=-5.0, fs=5e6, dc=(0.0, 0.0)):
    """One-channel-per-file recording (rx_multi_samples.cpp:214-215): the code repeats back to back for
    nsamples sample instants, starting `delay` samples late; [I Q] int16[2*nsamples]."""
    rng = np.random.default_rng(seed)
    n = len(code)
    reps = int(np.ceil((nsamples + n) / n)) + 1
    s = np.tile(np.asarray(code, dtype=np.complex128), reps)
    di = int(np.floor(delay)) % n
    s = s[n - di:n - di + nsamples]
    t = np.arange(nsamples) / fs
    s = s * np.exp(1j * (2 * np.pi * df * t + phase))
    ps = 10.0 ** (snr_db / 10.0)
    noise = (rng.standard_normal(nsamples) + 1j * rng.standard_normal(nsamples)) / np.sqrt(2.0)
    x = (np.sqrt(ps) * s + noise) * (3000.0 / np.sqrt((1.0 + ps) / 2.0)) + (dc[0] + 1j * dc[1])
    i, q = quantize(x)
    out = np.empty(2 * nsamples, dtype=np.int16)
    out[0::2], out[1::2] = i, q
    return out

"""Per-second reducer of the K-variant result records: host arithmetic only.

Mirrors acquisition/go_1s.m: validity selection (:80-87), conversion of (indice1, correction1) to nanoseconds
(:88-90), and the `.1s` file -- one row per second, each value the 25-point linear fit of the 25 window delays
of that second evaluated at its 13th point (:251-266).  Two-way combination and plots stay in the scripts.
The C program processing/C/go_1s.c is the authoritative implementation of this loop; this Python mirror over the same
C ABI exists so that the parity tests can drive the loop step by step.
"""
import numpy as np

HEADER = "# MJD\t\tOPlocal\tOPremote\tLTFBlocal\tLTBBremote\n"          # go_1s.m:253 (typo kept: file format)


def select_valid(xval1):
    """k=find(abs(xval1)>max(abs(xval1))/2); first gap ends the run; drop the first 10 windows (275 ms)
    and the last one (go_1s.m:80-87).  Returns 0-based indices."""
    a = np.abs(np.asarray(xval1))
    k = np.nonzero(a > a.max() / 2)[0]
    gaps = np.nonzero(np.diff(k) > 1)[0]
    if len(gaps):
        return k[10:gaps[0] + 1]          # k(11:kk(1))
    return k[10:-1]                       # k(11:end-1)


def delays_ns(indice1, correction1, fs=5e6, nint=1):
    """(indice1 + correction1/(2N+1))/fs*1e9 (go_1s.m:88,90)."""
    return (np.asarray(indice1, dtype=np.float64) + np.asarray(correction1, dtype=np.float64) / (2 * nint + 1)) / fs * 1e9


def cut_at_sample_loss(series_ns, jump_ns=2.0):
    """kk=find(abs(diff(oplo))>2): keep the part before the first jump (go_1s.m:93-100)."""
    s = np.asarray(series_ns)
    kk = np.nonzero(np.abs(np.diff(s)) > jump_ns)[0]
    if len(kk) and kk[0] > 0:
        return s[:kk[0]]
    return s


def one_second_points(series_ns, per_second=25):
    """for k=1:25:length-25: [a,b]=polyfit([k-1:k+23]/25, s(k:k+24), 1); b.yf(13)  (go_1s.m:255-263)."""
    s = np.asarray(series_ns, dtype=np.float64)
    out = []
    for k in range(0, len(s) - per_second, per_second):
        x = np.arange(k, k + per_second) / per_second
        c = np.polyfit(x, s[k:k + per_second], 1)
        out.append(np.polyval(c, x[12]))
    return np.asarray(out)


def write_1s(path, mjd0, oplo, opre, ltlo, ltre, per_second=25):
    """The `<MJD>.1s` file (go_1s.m:251-266): header + "%f\\t%f\\t%f\\t%f\\t%f\\n" rows, MJD advancing by 1/86400."""
    n = min(len(oplo), len(opre), len(ltlo), len(ltre))
    cols = [one_second_points(np.asarray(s)[:n], per_second) for s in (oplo, opre, ltlo, ltre)]
    with open(path, "w") as f:
        f.write(HEADER)
        for i in range(len(cols[0])):
            f.write("%f\t%f\t%f\t%f\t%f\n" % (mjd0 + i / 86400.0, cols[0][i], cols[1][i], cols[2][i], cols[3][i]))
    return len(cols[0])

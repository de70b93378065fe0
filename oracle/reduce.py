"""ORACLE (test infrastructure): numpy restatement of the per-second reducer acquisition/go_1s.m for ONE session,
i.e. one quadruple of K-variant result files (OP local / OP remote / LTFB local / LTFB remote), statement by
statement.  Only tests/ may import it.  Directory scanning, dates from file names and plots are not restated."""
import numpy as np


def _select(xval1):
    """go_1s.m:80-87 (= :139-146): k=find(abs(xval1)>max(abs(xval1))/2); cut at the first gap; drop 10 + last."""
    a = np.abs(xval1)
    k = np.nonzero(a > a.max() / 2)[0]
    kk = np.nonzero(np.diff(k) > 1)[0]
    if len(kk):
        return k[10:kk[0] + 1]            # k(11:kk(1))
    return k[10:-1]                       # k(11:end-1)


def _ns(m, k, fs, nint):
    return (m["indice1"][k] + m["correction1"][k] / (2 * nint + 1)) / fs * 1e9      # :88,90


def _nanmedian(x):
    return np.nanmedian(x) if np.any(~np.isnan(x)) else np.nan


def session(oplo_m, opre_m, ltlo_m, ltre_m, fs=5e6, nint=1):
    """Returns dict(oplo, opre, ltlo, ltre [ns, equal lengths], res [ns, NaN = outlier], snrop, snrlt) or None
    when the OP loopback run is too short (:101)."""
    k = _select(oplo_m["xval1"])
    oplo = _ns(oplo_m, k, fs, nint)
    kk = np.nonzero(np.abs(np.diff(oplo)) > 2)[0]              # :93 sample loss at OP loopback
    if len(kk) and kk[0] + 1 > 1:
        oplo = oplo[:kk[0]]                                    # oplo(1:kk-1), kk 1-based = kk[0]+1
    if len(oplo) <= 102:                                       # :101
        return None
    x = opre_m["xval1"][k]                                     # :106
    kkk = np.nonzero(np.abs(x) > np.abs(x).max() / 2)[0]
    kkkk = np.nonzero(np.diff(kkk) > 1)[0]
    if len(kkkk):                                              # :110-117
        k = k[:kkkk[0] + 1]
        if kkkk[0] + 1 < len(oplo):
            oplo = oplo[:kkkk[0] + 1]
    opre = _ns(opre_m, k, fs, nint)[:len(oplo)]                # :119-122
    snrop = float(np.median(10 * np.log10(np.abs(opre_m["SNR1r"][k] + opre_m["SNR1i"][k]) * fs)))   # :123
    k = _select(ltlo_m["xval1"])                               # :139-146
    ltlo = _ns(ltlo_m, k, fs, nint)
    x = ltre_m["xval1"][k]                                     # :157
    kkk = np.nonzero(np.abs(x) > np.abs(x).max() / 2)[0]
    kkkk = np.nonzero(np.diff(kkk) > 1)[0]
    if len(kkkk):                                              # :160-164
        k = k[:kkkk[0] + 1]
        ltlo = ltlo[:kkkk[0] + 1]
    if len(kkk) < len(ltlo):                                   # :165-169
        k = k[:kkk[-1] + 1]
        ltlo = ltlo[:kkk[-1] + 1]
    ltre = _ns(ltre_m, k, fs, nint)                            # :170-172
    snrlt = float(np.median(10 * np.log10(np.abs(ltre_m["SNR1r"][k] + ltre_m["SNR1i"][k]) * fs)))
    if len(oplo) > len(ltlo):                                  # :175-181
        oplo, opre = oplo[:len(ltlo)], opre[:len(ltlo)]
    else:
        ltlo, ltre = ltlo[:len(oplo)], ltre[:len(oplo)]
    n = min(len(oplo), len(opre), len(ltlo), len(ltre))
    res = 0.5 * ((opre[:n] - oplo[:n]) - (ltre[:n] - ltlo[:n]))      # :191
    res[np.abs(res - np.median(res)) > 5] = np.nan                 # :192-193
    step = 200.0 / (2 * nint + 1)
    ki = res > _nanmedian(res) + 10                                # :208-211, as written (pkg nan: NaN-skipping median)
    res[ki] -= step
    ki = res > _nanmedian(res) - 10
    res[ki] += step
    return dict(oplo=oplo, opre=opre, ltlo=ltlo, ltre=ltre, res=res, snrop=snrop, snrlt=snrlt)


def one_second_rows(s, mjd0, per=25):
    """go_1s.m:251-266: for k=1:25:length(opre)-25, 25-point linear fits evaluated at their 13th point."""
    rows = []
    cpt = 0
    for k in range(0, len(s["opre"]) - per, per):
        x = np.arange(k, k + per) / per
        vals = []
        for name in ("oplo", "opre", "ltlo", "ltre"):
            c = np.polyfit(x, s[name][k:k + per], 1)
            vals.append(np.polyval(c, x[12]))
        rows.append((mjd0 + cpt / 86400.0, *vals))
        cpt += 1
    return rows


# --------------------------------------------------------------------------------------------------------------
# ORACLE (test infrastructure) for processing/C/gofinal.c: processing/Octave/gofinal_{op,ltfb}.m and tw.m restated.
def _station(mat, cflavour, two):
    """gofinal_ltfb.m:35-47 / gofinal_op.m:30-43: solutions, SNR (dB) and df of one result file."""
    g = lambda k: np.asarray(mat[k], dtype=float).ravel()
    corr = lambda c: g(f"correction{c}_a") if f"correction{c}_a" in mat else g(f"correction{c}")
    out = {}
    for c in ((1, 2) if two else (1,)):
        if cflavour:
            out[f"sol{c}"], out[f"snr{c}"] = corr(c), g(f"SNR{c}")
        else:
            out[f"sol{c}"], out[f"snr{c}"] = g(f"indice{c}") + corr(c), 10 * np.log10(g(f"SNR{c}r") + g(f"SNR{c}i"))
        out[f"df{c}"] = g(f"df{c}")
    out["codelen"] = len(g("code")) if "code" in mat else 5000000
    return out


def gofinal_rows(local, remote, end_epoch, ltfb, local_c=False, remote_c=False, fs=5e6):
    """Rows (epoch, 9 values) written by gofinal_op.m:70-78 (every record) / gofinal_ltfb.m:84-90 (selected)."""
    lo, re = _station(local, local_c, True), _station(remote, remote_c, False)
    n = len(lo["sol1"])
    # gofinal_ltfb.m:48 counts the LOCAL file's records; gofinal_op.m:36-37 evaluates length(indice1) while the
    # REMOTE file is the one loaded
    dateini = end_epoch - (n if ltfb else len(re["sol1"]))
    ind = np.nonzero(lo["snr1"] > lo["snr1"].max() - 10)[0]
    res = (lo["sol1"][ind] - lo["sol2"][ind]) / fs / 3
    x = np.arange(1, len(res) + 1, dtype=float)
    fit = lambda r, xx: r - np.polyval(np.polyfit(xx - xx.mean(), r - r.mean(), 2), xx - xx.mean()) - r.mean()
    r = fit(res, x)
    if np.std(r, ddof=1) > 1e-6 and len(res) >= 150:
        i1 = np.nonzero(np.abs(res - res[29:40].mean()) < 1e-5)[0]
        i2 = np.nonzero(np.abs(res - res[139:150].mean()) < 1e-5)[0]
        if len(i1) and len(i2):
            ind = i1 if len(i1) > len(i2) else i2            # indices into the filtered list (the script's quirk)
            res = res[ind]
            r = fit(res, np.arange(1, len(res) + 1, dtype=float))
    stats = (float(np.std(r * 1e9, ddof=1)), float(res.mean() + (1 if res.mean() < 0 else 0)))
    step = (lo["codelen"] / 2) / 2.5e6
    rows = []
    for p in (ind if ltfb else range(n)):
        rows.append((int(np.floor(dateini + p * step)), lo["sol1"][p] / 3 / fs, lo["df1"][p], lo["snr1"][p],
                     lo["sol2"][p] / 3 / fs, lo["df2"][p], lo["snr2"][p], re["sol1"][p] / 3 / fs, re["df1"][p], re["snr1"][p]))
    return rows, stats


def two_way(op, lt):
    """tw.m:16-50 on the two numeric tables (15 columns): returns (difference, std_ns, mean_ns, n) or (difference,)."""
    stamp = lambda t: ((((t[0, 1] * 31 + t[0, 2]) * 24) + t[0, 3] * 60) + t[0, 4] * 60) + t[0, 5]
    a, b = stamp(op), stamp(lt)
    d = int(abs(a - b))
    if a > b:
        lt = lt[d:]
    else:
        op = op[d:]
    m = min(len(op), len(lt))
    op, lt = op[:m], lt[:m]
    if d >= 180 or m == 0:
        return (d,)
    for first in ("lt", "op"):
        t = lt if first == "lt" else op
        k = np.nonzero(t[:, 8] > t[:, 8].max() - 10)[0]
        k = np.arange(k[0] + 1, k[-1])
        lt, op = lt[k], op[k]
    x, y = lt[:, 9] - lt[:, 12], op[:, 9] - op[:, 12]
    x = x + (1 if x.mean() < 0 else 0)
    y = y + (1 if y.mean() < 0 else 0)
    res = 0.5 * (x - y)
    res = res[np.abs(res) < 1e-6] * 1e9
    return d, float(np.std(res, ddof=1)), float(res.mean()), len(res)

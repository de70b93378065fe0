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

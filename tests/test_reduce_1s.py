"""Host-side `.1s` reducer (acquisition/go_1s.m:80-101,251-266)."""
import numpy as np

from amaranth_twstft_b200 import reduce_1s as r1


def test_selection_and_ns_conversion():
    x = np.r_[np.full(3, 0.1), np.full(40, 1.0), np.full(2, 0.1), np.full(5, 1.0)]
    k = r1.select_valid(x)
    assert k[0] == 13 and k[-1] == 42                     # 11th strong window .. last before the gap
    k2 = r1.select_valid(np.r_[np.full(3, 0.1), np.full(40, 1.0)])
    assert k2[0] == 13 and k2[-1] == 41                   # k(11:end-1)
    assert np.allclose(r1.delays_ns([21, 22], [0.3, -0.6]), [(21 + 0.1) / 5e6 * 1e9, (22 - 0.2) / 5e6 * 1e9])
    assert len(r1.cut_at_sample_loss(np.r_[np.zeros(30), 5 * np.ones(30)])) == 29


def test_one_second_file(tmp_path):
    rng = np.random.default_rng(0)
    t = np.arange(110) / 25.0
    series = [1000 + 3 * t + 0.01 * rng.standard_normal(110) for _ in range(4)]
    pts = r1.one_second_points(series[0])
    assert len(pts) == 4                                   # k = 1,26,51,76 (k <= length-25)
    for i, p in enumerate(pts):                            # least squares on 25 points, evaluated at the 13th
        k = 25 * i
        a = np.vstack([t[k:k + 25], np.ones(25)]).T
        slope, icpt = np.linalg.lstsq(a, series[0][k:k + 25], rcond=None)[0]
        assert abs(p - (slope * t[k + 12] + icpt)) < 1e-9
    n = r1.write_1s(tmp_path / "60000.1s", 60000.0, *series)
    lines = open(tmp_path / "60000.1s").read().splitlines()
    assert n == 4 and lines[0] == "# MJD\t\tOPlocal\tOPremote\tLTFBlocal\tLTBBremote" and len(lines) == 5
    f = lines[2].split("\t")
    assert len(f) == 5 and f[0] == "%f" % (60000.0 + 1 / 86400.0) and f[1] == "%f" % pts[1]

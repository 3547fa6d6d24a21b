"""Pins the oracle's restatement of un-vendored third-party arithmetic against scipy/numpy and
the committed golden fixtures.  CPU only."""
import os

import numpy as np
import pytest
from scipy import interpolate, special

from conftest import GOLDEN
from oracle import torj_oracle as O


def test_gausshermite_matches_numpy():
    for n in (8, 44, 134, 178):
        x, w = O.gausshermite(n)
        xr, wr = np.polynomial.hermite.hermgauss(n)
        assert np.abs(x - xr).max() < 1e-13
        mid = slice(n // 2 - 3, n // 2 + 3)
        assert np.abs((w - wr) / wr)[mid].max() < 1e-13
        assert abs(w.sum() - np.sqrt(np.pi)) < 1e-14


def test_launch_known_answers():
    """reference test/tests/test_launch_weights.jl:15-56: N_rings=21, min_az=11, un-normalised sum within 1 % of 1."""
    g = np.load(os.path.join(GOLDEN, "launch_known.npz"))
    p, d, w = O.launch_peripheral_rays([0, 0, 0.0], [0, 0, 1.0], 0.0174, 1 / 3.99, 92.5e9, N_rings=21, min_azimuthal_points=11,
                                       normalize_weight_sum=False)
    assert len(w) == int(g["n_21_11"]) == 5165
    assert abs(w.sum() - float(g["sumw_21_11"])) < 1e-12
    assert abs(w.sum() - 1.0) < 0.01 and abs(w.sum() - 1.009424) < 1e-6
    for nr, ma, n in ((3, 5, 46), (7, 20, 1025), (66, 14, 65543), (88, 120, 999996)):
        assert int(g[f"n_{nr}_{ma}"]) == n
    p, d, w = O.launch_peripheral_rays([2.5, 0, 0.4], [-0.8660254, 0.0, -0.5], 0.0174, 1 / 3.99, 95e9)
    assert len(w) == 46 and abs(w.sum() - 1.0) < 1e-14
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-14)
    with pytest.raises(ValueError):
        O.launch_peripheral_rays([0, 0, 0.0], [0, 0, 1.0], 0.0174, 1 / 3.99, 92.5e9, N_rings=1)


def test_bessel_matches_scipy_and_golden():
    g = np.load(os.path.join(GOLDEN, "bessel_gl.npz"))
    for n in range(1, 5):
        mine = np.array([O.besselj(n, z) for z in g["z"]])
        assert np.abs(mine - g["J"][n - 1]).max() < 2e-15
        assert np.abs(mine - special.jv(n, g["z"])).max() < 2e-15


def test_fitpack_matches_scipy_golden():
    """Dierckx.Spline1D / roots / integrate (reference src/plasma.jl:101-134) == scipy splrep/sproot/splint."""
    g = np.load(os.path.join(GOLDEN, "fitpack_case.npz"))
    sp = O.FitpackSpline(g["x"], g["y"])
    t, c = sp.tc()
    assert np.array_equal(t, g["t"])
    assert np.abs(c - g["c"]).max() < 5e-14
    for lv, want in zip(g["levels"], g["roots"]):
        want = want[~np.isnan(want)]
        got = sp.roots(lv, 64)
        assert len(got) == len(want)
        assert np.abs(got - want).max() < 1e-11
    for (a, b), want in zip(g["pairs"], g["ints"]):
        assert abs(sp.integral(a, b) - want) < 1e-13
    # maxn cap of Dierckx.roots (default 8)
    assert len(sp.roots(0.0, 8)) <= 8


def test_fitpack_random_against_live_scipy():
    rng = np.random.default_rng(7)
    for trial in range(5):
        x = np.cumsum(rng.uniform(1e-4, 1e-3, 400))
        y = np.sin(40 * x) + 0.3 * x
        sp = O.FitpackSpline(x, y)
        tck = interpolate.splrep(x, y, s=0, k=3)
        xs = rng.uniform(x[0], x[-1], 50)
        assert np.abs(sp(xs) - interpolate.splev(xs, tck)).max() < 1e-12
        r = sp.roots(0.05, 64)
        rr = interpolate.sproot((tck[0], tck[1] - 0.05 * (np.arange(len(tck[1])) < len(x)), 3), mest=64)
        assert len(r) == len(rr) and np.abs(r - rr).max() < 1e-12
        assert abs(sp.integral(x[3], x[-7]) - interpolate.splint(x[3], x[-7], tck)) < 1e-14


def test_bspline_is_natural_cubic_spline(oracle_small, arrays_small):
    """Interpolations.jl BSpline(Cubic(Line(OnGrid()))) == natural cubic spline."""
    R, Z = arrays_small["R_coords"], arrays_small["Z_coords"]
    j = 40
    for name, key in (("Bphi", "Bphi_data"), ("BZ", "Bz_data"), ("psi", "psi_norm_data")):
        cs = interpolate.CubicSpline(R, arrays_small[key][:, j], bc_type="natural")
        Rq = np.linspace(R[0], R[-1], 301)
        v, dR, dZ = oracle_small.spline(name, Rq, np.full_like(Rq, Z[j]))
        assert np.abs(v - cs(Rq)).max() < 1e-13 * max(1.0, np.abs(v).max())
        assert np.abs(dR - cs(Rq, 1)).max() < 1e-11 * max(1.0, np.abs(dR).max())
    i = 22
    cs = interpolate.CubicSpline(Z, arrays_small["Br_data"][i, :], bc_type="natural")
    Zq = np.linspace(Z[0], Z[-1], 301)
    v, dR, dZ = oracle_small.spline("BR", np.full_like(Zq, R[i]), Zq)
    assert np.abs(v - cs(Zq)).max() < 1e-14 and np.abs(dZ - cs(Zq, 1)).max() < 1e-12


def test_line_extrapolation(oracle_small, arrays_small):
    R, Z = arrays_small["R_coords"], arrays_small["Z_coords"]
    v0, dR0, dZ0 = oracle_small.spline("psi", [R[-1]], [0.3])
    v1, dR1, dZ1 = oracle_small.spline("psi", [R[-1] + 0.25], [0.3])
    assert abs(v1[0] - (v0[0] + 0.25 * dR0[0])) < 1e-13 and abs(dR1[0] - dR0[0]) < 1e-14
    v2, _, _ = oracle_small.spline("psi", [R[0] - 0.1], [Z[0] - 0.2])
    vc, dRc, dZc = oracle_small.spline("psi", [R[0]], [Z[0]])
    assert abs(v2[0] - (vc[0] - 0.1 * dRc[0] - 0.2 * dZc[0])) < 1e-12


def _rk_solve(a, b, h, n):
    S = len(b)
    y, t = 1.0, 0.0
    f = lambda t, y: np.cos(t) * y + 0.1 * np.sin(3 * t)
    c = a.sum(1)
    for _ in range(n):
        k = np.zeros(S)
        for i in range(S):
            k[i] = f(t + c[i] * h, y + h * (a[i, :i] @ k[:i]))
        y += h * (b @ k)
        t += h
    return y


@pytest.mark.parametrize("scheme,order", [(0, 5), (1, 3)])
def test_rk_tableau_order(scheme, order):
    """Tsit5 is 5th order, OwrenZen3 3rd; the embedded estimate is one order lower."""
    a, bt = O.tableau(scheme)
    b = a[-1]                       # FSAL: last stage row = weights
    assert abs(b.sum() - 1.0) < 1e-14 and abs(bt.sum()) < 1e-14
    errs = []
    ref = _rk_solve(a, b, 1.0 / 2048, 2048)
    for n in (8, 16, 32):
        errs.append(abs(_rk_solve(a, b, 1.0 / n, n) - ref))
    rate = np.log2(errs[0] / errs[1]), np.log2(errs[1] / errs[2])
    assert abs(rate[1] - order) < 0.35, rate
    # error estimator h*sum(btilde k) scales like h^order for one step
    est = []
    f = lambda t, y: np.cos(t) * y + 0.1 * np.sin(3 * t)
    c = a.sum(1)
    for h in (0.1, 0.05):
        k = np.zeros(len(b))
        for i in range(len(b)):
            k[i] = f(c[i] * h, 1.0 + h * (a[i, :i] @ k[:i]))
        est.append(abs(h * (bt @ k)))
    assert abs(np.log2(est[0] / est[1]) - order) < 0.5

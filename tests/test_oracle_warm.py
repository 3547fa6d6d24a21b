"""The oracle's restatement of reference src/general_absorption.jl (warm-plasma α), checked WITHOUT the CUDA path.

The file is dead code in the reference (never include()d) and has no tests or golden vectors, so parity against Julia is
unpinned. What can be pinned here is pinned independently:
  * every special function against scipy (expi, gammaln, spherical_in, wofz);
  * the 501-node Hermitian quadrature against adaptive quadrature of the same integrand;
  * the two branches of antihermitian (closed-form recursion / spherical-Bessel series) against each other where they meet;
  * the weakly relativistic tensor (iwarm = 1: fsup, zetac, dieltens_maxw_wr) against the fully relativistic one at low T_e;
  * the cold limit of N_perp against the Appleton-Hartree index of src/dispersion.jl;
  * α itself, with the build-defined wiring v_g_perp = 1/|dΛ/dN|, against the independent Albajar model of
    src/absorption.jl across the second-harmonic layer (same physics, different derivation: agreement to ~10 %).
tests/golden/warm_alpha.npz then freezes the oracle's own numbers (regression only).
"""
import os

import numpy as np
import pytest
import scipy.integrate as si
import scipy.special as sp

from conftest import GOLDEN
from oracle import torj_oracle as O

C, E, ME = 2.99792458e8, 1.602176634e-19, 9.1093837015e-31


def mu_of(te_ev):
    return ME * C * C / (te_ev * E)


def test_expei_against_scipy_expi():
    xs = np.concatenate([-np.logspace(-4, 2.6, 150), np.logspace(-4, 2.6, 150),
                         [-1.0, -4.0, -4.0001, 5.9999, 6.0, 11.9999, 12.0, 24.0, 24.0001, 0.33, 0.41]])
    for x in xs:
        ref = np.exp(-x) * sp.expi(x)
        if abs(x - 0.37250741078136663) < 0.03:   # zero of Ei: compare absolutely
            assert abs(O.expei(x) - ref) < 1e-15
        else:
            assert abs(O.expei(x) - ref) <= 2e-14 * abs(ref), x   # exp(-x) expi(x) carries its own roundings
    assert O.expei(0.0) == -1.79e308           # reference src/general_absorption.jl:144-145


def test_gammln_fact_against_scipy():
    for x in np.linspace(0.5, 20.0, 79):
        assert abs(O.gammln(x) - sp.gammaln(x)) < 5e-10   # the reference's 6-term Lanczos formula is this accurate, no more
    assert [O.fact(k) for k in (-1, 0, 1, 5, 15)] == [0.0, 1.0, 1.0, 120.0, 1307674368000.0]
    assert O.fact(20) == pytest.approx(sp.factorial(20), rel=1e-15)


def test_ssbi_is_the_modified_spherical_bessel_series():
    """sum_k (z^2/4)^k / (k! Gamma(m+k+3/2)) = (2/z)^m (2/sqrt(pi)) i_m(z): NOT sphericalbesselj, which is what the
    reference's (dropped) assertion at :314-316 compares with."""
    for z in (0.05, 0.7, 2.0, 4.9):
        for n in (1, 2, 4):
            v = O.ssbi(z, n, 5)
            assert len(v) == 7
            for k, m in enumerate(range(n, 8)):
                ref = (2.0 / z) ** m * 2.0 / np.sqrt(np.pi) * sp.spherical_in(m, z)
                assert abs(v[k] - ref) <= 2e-9 * ref      # series stops at 1e-10, Gamma from the Lanczos gammln
                if z > 1.0:
                    assert abs(v[k] - (2.0 / z) ** m * 2.0 / np.sqrt(np.pi) * sp.spherical_jn(m, z)) > 1e-3 * ref


def test_zetac_against_faddeeva():
    """Z(z) = i sqrt(pi) w(z), upper half plane (the only one fsup reaches: yp, ym, y0 >= 0, :489-503)."""
    worst = 0.0
    for xr in np.linspace(-9.0, 9.0, 73):
        for yi in (0.0, 1e-3, 0.2, 1.0, 3.0, 7.0):
            ref = 1j * np.sqrt(np.pi) * sp.wofz(xr + 1j * yi)
            worst = max(worst, abs(O.zetac(xr, yi) - ref) / abs(ref))
    assert worst < 5e-14


def test_larmornumber():
    assert O.larmornumber(0.5, 0.0, mu_of(2e3)) == 3     # 2Y = 1: resonant; 3Y - 1 = 0.5 -> mu 0.5 = 128 > 15
    assert O.larmornumber(0.5, 0.0, mu_of(25e3)) == 4    # mu = 20.4: 3Y-1 -> 10.2 < 15, 4Y-1 -> 20.4
    assert O.larmornumber(1.2, 0.3, mu_of(5e3)) == 2     # first harmonic: mu (gamma_1 - 1) = 8 < 15, so one more
    assert O.larmornumber(0.3, 0.2, mu_of(10e3)) == 5


@pytest.mark.parametrize("yg,anpl,te", [(0.51, 0.12, 5e3), (0.34, -0.3, 15e3), (0.49, 0.0, 25e3)])
def test_hermitian_quadrature_against_adaptive_integration(yg, anpl, te):
    """rr(n, k, m) of :669-710 is sum_i exp(-t_i^2) dt g(t_i) over 501 nodes in [-5, 5]; the integrand is smooth apart from
    the integrable log singularity of expei at zm = 0, which costs the equidistant rule its spectral accuracy."""
    amu = mu_of(te)
    rr = O.hermitian(yg, anpl, amu, 3)
    cmxw = 1.0 + 15.0 / (8.0 * amu) + 105.0 / (128.0 * amu * amu)
    cr = -amu * amu / (np.sqrt(np.pi) * cmxw)
    bth2 = 2.0 / amu

    def integrand(t, n, k, m):
        rxt = np.sqrt(1.0 + t * t / (2.0 * amu)); x = t * rxt
        upl = np.sqrt(bth2) * x; gx = 1.0 + t * t / amu
        gr = anpl * upl + n * yg; zm = -amu * (gx - gr); s = amu * (gx + gr)
        fe = np.exp(-zm) * sp.expi(zm)
        base = cr * np.exp(-t * t) * gx / rxt
        if m == 0:
            return -base * fe * upl ** 2
        ffe = {1: (1.0 + s * (1.0 - zm * fe)) / amu ** 2,
               2: (6.0 - 2.0 * zm + 4.0 * s + s * s * (1.0 + zm - zm * zm * fe)) / amu ** 4,
               3: (18.0 * s * (s + 4.0 - zm) + 6.0 * (20.0 - 8.0 * zm + zm * zm) + s ** 3 * (2.0 + zm + zm * zm - zm ** 3 * fe)) / amu ** 6}[m]
        return base * ffe * upl ** k

    for n, k, m in [(0, 2, 0), (0, 0, 1), (-1, 0, 1), (1, 1, 1), (-2, 2, 2), (2, 0, 2), (3, 2, 3), (-3, 1, 3), (0, 1, 3)]:
        ref = si.quad(integrand, -5.0, 5.0, args=(n, k, m), limit=400, epsabs=0, epsrel=1e-11)[0]
        got = rr[n + 3, k, m]
        assert abs(got - ref) <= 2e-4 * abs(ref) + 1e-14, (n, k, m, got, ref)
    assert np.all(rr[:, :, 0].ravel()[np.arange(21) != 3 * 3 + 2] == 0.0)    # only rr(0,2,0) is filled at m = 0
    assert np.all(O.hermitian(yg, anpl, amu, 5)[:, :, 4:] == 0.0)             # llm = min(3, lrm): orders 4, 5 stay zero


def test_antihermitian_branches_agree_where_they_meet():
    """|aa| = mu |N_par| du = 5 separates the closed-form recursion (:974-1015) from the series (:1022-1037); the two are
    different evaluations of the same integrals, so they must join continuously."""
    yg, anpl = 0.52, 0.25
    dnl = 1.0 - anpl * anpl
    for n in (2, 3):
        du = np.sqrt((n * yg) ** 2 - dnl) / dnl
        amu0 = 5.0 / (anpl * du)
        lo, hi = O.antihermitian(yg, anpl, amu0 * (1 - 1e-6), 4), O.antihermitian(yg, anpl, amu0 * (1 + 1e-6), 4)
        blk_lo, blk_hi = lo[n - 1, :, n - 1:], hi[n - 1, :, n - 1:]
        assert np.all(blk_lo != 0.0)
        assert np.max(np.abs(blk_lo - blk_hi) / np.abs(blk_lo)) < 2e-4      # 1e-6 in mu moves exp(-mu ...) by ~1e-4
    assert np.all(O.antihermitian(0.3, 0.1, 100.0, 3)[:2] == 0.0)           # (n Y)^2 < 1 - N_par^2: no resonance, ri = 0


def test_weakly_and_fully_relativistic_tensors_agree_at_low_temperature():
    X, Y, npl = 0.25, 0.515, 0.2
    for te in (500.0, 2000.0):
        amu = mu_of(te)
        lrm = min(5, O.larmornumber(Y, npl, amu))
        ncold = 0.8
        fr = O.warmdisp(X, Y, npl, amu, ncold, 1, 3, lrm)
        wr = O.warmdisp(X, Y, npl, amu, ncold, 1, 1, lrm)
        assert fr["ierr"] == wr["ierr"] == 0
        assert abs(fr["anpr"].real - wr["anpr"].real) < 2e-3 * fr["anpr"].real
        assert abs(fr["anpr"].imag - wr["anpr"].imag) < (0.12 if te > 1000 else 0.05) * abs(fr["anpr"].imag)
        for k in ("ex", "ey", "ez"):
            assert abs(abs(fr[k]) - abs(wr[k])) < 2e-2


def test_cold_limit_is_appleton_hartree():
    """T_e -> 0 away from the resonances: N_perp^2 + N_par^2 -> Ns^2 of reference src/dispersion.jl:29-32."""
    X, Y, npl, te = 0.3, 0.37, 0.25, 50.0
    for sox in (1, -1):
        d = (1 - npl ** 2) ** 2 + 4 * npl ** 2 * (1 - X) / Y ** 2
        ns2 = 1 - X + (1 + sox * np.sqrt(d) + npl ** 2) / (2 * (-1 + X + Y ** 2)) * X * Y ** 2
        w = O.warmdisp(X, Y, npl, mu_of(te), np.sqrt(ns2 - npl ** 2), sox, 3, min(5, O.larmornumber(Y, npl, mu_of(te))))
        assert w["ierr"] == 0
        assert abs(w["anpr"].real ** 2 + npl ** 2 - ns2) < 2e-3 * ns2
        assert abs(w["anpr"].imag) < 1e-12


def test_alpha_is_zero_without_resonant_electrons():
    """Hermitian tensor (ri = 0) and a real start value: every coefficient of the biquadratic is real, so Im N_perp^2 = 0."""
    o = O.warm_alpha(2 * np.pi * 95e9, 0.2, 0.15, 0.85, 1.45, 300.0, 0.5, 1)   # lrm = 5 and 5 Y < sqrt(1 - N_par^2)
    assert o["lrm"] == 5 and o["ierr"] == 0 and o["alpha"] == 0.0 and o["N_warm"] > 0.5


def test_alpha_agrees_with_the_albajar_model_across_the_layer(oracle_small, gl24):
    """Independent derivations of the same damping (Albajar: src/absorption.jl; warm tensor: src/general_absorption.jl)
    at the same points of the 95 GHz second-harmonic layer of the Solov'ev equilibrium. Pins the build-defined wiring
    v_g_perp = 1/|dΛ/dN| (a factor 2 or a missing 1/N would show at once)."""
    f = 95e9
    om = 2 * np.pi * f
    for tor in (0.0, 0.15, 0.35):
        d = np.array([-np.cos(tor), np.sin(tor), 0.0])
        ratios = []
        for R in (1.90, 1.93, 1.96):   # up to the peak; on the steep cold side the two models differ by 30 %
            x = np.array([R, 0.0, 0.05])
            lo, hi = 0.1, 1.2
            for _ in range(70):   # |N| from the cold dispersion relation Λ = 0
                mid = 0.5 * (lo + hi)
                if oracle_small.eval_plasma(x, mid * d, om, 1)["Lambda"] > 0: hi = mid
                else: lo = mid
            u = np.concatenate([x, 0.5 * (lo + hi) * d, [1.0]])
            a_alb = -oracle_small.rhs(u, f, 1, gl24, absorption_model=0)[6]
            a_warm = -oracle_small.rhs(u, f, 1, gl24, absorption_model=1)[6]
            assert a_alb > 1.0 and a_warm > 1.0
            ratios.append(a_warm / a_alb)
        assert 0.9 < min(ratios) and max(ratios) < 1.25, ratios


def test_golden_regression():
    g = np.load(os.path.join(GOLDEN, "warm_alpha.npz"))
    for row, ref in zip(g["inputs"], g["outputs"]):
        o = O.warm_alpha(*row[:7], int(row[7]))
        got = np.array([o["N_warm"], o["alpha"], o["lrm"], o["ierr"], o["iterations"]])
        assert np.allclose(got, ref, rtol=1e-10, atol=1e-300), (row, got, ref)

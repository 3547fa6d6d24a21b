// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Fully relativistic EC absorption coefficient (Albajar), harmonics 2..3, restating
//   reference src/absorption.jl:1-7     abs_Al_init            (Gauss-Legendre nodes are passed in)
//   reference src/absorption.jl:10-64   abs_Al_N_with_pol_vec
//   reference src/absorption.jl:132-168 abs_Al_pol_fact
//   reference src/absorption.jl:170-189 abs_Al_integral_nume_fast
//   reference src/absorption.jl:191-226 abs_Albajar_fast
//   reference src/absorption.jl:228-235 α_approx
// Bessel functions: the reference calls SpecialFunctions.besselj (AMOS, un-vendored, compat 2.5.1);
// the oracle uses libm jn(), cross-checked against scipy.special.jv in tests/.
#pragma once
#include <cmath>
#include <complex>
#include <vector>
#include "general_absorption.hpp"
#include "plasma.hpp"

namespace torj_oracle {

struct AbsQuad {  // reference globals _int_absz/_int_weights (src/constants.jl:7-8)
    std::vector<double> t, w;
};

struct AbsCounters {
    long n_alpha = 0;  // calls of abs_Albajar_fast that passed the Te gate
    long n_harm = 0;   // harmonic integrals evaluated
    long n_prune = 0;  // harmonic integrals skipped by the alpha_floor bound (alpha_floor > 0 only)
    long n_askip = 0;  // inner Runge-Kutta stages that took alpha = 0 (alpha_floor > 0 only)
};

typedef std::complex<double> cplx;

// reference src/absorption.jl:10-64
inline double abs_Al_N_with_pol_vec(double X, double Y, double cos_theta, double sin_theta, int mode, cplx e[3]) {
    e[0] = e[1] = e[2] = cplx(0.0, 0.0);
    if (X >= 1.0) return 0.0;
    double st2 = sin_theta * sin_theta, ct2 = cos_theta * cos_theta;
    double rho = Y * Y * st2 * st2 + 4.0 * (1.0 - X) * (1.0 - X) * ct2;
    if (rho < 0.0) return 0.0;
    rho = std::sqrt(rho);
    double f = (2.0 * (1.0 - X)) / (2.0 * (1.0 - X) - Y * Y * st2 - (double)mode * Y * rho);
    double N = 1.0 - X * f;
    if (N < 0.0) return 0.0;
    N = std::sqrt(N);
    double g = 1.0 / Y * (1.0 - (1.0 - Y * Y) * f);
    if (ct2 < 1e-5 || 1.0 - st2 < 1e-5) {
        if (mode > 0) {
            e[1] = cplx(0.0, std::sqrt(1.0 / N));
            e[0] = cplx(0.0, g) * e[1];
        } else {
            e[2] = std::sqrt(1.0 / N);
        }
    } else {
        double den = 1.0 - X - N * N * st2;
        double a_in = 1.0 + (((1.0 - X) * N * N * ct2) / (den * den)) * 1.0 / (Y * Y) * (1.0 - (1.0 - Y * Y) * f) * (1.0 - (1.0 - Y * Y) * f);
        double a_sq = st2 * a_in * a_in;
        double b_in = 1.0 + ((1.0 - X) / den) * 1.0 / (Y * Y) * (1.0 - (1.0 - Y * Y) * f) * (1.0 - (1.0 - Y * Y) * f);
        double b_sq = ct2 * b_in * b_in;
        double ey = std::sqrt(1.0 / (N * std::sqrt(a_sq + b_sq)));
        e[1] = cplx(0.0, mode > 0 ? ey : -ey);
        e[0] = cplx(0.0, g) * e[1];
        e[2] = -cplx((N * N * sin_theta * cos_theta) / den, 0.0) * e[0];
    }
    return N;
}

// reference src/absorption.jl:132-168 + :170-189 fused over the quadrature nodes
inline double abs_Al_integral_nume_fast(const AbsQuad& q, double mu, double omega_bar, double m_0, double N_par,
                                        double N_perp, const cplx e[3], int m) {
    const double fm = (double)m;
    double x_m = N_perp * omega_bar * std::sqrt((fm / m_0) * (fm / m_0) - 1.0);
    double N_eff = (N_perp * N_par) / (1.0 - N_par * N_par);
    cplx Axz = e[0] + N_eff * e[2];
    double Axz_sq = std::abs(Axz) * std::abs(Axz);
    const cplx I(0.0, 1.0);
    double Re_Axz_ey = std::real(I * Axz * std::conj(e[1]));
    double Re_Axz_ez = std::real(Axz * std::conj(e[2]));
    double Re_ey_ez = std::real(I * std::conj(e[1]) * e[2]);
    double ey_sq = std::abs(e[1]) * std::abs(e[1]);
    double ez_sq = std::abs(e[2]) * std::abs(e[2]);
    double spar = std::sqrt(1.0 - N_par * N_par);
    double xs = x_m / (fm * spar);
    double c_abs = 0.0;
    for (size_t k = 0; k < q.t.size(); ++k) {
        double t = q.t[k];
        double u_par = 1.0 / spar * (fm / m_0 * N_par + std::sqrt((fm / m_0) * (fm / m_0) - 1.0) * t);
        double u_perp_sq = ((fm / m_0) * (fm / m_0) - 1.0) * (1.0 - t * t);
        double gamma = std::sqrt(1.0 + u_par * u_par + u_perp_sq);
        double arg = x_m * std::sqrt(1.0 - t * t);
        double Jl = jn(m - 1, arg), Jn = jn(m, arg), Ju = jn(m + 1, arg);
        double Jn2 = Jn * Jn;
        double dsq = arg / x_m * Jn * (Jl - Ju);
        double pf = (Axz_sq + ey_sq) * Jn2;
        pf += Re_Axz_ey * x_m / fm * dsq;
        pf -= (arg / fm) * (arg / fm) * ey_sq * Jl * Ju;
        pf += xs * xs * ez_sq * t * t * Jn2;
        pf += xs * 2.0 * Re_Axz_ez * t * Jn2;
        pf += xs * Re_ey_ez * t * x_m / fm * dsq;
        pf *= (fm / (N_perp * omega_bar)) * (fm / (N_perp * omega_bar));
        c_abs += q.w[k] * pf * (-mu) * std::exp(mu * (1.0 - gamma));
    }
    double a = 1.0 / (1.0 + 105.0 / (128.0 * mu * mu) + 15.0 / (8.0 * mu));
    double s = std::sqrt(mu / (2.0 * M_PI));
    return c_abs * a * (s * s * s);
}

// Rigorous upper bound of one harmonic's contribution to alpha [1/m] — the work-saving rule of the CUDA path
// (include/torj_cuda.h, torj_options.alpha_floor), restated so that bench.py can time the CPU arm doing the same work:
// |J_n| <= 1, |z J_m'| <= x_m, sum of the Gauss-Legendre weights = 2, gamma is linear in t on the resonance curve, so its
// minimum is at an end point. Never used by the checker role of the oracle (alpha_floor = 0 there).
inline double harmonic_upper_bound(double mu, double omega_bar, double m_0, double N_par, double N_perp, const cplx e[3], int m,
                                   double X, double omega, double Y) {
    const double fm = (double)m;
    const double q2 = (fm / m_0) * (fm / m_0) - 1.0, qq = std::sqrt(q2);
    const double x_m = N_perp * omega_bar * qq;
    const double N_eff = (N_perp * N_par) / (1.0 - N_par * N_par);
    const cplx Axz = e[0] + N_eff * e[2];
    const cplx I(0.0, 1.0);
    const double Axz_sq = std::norm(Axz), ey_sq = std::norm(e[1]), ez_sq = std::norm(e[2]);
    const double Re_Axz_ey = std::real(I * Axz * std::conj(e[1])), Re_Axz_ez = std::real(Axz * std::conj(e[2]));
    const double Re_ey_ez = std::real(I * std::conj(e[1]) * e[2]);
    const double spar = std::sqrt(1.0 - N_par * N_par), xs = x_m / (fm * spar);
    const double pmax = (Axz_sq + ey_sq) + 2.0 / fm * std::fabs(Re_Axz_ey) * x_m + ey_sq + ey_sq / (fm * fm) * x_m * x_m + xs * xs * ez_sq
                        + 2.0 * xs * std::fabs(Re_Axz_ez) + 2.0 / fm * xs * std::fabs(Re_ey_ez) * x_m;
    double gmin = 1e300;
    for (double t : {-1.0, 1.0}) {
        double u_par = 1.0 / spar * (fm / m_0 * N_par + qq * t);
        gmin = std::min(gmin, std::sqrt(1.0 + u_par * u_par + q2 * (1.0 - t * t)));
    }
    const double a = 1.0 / (1.0 + 105.0 / (128.0 * mu * mu) + 15.0 / (8.0 * mu));
    const double sm = std::sqrt(mu / (2.0 * M_PI));
    const double ex = std::exp(std::min(0.0, mu * (1.0 - gmin)));
    return 2.0 * pmax * (fm / (N_perp * omega_bar)) * (fm / (N_perp * omega_bar)) * mu * ex * a * sm * sm * sm * qq * 2.0 * M_PI * M_PI / m_0 * X
           * omega / (Y * constants::c);
}

// reference src/absorption.jl:191-226. alpha_floor / safe: see harmonic_upper_bound (0 / nullptr = the reference's arithmetic)
inline double abs_Albajar_fast(const AbsQuad& q, double omega, double X, double Y, double N_abs, double N_par,
                               double Te, int mode, double te_min, int max_harmonic, AbsCounters* cnt, double alpha_floor = 0.0,
                               bool* safe = nullptr) {
    if (safe) *safe = false;
    if (Te < te_min) return 0.0;
    if (cnt) cnt->n_alpha++;
    double mu = constants::m_e * constants::c * constants::c / (constants::e * Te);
    double omega_bar = 1.0 / Y;
    double c_abs = 0.0;
    double cos_theta = N_par / N_abs;
    double sin_theta = std::sin(std::acos(cos_theta));
    double N_perp = std::sqrt(N_abs * N_abs - N_par * N_par);
    cplx e[3];
    double N_test = abs_Al_N_with_pol_vec(X, Y, cos_theta, sin_theta, mode, e);
    if (std::isnan(N_test) || N_test <= 0.0 || N_test > 1.0) return 0.0;
    double m_0 = std::sqrt(1.0 - N_par * N_par) * omega_bar;
    bool ok = alpha_floor > 0.0;
    for (int m = 2; m <= max_harmonic; ++m) {
        if ((double)m < m_0) {
            if (!(m_0 - (double)m > 0.02 * m_0)) ok = false;
            continue;
        }
        if (alpha_floor > 0.0) {
            double bound = harmonic_upper_bound(mu, omega_bar, m_0, N_par, N_perp, e, m, X, omega, Y);
            if (bound < alpha_floor) {
                if (cnt) cnt->n_prune++;
                if (!(bound < alpha_floor * 1e-10)) ok = false;
                continue;
            }
        }
        ok = false;
        if (cnt) cnt->n_harm++;
        double c_m = abs_Al_integral_nume_fast(q, mu, omega_bar, m_0, N_par, N_perp, e, m);
        c_abs += std::sqrt(((double)m / m_0) * ((double)m / m_0) - 1.0) * c_m;
    }
    c_abs = -(c_abs * 2.0 * M_PI * M_PI / m_0);
    c_abs = c_abs * X * omega / (Y * constants::c);
    if (safe) *safe = ok;
    return c_abs;
}

struct RayParams {
    double omega;
    int mode;
    double te_min = 20.0;   // reference src/absorption.jl:194
    int max_harmonic = 3;   // reference src/absorption.jl:199
    int absorption_model = 0;  // 0 = Albajar (src/absorption.jl), 1 = warm-plasma α (src/general_absorption.jl:1328-1337)
    double alpha_floor = 0.0;  // > 0: the CUDA path's work-saving rules (timing of a like-for-like CPU arm only)
};

// reference src/absorption.jl:228-235
inline double alpha_approx(const Plasma& pl, const AbsQuad& q, const RayParams& rp, const double x[3],
                           const double N[3], AbsCounters* cnt, bool* safe = nullptr) {
    double N_abs = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
    PlasmaPoint<double> p = eval_plasma<double>(pl, x, N, rp.omega);
    double Te = std::exp(evaluate<double>(pl.lnTe, x));
    return abs_Albajar_fast(q, rp.omega, p.X, p.Y, N_abs, p.N_par, Te, rp.mode, rp.te_min, rp.max_harmonic, cnt, rp.alpha_floor,
                            safe);
}

// Warm-plasma damping in the place of α_approx. The reference never wires general_absorption.jl's α into gradΛ!
// (the file is not even include()d), so this wiring is BUILD-DEFINED (SURVEY.md §8(a) a21 "proposed wiring"):
//   α_warm = α(ω, X, Y, |N|, acos(N∥/|N|), Te, v_g_perp = 1/|∂Λ/∂N|, mode)[2]
// with the same Te gate as abs_Albajar_fast (src/absorption.jl:194). n_harm counts the (2 llm + 1) resonance indices n
// of the Hermitian quadrature (src/general_absorption.jl:669-710), i.e. 501 expei evaluations each.
inline double alpha_warm_approx(const Plasma& pl, const RayParams& rp, const double x[3], const double N[3],
                                double v_g_perp, AbsCounters* cnt) {
    double N_abs = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
    PlasmaPoint<double> p = eval_plasma<double>(pl, x, N, rp.omega);
    double Te = std::exp(evaluate<double>(pl.lnTe, x));
    if (Te < rp.te_min) return 0.0;
    double theta = std::acos(p.N_par / N_abs);
    warm::AlphaOut o = warm::alpha(rp.omega, p.X, p.Y, N_abs, theta, Te, v_g_perp, rp.mode);
    if (cnt) { cnt->n_alpha++; cnt->n_harm += 2 * std::min(3, o.lrm) + 1; }
    return o.alpha;
}

}  // namespace torj_oracle

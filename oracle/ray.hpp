// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Single-ray driver: ray initialisation, Hamiltonian RHS, adaptive RK integrator, sample storage.
//   reference src/solve.jl:18-38    first_point
//   reference src/solve.jl:40-74    refraction_equations!, vacuum_plasma_refraction
//   reference src/solve.jl:85-95    gradΛ!  (ForwardDiff duals here, as in the reference)
//   reference src/solve.jl:135-181  make_ray (100 chained ODEProblems, termination, storage)
// Third-party semantics restated (un-vendored, versions unpinned — SURVEY.md §8(c), App. A.2/A.3):
//   OrdinaryDiffEq: default algorithm at reltol=1e-6 -> Tsit5 (the `OwrenZen3()` in the source is passed
//   as the ODEProblem parameter object, reference src/solve.jl:155-161); PI step controller;
//   Hairer initial-dt rule; tstop snapping; every accepted step saved.
//   Roots.Bisection: exact bisection (the `xtol` keyword is not one Roots recognises, so the default
//   zero tolerances apply); NLsolve trust region with ftol=1e-12 -> Newton to |F|_inf < 1e-12.
//   IMAS.toroidal_intersection: first hit of the revolved (R,Z) grid box.
// PARITY UNPINNED against Julia: neither Julia nor the reference's golden artifact exist here.
#pragma once
#include <cmath>
#include <limits>
#include <vector>
#include "absorption.hpp"
#include "plasma.hpp"

namespace torj_oracle {

enum RayStatus {
    RAY_OK = 0,
    RAY_CUTOFF_AT_ENTRY = 1,  // reference src/solve.jl:57-59 returns (false, nothing)
    RAY_INIT_FAILED = 2,      // bisection bracket / assert src/solve.jl:32,138,141 / Newton failure
    RAY_LEFT_GRID = 3,
    RAY_MAX_STEPS = 4,
    RAY_NAN = 5,
    RAY_TRAJ_TRUNCATED = 6
};

struct Options {
    int scheme = 0;            // 0 = Tsit5 (effective reference default), 1 = OwrenZen3 (named in the source)
    int n_segments = 100;      // reference src/solve.jl:145
    double dtmax = 1e-4;       // reference src/solve.jl:157
    double abstol = 1e-6;
    double reltol = 1e-6;
    double psi_stop = 1.0;     // reference src/solve.jl:174
    double p_stop = 1e-6;      // reference src/solve.jl:176
    double te_min = 20.0;      // reference src/absorption.jl:194
    int max_harmonic = 3;      // reference src/absorption.jl:199
    int max_steps_per_segment = 100000;
    int absorption_model = 0;  // 0 = Albajar, 1 = warm-plasma α of src/general_absorption.jl (wiring build-defined)
    // The two work-saving rules of the CUDA path (include/torj_cuda.h: alpha_floor), restated so that the CPU arm of
    // bench.py can be timed doing the same work; 0 = evaluate everything as the reference does.
    double alpha_floor = 0.0;
};

struct RayCounters {
    long n_acc = 0, n_rej = 0, n_rhs = 0;
    AbsCounters abs;
};

struct RayResult {
    std::vector<double> s, x, y, z, P, dP_ds;  // all saved points (first two: launch point, plasma entry)
    std::vector<double> psi, dpsi_ds, aP;      // extras for the streaming deposition: psi_N, grad(psi).dx/ds, P*alpha
    double u_final[7] = {0, 0, 0, 0, 0, 0, 0};
    int status = RAY_OK;
    int segments_done = 0;
    RayCounters cnt;
};

// ---------------------------------------------------------------------------------------------
// RHS: reference src/solve.jl:85-95
// ---------------------------------------------------------------------------------------------
// skip_alpha / safe: the inner-stage rule of the CUDA path (only ever set when rp.alpha_floor > 0, i.e. for timing)
inline void grad_lambda(const Plasma& pl, const AbsQuad& q, const RayParams& rp, const double u[7], double du[7],
                        RayCounters* cnt, bool skip_alpha = false, bool* safe = nullptr) {
    typedef Dual<3> D;
    D xd[3] = {D::seed(u[0], 0), D::seed(u[1], 1), D::seed(u[2], 2)};
    D Nc[3] = {D(u[3]), D(u[4]), D(u[5])};
    D Lx = dispersion_relation<D>(pl, xd, Nc, rp.omega, rp.mode);  // d/dx -> du[4:6]
    D xc[3] = {D(u[0]), D(u[1]), D(u[2])};
    D Nd[3] = {D::seed(u[3], 0), D::seed(u[4], 1), D::seed(u[5], 2)};
    D LN = dispersion_relation<D>(pl, xc, Nd, rp.omega, rp.mode);  // d/dN -> du[1:3]
    double nrm = std::sqrt(LN.d[0] * LN.d[0] + LN.d[1] * LN.d[1] + LN.d[2] * LN.d[2]);
    for (int k = 0; k < 3; ++k) {
        du[k] = LN.d[k] / nrm;
        du[3 + k] = -(Lx.d[k] / nrm);
    }
    if (skip_alpha) { du[6] = 0.0; if (cnt) cnt->abs.n_askip++; }
    else if (rp.absorption_model == 1) du[6] = -u[6] * alpha_warm_approx(pl, rp, u, u + 3, 1.0 / nrm, cnt ? &cnt->abs : nullptr);
    else du[6] = -u[6] * alpha_approx(pl, q, rp, u, u + 3, cnt ? &cnt->abs : nullptr, safe);
    if (cnt) cnt->n_rhs++;
}

// ---------------------------------------------------------------------------------------------
// Ray initialisation
// ---------------------------------------------------------------------------------------------
// IMAS.toroidal_intersection restated for the rectangular grid box (reference src/solve.jl:22-24):
// smallest t>0 at which p0 + t*N0 hits the revolved rectangle [R1,R2]x[Z1,Z2].
inline bool box_intersection(const Plasma& pl, const double p0[3], const double N0[3], double* t_out) {
    double best = std::numeric_limits<double>::infinity();
    const double R1 = pl.R_first, R2 = pl.R_last, Z1 = pl.Z_first, Z2 = pl.Z_last;
    const double tiny = 1e-12;
    // cylinders R = Rc, accepted where Z within [Z1,Z2]
    double a = N0[0] * N0[0] + N0[1] * N0[1];
    double b = 2.0 * (p0[0] * N0[0] + p0[1] * N0[1]);
    for (int k = 0; k < 2; ++k) {
        double Rc = k == 0 ? R1 : R2;
        double c = p0[0] * p0[0] + p0[1] * p0[1] - Rc * Rc;
        if (a <= 0.0) continue;
        double disc = b * b - 4.0 * a * c;
        if (disc < 0.0) continue;
        double sq = std::sqrt(disc);
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            double t = (-b + sgn * sq) / (2.0 * a);
            if (t <= tiny) continue;
            double z = p0[2] + t * N0[2];
            if (z >= Z1 && z <= Z2 && t < best) best = t;
        }
    }
    // planes Z = Zc, accepted where R within [R1,R2]
    if (N0[2] != 0.0) {
        for (int k = 0; k < 2; ++k) {
            double Zc = k == 0 ? Z1 : Z2;
            double t = (Zc - p0[2]) / N0[2];
            if (t <= tiny) continue;
            double R = std::hypot(p0[0] + t * N0[0], p0[1] + t * N0[1]);
            if (R >= R1 && R <= R2 && t < best) best = t;
        }
    }
    if (!std::isfinite(best)) return false;
    *t_out = best;
    return true;
}

// reference src/solve.jl:18-38
inline int first_point(const Plasma& pl, const double p0[3], const double N0[3], double p_plasma[3]) {
    double p[3] = {p0[0], p0[1], p0[2]};
    if (!pl.on_grid(p0)) {
        double t;
        if (!box_intersection(pl, p0, N0, &t)) return RAY_INIT_FAILED;
        for (int k = 0; k < 3; ++k) p[k] = p0[k] + N0[k] * t;
    }
    auto g = [&](double t) {
        double xx[3] = {p[0] + t * N0[0], p[1] + t * N0[1], p[2] + t * N0[2]};
        return pl.psi_at(xx) - pl.psi_prof_max;
    };
    double a = 0.0, b = 0.5;  // reference src/solve.jl:29
    double ga = g(a), gb = g(b);
    double root;
    if (ga == 0.0) root = a;
    else if (gb == 0.0) root = b;
    else {
        if ((ga < 0.0) == (gb < 0.0)) return RAY_INIT_FAILED;  // Roots throws: not a bracketing interval
        for (int it = 0; it < 200; ++it) {
            double m = 0.5 * (a + b);
            if (m <= a || m >= b) break;
            double gm = g(m);
            if (gm == 0.0) { a = b = m; break; }
            if ((gm < 0.0) == (ga < 0.0)) { a = m; ga = gm; } else { b = m; gb = gm; }
        }
        // final bracket is one ulp wide; take the end with psi <= psi_prof_max so that the
        // reference's assertion src/solve.jl:138 holds without relying on the nudge at :35
        root = (ga <= 0.0) ? a : b;
    }
    for (int k = 0; k < 3; ++k) p[k] += root * N0[k];
    double psi_ref = pl.psi_at(p);
    if (!(std::fabs(psi_ref - pl.psi_prof_max) < 1e-6)) return RAY_INIT_FAILED;  // src/solve.jl:32
    if (psi_ref > pl.psi_prof_max)
        for (int k = 0; k < 3; ++k) p[k] += 2.0 * (psi_ref - pl.psi_prof_max) * N0[k];  // src/solve.jl:35
    for (int k = 0; k < 3; ++k) p_plasma[k] = p[k];
    return RAY_OK;
}

// reference src/solve.jl:40-49 on duals (the Jacobian comes from autodiff=:forward, src/solve.jl:72)
template <class T>
inline void refraction_equations(T F[3], const T N[3], double X, double Y, const double n0[3], const double n[3],
                                 const double b[3], int mode) {
    double n_dot_N = -(n[0] * n0[0] + n[1] * n0[1] + n[2] * n0[2]);
    T N_par = N[0] * b[0] + N[1] * b[1] + N[2] * b[2];
    T Ns = dsqrt(refractive_index_sq<T>(T(X), T(Y), N_par, mode));
    T coef = 1.0 / Ns * n_dot_N - dsqrt(1.0 - 1.0 / (Ns * Ns) * (1.0 - n_dot_N * n_dot_N));
    for (int k = 0; k < 3; ++k) {
        T Fk = n0[k] / Ns + coef * n[k];
        F[k] = Fk * Fk * (Ns * Ns) - N[k] * N[k];
    }
}

// Solve a 3x3 system with complete pivoting; pivots below tol are skipped (their unknown gets 0),
// which is the minimum-norm answer NLsolve's SVD fallback gives for a zero row+column
// (tor=0 central ray: F_2 and N_2 vanish identically).
inline void solve3_robust(double A[3][3], double rhs[3], double x[3]) {
    int rp[3] = {0, 1, 2}, cp[3] = {0, 1, 2};
    int rank = 0;
    double scale = 0.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(A[i][j]));
    for (int k = 0; k < 3; ++k) {
        int pi = k, pj = k; double best = 0.0;
        for (int i = k; i < 3; ++i) for (int j = k; j < 3; ++j)
            if (std::fabs(A[rp[i]][cp[j]]) > best) { best = std::fabs(A[rp[i]][cp[j]]); pi = i; pj = j; }
        if (best <= 1e-14 * scale) break;
        std::swap(rp[k], rp[pi]); std::swap(cp[k], cp[pj]);
        for (int i = k + 1; i < 3; ++i) {
            double f = A[rp[i]][cp[k]] / A[rp[k]][cp[k]];
            for (int j = k; j < 3; ++j) A[rp[i]][cp[j]] -= f * A[rp[k]][cp[j]];
            rhs[rp[i]] -= f * rhs[rp[k]];
        }
        rank = k + 1;
    }
    x[0] = x[1] = x[2] = 0.0;
    for (int k = rank - 1; k >= 0; --k) {
        double s = rhs[rp[k]];
        for (int j = k + 1; j < rank; ++j) s -= A[rp[k]][cp[j]] * x[cp[j]];
        x[cp[k]] = s / A[rp[k]][cp[k]];
    }
}

// reference src/solve.jl:51-74
inline int vacuum_plasma_refraction(const Plasma& pl, const double p[3], const double N0[3], double omega, int mode,
                                    double N_out[3]) {
    PlasmaPoint<double> pp = eval_plasma<double>(pl, p, N0, omega);
    double N_est = refractive_index_sq<double>(pp.X, pp.Y, 0.0, mode);
    if (!(N_est > 0.0)) return RAY_CUTOFF_AT_ENTRY;
    N_est = std::sqrt(N_est);
    double R = std::hypot(p[0], p[1]);
    double v, dR, dZ;
    pl.psi.eval(R, p[2], &v, &dR, &dZ);
    double n[3] = {dR * p[0] / R, dR * p[1] / R, dZ};
    double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int k = 0; k < 3; ++k) n[k] /= nn;
    double n0n = std::sqrt(N0[0] * N0[0] + N0[1] * N0[1] + N0[2] * N0[2]);
    double n0[3] = {N0[0] / n0n, N0[1] / n0n, N0[2] / n0n};
    double b[3] = {pp.b[0], pp.b[1], pp.b[2]};
    double N[3] = {N0[0] * N_est, N0[1] * N_est, N0[2] * N_est};
    typedef Dual<3> D;
    auto resid = [&](const double Nv[3], double F[3], double J[3][3]) {
        D Nd[3] = {D::seed(Nv[0], 0), D::seed(Nv[1], 1), D::seed(Nv[2], 2)};
        D Fd[3];
        refraction_equations<D>(Fd, Nd, pp.X, pp.Y, n0, n, b, mode);
        for (int i = 0; i < 3; ++i) { F[i] = Fd[i].v; if (J) for (int j = 0; j < 3; ++j) J[i][j] = Fd[i].d[j]; }
    };
    double F[3], J[3][3];
    for (int it = 0; it < 50; ++it) {
        resid(N, F, J);
        double fn = std::max(std::fabs(F[0]), std::max(std::fabs(F[1]), std::fabs(F[2])));
        if (!(fn == fn)) return RAY_INIT_FAILED;
        if (fn < 1e-12) {  // ftol, reference src/solve.jl:72
            // The reference next asserts |Λ| < 1e-12 (src/solve.jl:141); |Λ| = |sum_k F_k| can reach 3 ftol, so an
            // iteration that stops at fn just below ftol can trip that assertion. Oracle and kernel both keep
            // iterating until the assertion has a 2x margin (a deviation only where the reference would throw).
            double Np = N[0] * b[0] + N[1] * b[1] + N[2] * b[2];
            double Lam = N[0] * N[0] + N[1] * N[1] + N[2] * N[2] - refractive_index_sq<double>(pp.X, pp.Y, Np, mode);
            if (std::fabs(Lam) < 5e-13 || it >= 40) {
                for (int k = 0; k < 3; ++k) N_out[k] = N[k];
                return RAY_OK;
            }
        }
        double rhs[3] = {-F[0], -F[1], -F[2]}, dx[3];
        solve3_robust(J, rhs, dx);
        // backtracking on the residual norm (globalisation standing in for the trust region)
        double lam = 1.0;
        double f2 = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
        for (int ls = 0; ls < 30; ++ls) {
            double Nt[3] = {N[0] + lam * dx[0], N[1] + lam * dx[1], N[2] + lam * dx[2]}, Ft[3];
            resid(Nt, Ft, nullptr);
            double f2t = Ft[0] * Ft[0] + Ft[1] * Ft[1] + Ft[2] * Ft[2];
            if (f2t == f2t && f2t < f2) { for (int k = 0; k < 3; ++k) N[k] = Nt[k]; break; }
            lam *= 0.5;
            if (ls == 29) return RAY_INIT_FAILED;
        }
    }
    return RAY_INIT_FAILED;
}

// ---------------------------------------------------------------------------------------------
// Runge-Kutta tableaux (OrdinaryDiffEq)
// ---------------------------------------------------------------------------------------------
struct Tableau {
    int stages;      // including the FSAL stage
    int order;       // for the controller exponents and the initial-dt rule
    double a[7][7];  // a[i][j], row i = stage i+1
    double btilde[7];
};

inline const Tableau& tableau_tsit5() {
    static Tableau t = [] {
        Tableau T{};
        T.stages = 7; T.order = 5;
        double (*a)[7] = T.a;
        a[1][0] = 0.161;
        a[2][0] = -0.008480655492356989; a[2][1] = 0.335480655492357;
        a[3][0] = 2.8971530571054935; a[3][1] = -6.359448489975075; a[3][2] = 4.3622954328695815;
        a[4][0] = 5.325864828439257; a[4][1] = -11.748883564062828; a[4][2] = 7.4955393428898365; a[4][3] = -0.09249506636175525;
        a[5][0] = 5.86145544294642; a[5][1] = -12.92096931784711; a[5][2] = 8.159367898576159; a[5][3] = -0.071584973281401; a[5][4] = -0.028269050394068383;
        a[6][0] = 0.09646076681806523; a[6][1] = 0.01; a[6][2] = 0.4798896504144996; a[6][3] = 1.379008574103742; a[6][4] = -3.290069515436081; a[6][5] = 2.324710524099774;
        double bt[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                        0.5823571654525552, -0.45808210592918697, 0.015151515151515152};
        for (int i = 0; i < 7; ++i) T.btilde[i] = bt[i];
        return T;
    }();
    return t;
}

inline const Tableau& tableau_owrenzen3() {
    static Tableau t = [] {
        Tableau T{};
        T.stages = 4; T.order = 3;
        double (*a)[7] = T.a;
        a[1][0] = 12.0 / 23.0;
        a[2][0] = -68.0 / 375.0; a[2][1] = 368.0 / 375.0;
        a[3][0] = 31.0 / 144.0; a[3][1] = 529.0 / 1152.0; a[3][2] = 125.0 / 384.0;
        T.btilde[0] = 25.0 / 144.0; T.btilde[1] = -575.0 / 1152.0; T.btilde[2] = 125.0 / 384.0; T.btilde[3] = 0.0;
        return T;
    }();
    return t;
}

inline double eps_of(double x) {  // Julia eps(x)
    x = std::fabs(x);
    return std::nextafter(x, std::numeric_limits<double>::infinity()) - x;
}

// ---------------------------------------------------------------------------------------------
// make_ray: reference src/solve.jl:135-181 (without the deposition call at :179, see deposition.hpp)
// ---------------------------------------------------------------------------------------------
inline void make_ray(const Plasma& pl, const AbsQuad& q, const Options& opt, const double x0[3], const double N_vac[3],
                     double f, int mode, double s_max, RayResult& res, bool store) {
    RayParams rp;
    rp.omega = 2.0 * M_PI * f;
    rp.mode = mode;
    rp.te_min = opt.te_min;
    rp.max_harmonic = opt.max_harmonic;
    rp.absorption_model = opt.absorption_model;
    rp.alpha_floor = opt.absorption_model == 0 ? opt.alpha_floor : 0.0;
    const Tableau& tb = opt.scheme == 1 ? tableau_owrenzen3() : tableau_tsit5();
    const int S = tb.stages;

    double p_plasma[3];
    res.status = first_point(pl, x0, N_vac, p_plasma);
    if (res.status != RAY_OK) return;
    if (!(pl.psi_at(p_plasma) <= pl.psi_prof_max)) { res.status = RAY_INIT_FAILED; return; }  // src/solve.jl:138
    double N_plasma[3];
    res.status = vacuum_plasma_refraction(pl, p_plasma, N_vac, rp.omega, mode, N_plasma);
    if (res.status != RAY_OK) return;
    {
        double L = dispersion_relation<double>(pl, p_plasma, N_plasma, rp.omega, mode);
        if (!(std::fabs(L) < 1e-12)) { res.status = RAY_INIT_FAILED; return; }  // src/solve.jl:141
    }
    double u[7] = {p_plasma[0], p_plasma[1], p_plasma[2], N_plasma[0], N_plasma[1], N_plasma[2], 1.0};
    const double s_step = s_max / (double)opt.n_segments;
    double dx[3] = {p_plasma[0] - x0[0], p_plasma[1] - x0[1], p_plasma[2] - x0[2]};
    const double s0 = std::sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
    // dir = dx/ds at the sample (nullptr for the launch point); aP = P*alpha there
    auto push = [&](double s, const double* xx, double P, double dP, const double* dir, double aP) {
        if (!store) return;
        res.s.push_back(s); res.x.push_back(xx[0]); res.y.push_back(xx[1]); res.z.push_back(xx[2]);
        res.P.push_back(P); res.dP_ds.push_back(dP);
        double R = std::hypot(xx[0], xx[1]), v, dR, dZ;
        pl.psi.eval(R, xx[2], &v, &dR, &dZ);
        res.psi.push_back(v);
        res.dpsi_ds.push_back(dir ? dR * (xx[0] * dir[0] + xx[1] * dir[1]) / R + dZ * dir[2] : 0.0);
        res.aP.push_back(aP);
    };
    push(0.0, x0, 1.0, 0.0, nullptr, 0.0);        // src/solve.jl:149-153
    {
        double f0[7];
        RayCounters scratch;
        grad_lambda(pl, q, rp, u, f0, &scratch);
        push(s0, p_plasma, 1.0, 0.0, f0, -f0[6]);
    }

    const double beta1 = 7.0 / (10.0 * tb.order), beta2 = 2.0 / (5.0 * tb.order);
    const double gamma = 0.9, qmin = 0.2, qmax = 10.0, qoldinit = 1e-4;
    double k[7][7];
    auto rms = [](const double v[7]) { double s = 0; for (int i = 0; i < 7; ++i) s += v[i] * v[i]; return std::sqrt(s / 7.0); };

    for (int seg = 1; seg <= opt.n_segments; ++seg) {
        double t = (double)(seg - 1) * s_step + s0;
        const double tstop = (double)seg * s_step + s0;
        // ---- fresh integrator: f0, initial dt (Hairer), controller memory
        bool a_skip = false;  // inner stages of the next step may take alpha = 0 (alpha_floor > 0 only)
        grad_lambda(pl, q, rp, u, k[0], &res.cnt, false, &a_skip);
        double dt;
        {
            double sk[7], a0[7], a1[7];
            for (int i = 0; i < 7; ++i) { sk[i] = opt.abstol + std::fabs(u[i]) * opt.reltol; a0[i] = u[i] / sk[i]; a1[i] = k[0][i] / sk[i]; }
            double d0 = rms(a0), d1 = rms(a1);
            const double smalldt = 1e-6;
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? smalldt : (d0 / d1) / 100.0;
            dt0 = std::min(dt0, opt.dtmax);
            if (dt0 < 10.0 * std::numeric_limits<double>::epsilon()) {
                dt = smalldt;
            } else {
                double u1[7], f1[7], a2[7];
                for (int i = 0; i < 7; ++i) u1[i] = u[i] + dt0 * k[0][i];
                grad_lambda(pl, q, rp, u1, f1, &res.cnt, false, &a_skip);
                for (int i = 0; i < 7; ++i) a2[i] = (f1[i] - k[0][i]) / sk[i];
                double d2 = rms(a2) / dt0;
                double md = std::max(d1, d2);
                double dt1 = (md <= 1e-15) ? std::max(smalldt, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(md)) / (double)tb.order);
                dt = std::min(std::min(100.0 * dt0, dt1), opt.dtmax);
            }
        }
        double qold = qoldinit;
        int nstep = 0;
        while (t < tstop) {
            if (++nstep > opt.max_steps_per_segment) { res.status = RAY_MAX_STEPS; return; }
            dt = std::min(dt, tstop - t);
            double unew[7], tmp[7];
            for (int st = 1; st < S; ++st) {
                for (int i = 0; i < 7; ++i) {
                    double acc = 0.0;
                    for (int j = 0; j < st; ++j) acc += tb.a[st][j] * k[j][i];
                    tmp[i] = u[i] + dt * acc;
                }
                if (st < S - 1) grad_lambda(pl, q, rp, tmp, k[st], &res.cnt, a_skip && rp.alpha_floor > 0.0, nullptr);
                else grad_lambda(pl, q, rp, tmp, k[st], &res.cnt, false, &a_skip);  // FSAL stage: full alpha
            }
            for (int i = 0; i < 7; ++i) unew[i] = tmp[i];  // last stage argument is the new state (FSAL)
            double at[7];
            bool bad = false;
            for (int i = 0; i < 7; ++i) {
                double ut = 0.0;
                for (int j = 0; j < S; ++j) ut += tb.btilde[j] * k[j][i];
                ut *= dt;
                at[i] = ut / (opt.abstol + std::max(std::fabs(u[i]), std::fabs(unew[i])) * opt.reltol);
                if (!(unew[i] == unew[i])) bad = true;
            }
            if (bad) { res.status = RAY_NAN; return; }
            double EEst = rms(at);
            double qq, q11 = 0.0;
            if (EEst == 0.0) qq = 1.0 / qmax;
            else {
                q11 = std::pow(EEst, beta1);
                qq = q11 / std::pow(qold, beta2);
                qq = std::max(1.0 / qmax, std::min(1.0 / qmin, qq / gamma));
            }
            if (EEst <= 1.0) {
                res.cnt.n_acc++;
                qold = std::max(EEst, qoldinit);
                // OrdinaryDiffEq step_accept_controller! (PIController): q := 1 inside [qsteady_min, qsteady_max] = [1, 1.2]
                if (qq >= 1.0 && qq <= 1.2) qq = 1.0;
                double dtnew = dt / qq;
                double ttmp = t + dt;
                if (std::fabs(ttmp - tstop) < 100.0 * eps_of(std::max(t, tstop))) ttmp = tstop;
                t = ttmp;
                for (int i = 0; i < 7; ++i) { u[i] = unew[i]; k[0][i] = k[S - 1][i]; }
                if (u[6] < 0.0) {  // positivity callback, reference src/solve.jl:78-83,159-160
                    u[6] = 0.0;
                    grad_lambda(pl, q, rp, u, k[0], &res.cnt, false, &a_skip);
                }
                // dP/ds sample = P*alpha at the saved state (src/solve.jl:171) = -k[0][6]
                push(t, u, u[6], -k[0][6], k[0], -k[0][6]);
                dt = std::min(opt.dtmax, dtnew);
            } else {
                res.cnt.n_rej++;
                dt = dt / std::min(1.0 / qmin, q11 / gamma);
            }
        }
        res.segments_done = seg;
        if (pl.psi_at(u) > opt.psi_stop) break;  // src/solve.jl:174
        if (u[6] < opt.p_stop) break;            // src/solve.jl:176
    }
    for (int i = 0; i < 7; ++i) res.u_final[i] = u[i];
}

}  // namespace torj_oracle

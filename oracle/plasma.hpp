// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Equilibrium model, cold dispersion and the Hamiltonian RHS, restating
//   reference src/plasma.jl:2-89     (Plasma, make_2d_prof_spline, evaluate, B_spline, n_e, T_e)
//   reference src/dispersion.jl:7-39 (eval_plasma, Δ, refractive_index_sq, dispersion_relation)
//   reference src/constants.jl:13-26 (CODATA constants, digit for digit)
#pragma once
#include <cmath>
#include <vector>
#include "bspline.hpp"
#include "dual.hpp"

namespace torj_oracle {

// reference src/constants.jl:13-26
namespace constants {
constexpr double c = 2.99792458e8;
constexpr double eps0 = 8.8541878128e-12;
constexpr double e = 1.602176634e-19;
constexpr double m_e = 9.1093837015e-31;
}  // namespace constants

struct Plasma {
    double R_first = 0, R_last = 0, Z_first = 0, Z_last = 0;
    int nR = 0, nZ = 0;
    Spline2D psi, lnne, lnTe, BR, BZ, Bphi;
    Spline1D volume;
    double psi_prof_max = 0;

    // reference src/plasma.jl:16-22
    void make_2d_prof_spline(Spline2D& out, const double* psi_prof, const double* prof, int nprof,
                             const double* psi_norm_data) {
        std::vector<double> psi_range(nprof), prof2(nprof), logp(nprof);
        for (int i = 0; i < nprof; ++i)
            psi_range[i] = psi_prof[0] + (psi_prof[nprof - 1] - psi_prof[0]) * (double)i / (double)(nprof - 1);
        psi_range[nprof - 1] = psi_prof[nprof - 1];
        natural_cubic_resample(psi_prof, prof, nprof, psi_range.data(), nprof, prof2.data());
        for (int i = 0; i < nprof; ++i) logp[i] = std::log(prof2[i]);
        Spline1D s;
        s.fit(psi_range[0], psi_range[nprof - 1], logp.data(), nprof);
        std::vector<double> data((size_t)nR * nZ);
        for (size_t k = 0; k < data.size(); ++k) data[k] = s(psi_norm_data[k]);
        out.fit(R_first, R_last, nR, Z_first, Z_last, nZ, data.data());
    }

    // reference src/plasma.jl:30-58. 2-D arrays are nR x nZ with R fastest (Julia column-major).
    void build(const double* R, int nR_, const double* Z, int nZ_, const double* psi_norm_data,
               const double* psi_prof, const double* ne_prof, const double* Te_prof, int nprof,
               const double* BR_data, const double* BZ_data, const double* Bphi_data,
               const double* psi1d, const double* vol1d, int n1d) {
        nR = nR_; nZ = nZ_;
        R_first = R[0]; R_last = R[nR - 1]; Z_first = Z[0]; Z_last = Z[nZ - 1];
        psi.fit(R_first, R_last, nR, Z_first, Z_last, nZ, psi_norm_data);
        make_2d_prof_spline(lnne, psi_prof, ne_prof, nprof, psi_norm_data);
        make_2d_prof_spline(lnTe, psi_prof, Te_prof, nprof, psi_norm_data);
        BR.fit(R_first, R_last, nR, Z_first, Z_last, nZ, BR_data);
        BZ.fit(R_first, R_last, nR, Z_first, Z_last, nZ, BZ_data);
        Bphi.fit(R_first, R_last, nR, Z_first, Z_last, nZ, Bphi_data);
        std::vector<double> pr(n1d), v2(n1d);
        for (int i = 0; i < n1d; ++i) pr[i] = psi1d[0] + (psi1d[n1d - 1] - psi1d[0]) * (double)i / (double)(n1d - 1);
        pr[n1d - 1] = psi1d[n1d - 1];
        natural_cubic_resample(psi1d, vol1d, n1d, pr.data(), n1d, v2.data());
        volume.fit(pr[0], pr[n1d - 1], v2.data(), n1d);
        psi_prof_max = psi_prof[0];
        for (int i = 1; i < nprof; ++i) psi_prof_max = std::max(psi_prof_max, psi_prof[i]);
    }

    // reference src/solve.jl:7-11
    bool on_grid(const double p[3]) const {
        double R = std::hypot(p[0], p[1]);
        return R_first <= R && R <= R_last && Z_first <= p[2] && p[2] <= Z_last;
    }
    // reference src/plasma.jl:61-65 on plain doubles
    double psi_at(const double x[3]) const { return psi(std::hypot(x[0], x[1]), x[2]); }
};

// spline evaluated at a (possibly dual) position: reference src/plasma.jl:61-65 under ForwardDiff
template <class T>
inline T evaluate(const Spline2D& spl, const T x[3]) {
    T R = dhypot(x[0], x[1]);
    double v, dR, dZ;
    spl.eval(value(R), value(x[2]), &v, &dR, &dZ);
    // first-order propagation: value v, partials dR*dR/dq + dZ*dz/dq
    T out = (R - value(R)) * dR + (x[2] - value(x[2])) * dZ;
    return out + v;
}

template <class T>
struct PlasmaPoint {
    T X, Y, N_par;
    T b[3];
};

// reference src/plasma.jl:73-81 + src/dispersion.jl:7-15
template <class T>
inline PlasmaPoint<T> eval_plasma(const Plasma& pl, const T x[3], const T N[3], double omega) {
    T Br = evaluate(pl.BR, x);
    T Bp = evaluate(pl.Bphi, x);
    T Bz = evaluate(pl.BZ, x);
    T phi = datan2(x[1], x[0]);
    T cph = dcos(phi), sph = dsin(phi);
    T B[3] = {Br * cph - Bp * sph, Br * sph + Bp * cph, Bz};
    T B_abs = dsqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
    PlasmaPoint<T> p;
    for (int k = 0; k < 3; ++k) p.b[k] = B[k] / B_abs;
    p.N_par = N[0] * p.b[0] + N[1] * p.b[1] + N[2] * p.b[2];
    T ne = dexp(evaluate(pl.lnne, x));
    p.X = ne * (constants::e * constants::e / (constants::eps0 * constants::m_e * omega * omega));
    p.Y = B_abs * (constants::e / (constants::m_e * omega));
    return p;
}

// reference src/dispersion.jl:21-32
template <class T>
inline T refractive_index_sq(const T& X, const T& Y, const T& N_par, int mode) {
    T Np2 = N_par * N_par;
    T om = 1.0 - Np2;
    T delta = om * om + 4.0 * Np2 * (1.0 - X) / (Y * Y);
    T sq = dsqrt(delta);
    return 1.0 - X + (1.0 + (double)mode * sq + Np2) / (2.0 * (-1.0 + X + Y * Y)) * X * Y * Y;
}

// reference src/dispersion.jl:34-39
template <class T>
inline T dispersion_relation(const Plasma& pl, const T x[3], const T N[3], double omega, int mode) {
    T N_abs = dsqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);  // LinearAlgebra.norm(N)
    PlasmaPoint<T> p = eval_plasma(pl, x, N, omega);
    return N_abs * N_abs - refractive_index_sq(p.X, p.Y, p.N_par, mode);
}

}  // namespace torj_oracle

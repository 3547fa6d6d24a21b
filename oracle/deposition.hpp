// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Per-ray power deposition on the caller's psi_N grid.
//   deposition_faithful : restates reference src/plasma.jl:91-151 (power_deposition_profile) with the
//                         FITPACK routines of fitpack.hpp.  The reference refits Spline1D(s, psi_s - psi_j)
//                         for every level j; B-splines form a partition of unity, so that spline equals
//                         the single fit of psi_s shifted by psi_j — fitted once here.
//   deposition_streaming: CPU restatement of the streaming algorithm the CUDA kernel uses:
//                         per accepted step, cubic-Hermite psi(s) and P(s), level crossings in order,
//                         power between consecutive crossings booked to the shell in between; the open
//                         interval before the first and after the last crossing is dropped, exactly the
//                         intervals the reference's root pairing drops (src/plasma.jl:120-128).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "fitpack.hpp"
#include "plasma.hpp"

namespace torj_oracle {

// returns deposited power P; dP_dV[npsi] filled (last entry always 0, as in the reference)
inline double deposition_faithful(const Plasma& pl, const double* s, const double* x, const double* y, const double* z,
                                  const double* dP_ds, int n, const double* psi_grid, int npsi, double* dP_dV,
                                  double* dP_shell /* optional, npsi: unnormalised δP */) {
    for (int j = 0; j < npsi; ++j) { dP_dV[j] = 0.0; if (dP_shell) dP_shell[j] = 0.0; }
    std::vector<double> psi_s(n);
    for (int i = 0; i < n; ++i) psi_s[i] = pl.psi(std::hypot(x[i], y[i]), z[i]);  // src/plasma.jl:95-98
    InterpSpline dp_spl, psi_spl;
    if (!dp_spl.fit(s, dP_ds, n) || !psi_spl.fit(s, psi_s.data(), n)) return 0.0;
    double P = 0.0;
    int j = npsi - 1;  // 0-based index of the outermost level
    std::vector<double> outer_roots, inner_roots, intervals;
    psi_spl.roots(psi_grid[j], 8, outer_roots);
    double outer_volume = pl.volume(psi_grid[j]);
    j -= 1;
    while (j >= 0) {
        double inner_volume = pl.volume(psi_grid[j]);
        double dV = outer_volume - inner_volume;
        psi_spl.roots(psi_grid[j], 8, inner_roots);
        intervals = outer_roots;
        intervals.insert(intervals.end(), inner_roots.begin(), inner_roots.end());
        std::sort(intervals.begin(), intervals.end());
        if (intervals.size() < 2) break;
        if (intervals.size() % 2 != 0) intervals.pop_back();
        double dP = 0.0;
        for (size_t k = 0; k + 1 < intervals.size(); k += 2) dP += std::fabs(dp_spl.integral(intervals[k], intervals[k + 1]));
        dP_dV[j] = dP / dV;
        if (dP_shell) dP_shell[j] = dP;
        P += dP;
        j -= 1;
        outer_volume = inner_volume;
        outer_roots = inner_roots;
    }
    return P;
}

// ---- streaming algorithm (shared definition with the CUDA kernel; see DESIGN.md "deposition") ----
struct DepoState {
    int shell = -1;        // current shell index b: psi_grid[b] <= psi < psi_grid[b+1]; -1 below all, npsi-1 above all
    bool valid = false;    // a crossing has been seen (power since then is bookable)
    double P_last = 1.0;   // P at the last crossing
};

inline int locate_shell(const double* g, int n, double psi) {  // -1 if psi < g[0]; n-1 if psi >= g[n-1]
    int b = (int)(std::upper_bound(g, g + n, psi) - g) - 1;
    return b;
}

inline double hermite(double f0, double f1, double d0, double d1, double h, double th) {
    double t2 = th * th, t3 = t2 * th;
    return (2 * t3 - 3 * t2 + 1) * f0 + (t3 - 2 * t2 + th) * h * d0 + (-2 * t3 + 3 * t2) * f1 + (t3 - t2) * h * d1;
}
inline double hermite_d(double f0, double f1, double d0, double d1, double h, double th) {  // d/dtheta
    double t2 = th * th;
    return (6 * t2 - 6 * th) * f0 + (3 * t2 - 4 * th + 1) * h * d0 + (-6 * t2 + 6 * th) * f1 + (3 * t2 - 2 * th) * h * d1;
}

// one step a -> b ; dep[npsi] accumulates unnormalised δP per shell
inline void depo_step(DepoState& st, const double* g, int npsi, double h, double psi_a, double psi_b, double dpsi_a,
                      double dpsi_b, double P_a, double P_b, double dP_a, double dP_b /* dP/ds = -P*alpha */, double* dep) {
    while (true) {
        int lvl;   // level that would be crossed next
        int nshell;
        if (psi_b < psi_a) {  // moving inward: next level is the lower edge of the current shell
            lvl = st.shell;
            if (lvl < 0 || !(psi_b < g[lvl])) return;
            nshell = lvl - 1;
        } else {              // moving outward: next level is the upper edge
            lvl = st.shell + 1;
            if (lvl > npsi - 1 || !(psi_b >= g[lvl])) return;
            nshell = lvl;
        }
        // crossing abscissa: Newton on the Hermite cubic from the chord estimate
        double th = (g[lvl] - psi_a) / (psi_b - psi_a);
        for (int it = 0; it < 3; ++it) {
            double f = hermite(psi_a, psi_b, dpsi_a, dpsi_b, h, th) - g[lvl];
            double d = hermite_d(psi_a, psi_b, dpsi_a, dpsi_b, h, th);
            if (d != 0.0) th -= f / d;
            th = std::min(1.0, std::max(0.0, th));
        }
        double Pc = hermite(P_a, P_b, dP_a, dP_b, h, th);
        if (st.valid && st.shell >= 0 && st.shell < npsi - 1) dep[st.shell] += std::fabs(st.P_last - Pc);
        st.P_last = Pc;
        st.valid = true;
        st.shell = nshell;
    }
}

// whole-ray streaming deposition from stored samples. dpsi_ds[i] = grad(psi).dx/ds at sample i (0 for the
// launch point; the vacuum leg carries P=1 so its crossing abscissae do not matter).
inline double deposition_streaming(const Plasma& pl, const double* s, const double* psi_s, const double* dpsi_ds,
                                   const double* P, const double* dP_ds_pos /* = P*alpha >= 0 */, int n,
                                   const double* psi_grid, int npsi, double* dP_dV, double* dP_shell) {
    std::vector<double> dep(npsi, 0.0);
    DepoState st;
    st.shell = locate_shell(psi_grid, npsi, psi_s[0]);
    st.valid = false;
    st.P_last = P[0];
    for (int i = 0; i + 1 < n; ++i) {
        double h = s[i + 1] - s[i];
        if (i == 0) {
            // vacuum leg launch -> plasma entry: straight line, no absorption
            double d = (psi_s[1] - psi_s[0]) / h;
            depo_step(st, psi_grid, npsi, h, psi_s[0], psi_s[1], d, d, P[0], P[1], 0.0, 0.0, dep.data());
        } else {
            depo_step(st, psi_grid, npsi, h, psi_s[i], psi_s[i + 1], dpsi_ds[i], dpsi_ds[i + 1], P[i], P[i + 1],
                      -dP_ds_pos[i], -dP_ds_pos[i + 1], dep.data());
        }
    }
    double Ptot = 0.0;
    for (int j = 0; j < npsi; ++j) { dP_dV[j] = 0.0; if (dP_shell) dP_shell[j] = 0.0; }
    for (int j = 0; j + 1 < npsi; ++j) {
        double dV = pl.volume(psi_grid[j + 1]) - pl.volume(psi_grid[j]);
        dP_dV[j] = dep[j] / dV;
        if (dP_shell) dP_shell[j] = dep[j];
        Ptot += dep[j];
    }
    return Ptot;
}

}  // namespace torj_oracle

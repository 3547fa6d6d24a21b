// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Beam -> ray bundle, restating reference src/launch.jl:24-132 (launch_peripheral_rays) including its
// quirks (e_χ[3] = -n0[1]/n0[3]; e_υ[2] = n0[3]-n0[1]).
// FastGaussQuadrature.gausshermite (un-vendored, compat 1.0.2) -> Golub-Welsch start + Newton polish on the
// orthonormal Hermite recurrence + Christoffel weights (physicists' weight exp(-x^2)); pinned against
// numpy.polynomial.hermite.hermgauss in tests/.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "plasma.hpp"

namespace torj_oracle {

inline void gausshermite(int n, std::vector<double>& x, std::vector<double>& w) {
    // Jacobi matrix: diag 0, off-diagonal^2 = k/2; eigenvalues by Sturm-count bisection (start values only)
    std::vector<double> d(n, 0.0);
    auto count_below = [&](double xx) {
        int cnt = 0;
        double qv = -xx;
        if (qv < 0) ++cnt;
        for (int k = 1; k < n; ++k) {
            if (qv == 0.0) qv = 1e-300;
            qv = -xx - (0.5 * k) / qv;
            if (qv < 0) ++cnt;
        }
        return cnt;
    };
    double bound = std::sqrt(2.0 * n + 1.0) + 1.0;
    for (int i = 0; i < n; ++i) {
        double lo = -bound, hi = bound;
        for (int it = 0; it < 80; ++it) {
            double mid = 0.5 * (lo + hi);
            if (count_below(mid) <= i) lo = mid; else hi = mid;
        }
        d[i] = 0.5 * (lo + hi);
    }
    std::sort(d.begin(), d.end());
    x.assign(n, 0.0); w.assign(n, 0.0);
    const double pim4 = std::pow(M_PI, -0.25);
    for (int i = 0; i < n; ++i) {
        double z = d[i];
        for (int it = 0; it < 10; ++it) {
            // orthonormal Hermite functions p_k(z) (w.r.t. exp(-z^2)): p_0 = pi^-1/4
            double p0 = pim4, p1 = 0.0;  // p_k, p_{k-1}
            for (int k = 0; k < n; ++k) {
                double p2 = z * std::sqrt(2.0 / (k + 1.0)) * p0 - std::sqrt((double)k / (k + 1.0)) * p1;
                p1 = p0; p0 = p2;
            }
            double dp = std::sqrt(2.0 * n) * p1;
            double dz = p0 / dp;
            z -= dz;
            if (std::fabs(dz) < 1e-16 * std::max(1.0, std::fabs(z))) break;
        }
        if (n % 2 == 1 && i == n / 2) z = 0.0;
        x[i] = z;
    }
    for (int i = 0; i < n / 2; ++i) { double a = 0.5 * (x[n - 1 - i] - x[i]); x[i] = -a; x[n - 1 - i] = a; }
    for (int i = 0; i < n; ++i) {
        double p0 = pim4, p1 = 0.0, sum = 0.0;
        for (int k = 0; k < n; ++k) {
            sum += p0 * p0;
            double p2 = x[i] * std::sqrt(2.0 / (k + 1.0)) * p0 - std::sqrt((double)k / (k + 1.0)) * p1;
            p1 = p0; p0 = p2;
        }
        w[i] = 1.0 / sum;
    }
}

struct Bundle {
    std::vector<double> pos, dir, weight;  // pos/dir: [3][n] (component-major = Julia N x 3 column-major)
    int n = 0;
};

// reference src/launch.jl:24-132
inline int launch_peripheral_rays(const double x0[3], const double N0[3], double w, double inv_Rc, double f,
                                  int N_rings, int min_az, bool normalize, Bundle& out) {
    if (N_rings < 2) return -1;  // ArgumentError, src/launch.jl:27-29
    double nn = std::sqrt(N0[0] * N0[0] + N0[1] * N0[1] + N0[2] * N0[2]);
    double n0[3] = {N0[0] / nn, N0[1] / nn, N0[2] / nn};
    bool curved = std::isfinite(inv_Rc);
    double w0 = w, x_waist[3] = {0, 0, 0};
    if (curved) {
        double R_curv = 1.0 / inv_Rc;
        double lam = constants::c / f;
        w0 = (lam * std::fabs(R_curv) * w) / std::sqrt(lam * lam * R_curv * R_curv + M_PI * M_PI * w * w * w * w);
        double z_waist = M_PI * M_PI * R_curv * w * w * w * w / (lam * lam * R_curv * R_curv + M_PI * M_PI * w * w * w * w);
        for (int k = 0; k < 3; ++k) x_waist[k] = x0[k] - n0[k] * z_waist;
    }
    double e_chi[3] = {1.0, 0.0, -n0[0] / n0[2]};
    double e_ups[3] = {-n0[0] * n0[1] / n0[2], n0[2] - n0[0], -n0[1]};
    double nc = std::sqrt(e_chi[0] * e_chi[0] + e_chi[1] * e_chi[1] + e_chi[2] * e_chi[2]);
    double nu = std::sqrt(e_ups[0] * e_ups[0] + e_ups[1] * e_ups[1] + e_ups[2] * e_ups[2]);
    for (int k = 0; k < 3; ++k) { e_chi[k] /= nc; e_ups[k] /= nu; }
    std::vector<double> gx, gw;
    gausshermite(2 * N_rings + 2, gx, gw);
    // v[N_rings+2:end] (1-based) of 2N+2 ascending nodes -> 0-based N_rings+1 .. 2N+1 : N_rings+1 entries,
    // of which only the first N_rings are used by the loops below (src/launch.jl:72-83)
    std::vector<double> r_pts, r_w;
    for (int i = N_rings + 1; i < 2 * N_rings + 2; ++i) { r_pts.push_back(gx[i] * (w / std::sqrt(2.0))); r_w.push_back(gw[i] * (w / std::sqrt(2.0))); }
    std::vector<long> Nth(N_rings);
    long total = 0;
    for (int i = 0; i < N_rings; ++i) {
        double v = (double)min_az * r_pts[i] / r_pts[0];
        long r = (long)std::nearbyint(v);  // Julia round(Int64, x): ties to even, like nearbyint in the default mode
        Nth[i] = std::max(1L, r);
        total += Nth[i];
    }
    out.n = (int)total;
    out.pos.assign(3 * total, 0.0); out.dir.assign(3 * total, 0.0); out.weight.assign(total, 0.0);
    long k = 0;
    double sgn = inv_Rc > 0 ? 1.0 : (inv_Rc < 0 ? -1.0 : 0.0);
    for (int i = 0; i < N_rings; ++i) {
        for (long j = 0; j < Nth[i]; ++j) {
            double th = 2.0 * M_PI * (double)j / (double)Nth[i];
            double chi = r_pts[i] * std::cos(th), ups = r_pts[i] * std::sin(th);
            double p[3], d[3];
            for (int c = 0; c < 3; ++c) p[c] = chi * e_chi[c] + ups * e_ups[c] + x0[c];
            if (curved) {
                for (int c = 0; c < 3; ++c) d[c] = w0 / w * (chi * e_chi[c] + ups * e_ups[c]) * sgn + x_waist[c];
                if (inv_Rc < 0.0) { for (int c = 0; c < 3; ++c) d[c] -= p[c]; }
                else { for (int c = 0; c < 3; ++c) d[c] = -d[c] + p[c]; }
                double dn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (int c = 0; c < 3; ++c) d[c] /= dn;
            } else {
                for (int c = 0; c < 3; ++c) d[c] = n0[c];
            }
            for (int c = 0; c < 3; ++c) { out.pos[c * total + k + j] = p[c]; out.dir[c * total + k + j] = d[c]; }
            out.weight[k + j] = r_pts[i] * r_w[i] * (2.0 * M_PI / (double)Nth[i]);
        }
        k += Nth[i];
    }
    if (normalize) {
        double sum = 0.0;
        for (long i = 0; i < total; ++i) sum += out.weight[i];
        for (long i = 0; i < total; ++i) out.weight[i] /= sum;
    } else {
        for (long i = 0; i < total; ++i) out.weight[i] *= 2.0 / (w * w * M_PI);
    }
    return 0;
}

}  // namespace torj_oracle

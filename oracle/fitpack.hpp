// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Restatement of the three FITPACK (Dierckx) routines the reference reaches through Dierckx.jl
// (un-vendored, compat 0.5.4, Fortran via Dierckx_jll) at reference src/plasma.jl:101-134:
//   Dierckx.Spline1D(x, y, k=3)   -> curfit, s=0 : interpolating cubic spline, knots
//                                    t = [x1 x1 x1 x1, x3 .. x_{m-2}, xm xm xm xm]  (fpcurf.f, s=0 branch)
//   Dierckx.roots(spl)            -> sproot : all zeros of a cubic spline (first `maxn`=8 kept)
//   Dierckx.integrate(spl, a, b)  -> splint : integral, limits clamped to the knot range
// scipy.interpolate.splrep/sproot/splint wrap the same Fortran, so tests pin this file against
// scipy (tests/test_oracle_fitpack.py, fixtures in tests/golden/).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace torj_oracle {

struct InterpSpline {
    int m = 0;                // data points = number of coefficients
    std::vector<double> t;    // m+4 knots
    std::vector<double> c;    // m B-spline coefficients
    // piecewise-polynomial form on the distinct knot intervals l = 3 .. m-1 : p(y) = a0+a1 y+a2 y^2+a3 y^3, y=x-t[l]
    std::vector<double> pp;   // 4 per interval
    std::vector<double> cum;  // integral from t[3] to t[l]

    static void bsplvals(const double* t, double x, int l, double h[4]) {  // fpbspl, k=3; t[l] <= x < t[l+1]
        double hh[3];
        h[0] = 1.0;
        for (int j = 1; j <= 3; ++j) {
            for (int i = 0; i < j; ++i) hh[i] = h[i];
            h[0] = 0.0;
            for (int i = 0; i < j; ++i) {
                int li = l + i + 1, lj = li - j;
                double f = hh[i] / (t[li] - t[lj]);
                h[i] += f * (t[li] - x);
                h[i + 1] = f * (x - t[lj]);
            }
        }
    }

    bool fit(const double* x, const double* y, int m_) {
        m = m_;
        if (m < 4) return false;
        for (int i = 1; i < m; ++i) if (!(x[i] > x[i - 1])) return false;
        t.assign(m + 4, 0.0);
        for (int i = 0; i < 4; ++i) { t[i] = x[0]; t[m + i] = x[m - 1]; }
        for (int i = 4; i < m; ++i) t[i] = x[i - 2];
        // banded collocation matrix, kl = ku = 3, row i holds columns i-3..i+3 in A[i][0..6]
        std::vector<double> A((size_t)m * 7, 0.0), rhs(y, y + m);
        int l = 3;
        for (int i = 0; i < m; ++i) {
            while (l < m - 1 && x[i] >= t[l + 1]) ++l;
            double h[4];
            bsplvals(t.data(), x[i], l, h);
            for (int q = 0; q < 4; ++q) {
                int col = l - 3 + q;
                A[(size_t)i * 7 + (col - i + 3)] = h[q];
            }
        }
        // Gaussian elimination without pivoting (B-spline collocation matrices are totally positive)
        for (int k = 0; k < m; ++k) {
            double piv = A[(size_t)k * 7 + 3];
            for (int i = k + 1; i <= std::min(k + 3, m - 1); ++i) {
                double f = A[(size_t)i * 7 + (k - i + 3)] / piv;
                if (f == 0.0) continue;
                for (int col = k; col <= std::min(k + 3, m - 1); ++col)
                    A[(size_t)i * 7 + (col - i + 3)] -= f * A[(size_t)k * 7 + (col - k + 3)];
                rhs[i] -= f * rhs[k];
            }
        }
        c.assign(m, 0.0);
        for (int k = m - 1; k >= 0; --k) {
            double sum = rhs[k];
            for (int col = k + 1; col <= std::min(k + 3, m - 1); ++col) sum -= A[(size_t)k * 7 + (col - k + 3)] * c[col];
            c[k] = sum / A[(size_t)k * 7 + 3];
        }
        build_pp();
        return true;
    }

    // derivatives 0..3 of the spline at x = t[l]+ (right-continuous), knot interval l
    void derivs_at_left(int l, double out[4]) const {
        double d[4];
        for (int i = 0; i < 4; ++i) d[i] = c[l - 3 + i];  // d[i] <-> index l-3+i
        double fact = 1.0;
        const double x = t[l];
        for (int r = 0; r <= 3; ++r) {
            int p = 3 - r;  // current degree; relevant coefficient indices l-p..l are d[r..3]
            double e[4];
            for (int i = 0; i < 4; ++i) e[i] = d[i];
            for (int j = 1; j <= p; ++j)
                for (int i = l; i >= l - p + j; --i) {
                    double den = t[i + p - j + 1] - t[i];
                    double alpha = (x - t[i]) / den;
                    e[i - l + 3] = (1.0 - alpha) * e[i - 1 - l + 3] + alpha * e[i - l + 3];
                }
            out[r] = e[3] / fact;  // divided by r! to give the Taylor coefficient
            fact *= (double)(r + 1);
            if (p > 0)
                for (int i = l; i >= l - p + 1; --i)
                    d[i - l + 3] = (double)p * (d[i - l + 3] - d[i - 1 - l + 3]) / (t[i + p] - t[i]);
        }
    }

    void build_pp() {
        int nint = m - 3;  // intervals l = 3..m-1
        pp.assign((size_t)nint * 4, 0.0);
        cum.assign(nint + 1, 0.0);
        for (int l = 3; l < m; ++l) {
            double a[4];
            derivs_at_left(l, a);
            double* q = &pp[(size_t)(l - 3) * 4];
            for (int r = 0; r < 4; ++r) q[r] = a[r];
            double h = t[l + 1] - t[l];
            cum[l - 3 + 1] = cum[l - 3] + h * (a[0] + h * (a[1] / 2.0 + h * (a[2] / 3.0 + h * a[3] / 4.0)));
        }
    }

    int interval_of(double x) const {  // l in 3..m-1 with t[l] <= x < t[l+1] (last interval closed)
        int l = (int)(std::upper_bound(t.begin() + 3, t.begin() + m + 1, x) - t.begin()) - 1;
        return std::min(std::max(l, 3), m - 1);
    }
    double value(double x) const {
        int l = interval_of(x);
        const double* q = &pp[(size_t)(l - 3) * 4];
        double y = x - t[l];
        return q[0] + y * (q[1] + y * (q[2] + y * q[3]));
    }
    double antideriv(double x) const {
        x = std::min(std::max(x, t[3]), t[m]);  // splint: the spline is taken as zero outside [t(k+1), t(n-k)]
        int l = interval_of(x);
        const double* q = &pp[(size_t)(l - 3) * 4];
        double y = x - t[l];
        return cum[l - 3] + y * (q[0] + y * (q[1] / 2.0 + y * (q[2] / 3.0 + y * q[3] / 4.0)));
    }
    double integral(double a, double b) const { return antideriv(b) - antideriv(a); }

    // zeros of (spline - level), ascending, at most maxn (Dierckx.roots default maxn=8)
    void roots(double level, int maxn, std::vector<double>& out) const {
        out.clear();
        for (int l = 3; l < m; ++l) {
            const double* q = &pp[(size_t)(l - 3) * 4];
            double h = t[l + 1] - t[l];
            double a0 = q[0] - level, a1 = q[1], a2 = q[2], a3 = q[3];
            auto P = [&](double y) { return a0 + y * (a1 + y * (a2 + y * a3)); };
            // split [0,h] at the critical points of the cubic into monotone pieces
            double br[4]; int nb = 0;
            br[nb++] = 0.0;
            double A = 3.0 * a3, B = 2.0 * a2, C = a1;
            double cr[2]; int nc = 0;
            if (A != 0.0) {
                double disc = B * B - 4.0 * A * C;
                if (disc > 0.0) {
                    double sq = std::sqrt(disc);
                    double qq = -0.5 * (B + (B >= 0 ? sq : -sq));
                    double r1 = qq / A, r2 = (qq != 0.0) ? C / qq : r1;
                    if (r1 > r2) std::swap(r1, r2);
                    cr[nc++] = r1; cr[nc++] = r2;
                }
            } else if (B != 0.0) {
                cr[nc++] = -C / B;
            }
            for (int i = 0; i < nc; ++i) if (cr[i] > 0.0 && cr[i] < h) br[nb++] = cr[i];
            br[nb++] = h;
            bool last = (l == m - 1);
            for (int i = 0; i + 1 < nb; ++i) {
                double ya = br[i], yb = br[i + 1];
                double fa = P(ya), fb = P(yb);
                double root;
                if (fa == 0.0) root = ya;
                else if (fb == 0.0) { if (i + 2 < nb || last) root = yb; else continue; }  // right knot belongs to the next interval
                else if ((fa < 0.0) != (fb < 0.0)) {
                    double lo = ya, hi = yb, flo = fa;
                    for (int it = 0; it < 200; ++it) {
                        double mid = 0.5 * (lo + hi);
                        if (mid <= lo || mid >= hi) break;
                        double fm = P(mid);
                        if (fm == 0.0) { lo = hi = mid; break; }
                        if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else hi = mid;
                    }
                    root = 0.5 * (lo + hi);
                } else continue;
                double xr = t[l] + root;
                if (!out.empty() && xr <= out.back()) continue;
                out.push_back(xr);
                if ((int)out.size() >= maxn) return;
            }
        }
    }
};

}  // namespace torj_oracle

// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Cubic B-spline interpolation with the semantics of Interpolations.jl
//   cubic_spline_interpolation(ranges, A; extrapolation_bc=Line())
//   = extrapolate(scale(interpolate(A, BSpline(Cubic(Line(OnGrid())))), ranges...), Line())
// as used by the reference at src/plasma.jl:19,21,36-44 and evaluated at src/plasma.jl:61-65,
// src/solve.jl:63.  Interpolations.jl is an un-vendored dependency (compat 0.16.1,
// reference Project.toml:37); the algorithm restated here is its published one:
//   prefilter : n+2 coefficients, interior rows (1/6, 2/3, 1/6), boundary rows c0-2c1+c2 = 0
//   evaluate  : u=(x-x1)/h+1, i=clamp(floor(u),1,n-1), d=u-i, cubic B-spline weights on c[i-1..i+2]
//   outside   : clamp, evaluate value+gradient there, add gradient*(x-x_clamped)   ("Line")
// Also: natural cubic spline on a non-uniform grid (IMAS.interp1d(x, y, :cubic), reference
// src/plasma.jl:18,43 — DataInterpolations.CubicSpline, natural end conditions).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace torj_oracle {

// Solve for the n+2 cubic B-spline coefficients of data y[0..n-1] (stride ys), write c[0..n+1] (stride cs).
// Boundary rows give c[1]=y[0], c[n]=y[n-1]; the rest is a tridiagonal (1/6,2/3,1/6) system.
inline void bspline_prefilter_1d(const double* y, int ys, int n, double* c, int cs) {
    std::vector<double> sol(n), cp(n), dp(n);
    sol[0] = y[0];
    sol[n - 1] = y[(n - 1) * ys];
    if (n > 2) {
        // unknowns sol[1..n-2]; rows: sol[k-1]/6 + 2 sol[k]/3 + sol[k+1]/6 = y[k]
        const double a = 1.0 / 6.0, b = 2.0 / 3.0;
        int m = n - 2;
        std::vector<double> rhs(m);
        for (int k = 0; k < m; ++k) rhs[k] = y[(k + 1) * ys];
        rhs[0] -= a * sol[0];
        rhs[m - 1] -= a * sol[n - 1];
        cp[0] = a / b;
        dp[0] = rhs[0] / b;
        for (int k = 1; k < m; ++k) {
            double den = b - a * cp[k - 1];
            cp[k] = a / den;
            dp[k] = (rhs[k] - a * dp[k - 1]) / den;
        }
        sol[m] = dp[m - 1];
        for (int k = m - 2; k >= 0; --k) sol[k + 1] = dp[k] - cp[k] * sol[k + 2];
    }
    for (int k = 0; k < n; ++k) c[(k + 1) * cs] = sol[k];
    c[0] = 2.0 * sol[0] - sol[1];
    c[(n + 1) * cs] = 2.0 * sol[n - 1] - sol[n - 2];
}

inline void bspline_weights(double d, double h, double w[4], double dw[4]) {
    double e = 1.0 - d;
    w[0] = e * e * e / 6.0;
    w[1] = 2.0 / 3.0 - d * d + d * d * d / 2.0;
    w[2] = 2.0 / 3.0 - e * e + e * e * e / 2.0;
    w[3] = d * d * d / 6.0;
    dw[0] = -e * e / 2.0 / h;
    dw[1] = (-2.0 * d + 1.5 * d * d) / h;
    dw[2] = (2.0 * e - 1.5 * e * e) / h;
    dw[3] = d * d / 2.0 / h;
}

// cell lookup shared by 1-D and 2-D evaluation: returns 0-based index of c[i-1] and fills weights
inline int bspline_locate(double x, double x0, double h, int n, double w[4], double dw[4]) {
    double u = (x - x0) / h + 1.0;
    int i = (int)std::floor(u);
    i = std::min(std::max(i, 1), n - 1);
    bspline_weights(u - (double)i, h, w, dw);
    return i - 1;
}

struct Spline1D {
    int n = 0;
    double x0 = 0, h = 1;
    std::vector<double> c;  // n+2

    void fit(double xfirst, double xlast, const double* y, int n_) {
        n = n_;
        x0 = xfirst;
        h = (xlast - xfirst) / (double)(n - 1);
        c.assign(n + 2, 0.0);
        bspline_prefilter_1d(y, 1, n, c.data(), 1);
    }
    double xlast() const { return x0 + h * (n - 1); }
    void eval_in(double x, double* v, double* dv) const {
        double w[4], dw[4];
        int b = bspline_locate(x, x0, h, n, w, dw);
        double s = 0, ds = 0;
        for (int k = 0; k < 4; ++k) { s += w[k] * c[b + k]; ds += dw[k] * c[b + k]; }
        *v = s; *dv = ds;
    }
    double operator()(double x) const {
        double xc = std::min(std::max(x, x0), xlast());
        double v, dv;
        eval_in(xc, &v, &dv);
        if (xc != x) v += (x - xc) * dv;
        return v;
    }
};

struct Spline2D {
    int nr = 0, nz = 0;
    double r0 = 0, hr = 1, z0 = 0, hz = 1;
    std::vector<double> c;  // (nr+2) x (nz+2), R fastest

    // data: nr x nz, R fastest (Julia column-major Matrix[nR, nZ])
    void fit(double rfirst, double rlast, int nr_, double zfirst, double zlast, int nz_, const double* data) {
        nr = nr_; nz = nz_;
        r0 = rfirst; z0 = zfirst;
        hr = (rlast - rfirst) / (double)(nr - 1);
        hz = (zlast - zfirst) / (double)(nz - 1);
        int sr = nr + 2;
        std::vector<double> tmp((size_t)sr * nz);
        for (int j = 0; j < nz; ++j) bspline_prefilter_1d(data + (size_t)j * nr, 1, nr, tmp.data() + (size_t)j * sr, 1);
        c.assign((size_t)sr * (nz + 2), 0.0);
        for (int i = 0; i < sr; ++i) bspline_prefilter_1d(tmp.data() + i, sr, nz, c.data() + i, sr);
    }
    double rlast() const { return r0 + hr * (nr - 1); }
    double zlast() const { return z0 + hz * (nz - 1); }

    // value, d/dR, d/dZ and mixed d2/dRdZ inside the grid
    void eval_in(double R, double Z, double* v, double* dR, double* dZ, double* dRZ) const {
        double wr[4], dwr[4], wz[4], dwz[4];
        int br = bspline_locate(R, r0, hr, nr, wr, dwr);
        int bz = bspline_locate(Z, z0, hz, nz, wz, dwz);
        int sr = nr + 2;
        double s = 0, sR = 0, sZ = 0, sRZ = 0;
        for (int j = 0; j < 4; ++j) {
            const double* row = c.data() + (size_t)(bz + j) * sr + br;
            double a = 0, aR = 0;
            for (int i = 0; i < 4; ++i) { a += wr[i] * row[i]; aR += dwr[i] * row[i]; }
            s += wz[j] * a; sR += wz[j] * aR; sZ += dwz[j] * a; sRZ += dwz[j] * aR;
        }
        *v = s; *dR = sR; *dZ = sZ; *dRZ = sRZ;
    }
    // with "Line" extrapolation: v = itp(xc) + sum_d (x_d - xc_d) g_d(xc); gradient of that expression
    void eval(double R, double Z, double* v, double* dR, double* dZ) const {
        double Rc = std::min(std::max(R, r0), rlast());
        double Zc = std::min(std::max(Z, z0), zlast());
        double s, gR, gZ, gRZ;
        eval_in(Rc, Zc, &s, &gR, &gZ, &gRZ);
        bool outR = (Rc != R), outZ = (Zc != Z);
        double val = s, vR = gR, vZ = gZ;
        if (outR) { val += (R - Rc) * gR; if (!outZ) vZ += (R - Rc) * gRZ; }
        if (outZ) { val += (Z - Zc) * gZ; if (!outR) vR += (Z - Zc) * gRZ; }
        *v = val; *dR = vR; *dZ = vZ;
    }
    double operator()(double R, double Z) const {
        double v, a, b;
        eval(R, Z, &v, &a, &b);
        return v;
    }
};

// Natural cubic spline through (x[i], y[i]), x strictly increasing, evaluated at xq in [x[0], x[n-1]].
inline void natural_cubic_resample(const double* x, const double* y, int n, const double* xq, int nq, double* yq) {
    std::vector<double> m(n, 0.0);  // second derivatives
    if (n > 2) {
        std::vector<double> a(n), b(n), cdiag(n), r(n);
        for (int i = 1; i < n - 1; ++i) {
            double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
            a[i] = h0; b[i] = 2.0 * (h0 + h1); cdiag[i] = h1;
            r[i] = 6.0 * ((y[i + 1] - y[i]) / h1 - (y[i] - y[i - 1]) / h0);
        }
        // Thomas on rows 1..n-2 with m[0]=m[n-1]=0
        std::vector<double> cp(n), dp(n);
        cp[1] = cdiag[1] / b[1];
        dp[1] = r[1] / b[1];
        for (int i = 2; i < n - 1; ++i) {
            double den = b[i] - a[i] * cp[i - 1];
            cp[i] = cdiag[i] / den;
            dp[i] = (r[i] - a[i] * dp[i - 1]) / den;
        }
        m[n - 2] = dp[n - 2];
        for (int i = n - 3; i >= 1; --i) m[i] = dp[i] - cp[i] * m[i + 1];
    }
    for (int q = 0; q < nq; ++q) {
        double t = xq[q];
        int i = (int)(std::upper_bound(x, x + n, t) - x) - 1;
        i = std::min(std::max(i, 0), n - 2);
        double h = x[i + 1] - x[i];
        double A = (x[i + 1] - t) / h, B = (t - x[i]) / h;
        yq[q] = A * y[i] + B * y[i + 1] + ((A * A * A - A) * m[i] + (B * B * B - B) * m[i + 1]) * h * h / 6.0;
    }
}

}  // namespace torj_oracle

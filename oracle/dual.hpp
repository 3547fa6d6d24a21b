// TEST INFRASTRUCTURE — CPU oracle for the TorJ ray-tracing hot path. Not part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Forward-mode dual numbers with N partials: the oracle differentiates the dispersion relation the
// way the reference does (ForwardDiff.gradient!, reference src/solve.jl:89-90), so that the CUDA
// kernel's hand-derived analytic gradients are checked against an independent derivation.
#pragma once
#include <cmath>

namespace torj_oracle {

template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
    Dual(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
    static Dual seed(double x, int k) { Dual r(x); r.d[k] = 1.0; return r; }
};

inline double value(double x) { return x; }
template <int N> inline double value(const Dual<N>& x) { return x.v; }

template <int N> inline Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r;
}
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r;
}
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r;
}
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r;
}
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v / b.v;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return b + a; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { return (-b) + a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) {
    Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r;
}
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) {
    Dual<N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r;
}
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }

inline double dsqrt(double x) { return std::sqrt(x); }
inline double dexp(double x) { return std::exp(x); }
inline double dsin(double x) { return std::sin(x); }
inline double dcos(double x) { return std::cos(x); }
inline double datan2(double y, double x) { return std::atan2(y, x); }
inline double dhypot(double x, double y) { return std::hypot(x, y); }

template <int N> inline Dual<N> dsqrt(const Dual<N>& a) {
    Dual<N> r; r.v = std::sqrt(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / (2.0 * r.v);
    return r;
}
template <int N> inline Dual<N> dexp(const Dual<N>& a) {
    Dual<N> r; r.v = std::exp(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * r.v;
    return r;
}
template <int N> inline Dual<N> dsin(const Dual<N>& a) {
    Dual<N> r; r.v = std::sin(a.v); double c = std::cos(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
    return r;
}
template <int N> inline Dual<N> dcos(const Dual<N>& a) {
    Dual<N> r; r.v = std::cos(a.v); double s = -std::sin(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s;
    return r;
}
template <int N> inline Dual<N> datan2(const Dual<N>& y, const Dual<N>& x) {
    Dual<N> r; r.v = std::atan2(y.v, x.v);
    double h2 = x.v * x.v + y.v * y.v;
    for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) / h2;
    return r;
}
template <int N> inline Dual<N> dhypot(const Dual<N>& x, const Dual<N>& y) {
    Dual<N> r; r.v = std::hypot(x.v, y.v);
    for (int i = 0; i < N; ++i) r.d[i] = (x.v * x.d[i] + y.v * y.d[i]) / r.v;
    return r;
}

}  // namespace torj_oracle

// Device-side physics of the TorJ ray-tracing hot path for sm_100a (FP64 throughout).
//
//   spline tables / evaluation   <- reference src/plasma.jl:61-89 (Interpolations.jl cubic B-splines, Line extrapolation)
//   cold dispersion + analytic   <- reference src/dispersion.jl:7-39 and the ForwardDiff gradients of
//   Hamiltonian RHS                 src/solve.jl:85-95 (derived by hand here; the CPU oracle keeps dual numbers)
//   Albajar absorption           <- reference src/absorption.jl:10-64,132-235
//   streaming psi-shell deposit  <- reference src/plasma.jl:91-151 (see DESIGN.md for the algorithm and its bound)
//
// Table layout in HBM (one node = one (R,Z) B-spline coefficient position, R fastest), ONE table of 3 double2 per node:
//   tab[node] = { (B_R, B_Z), (B_phi, ln n_e), (ln T_e, psi_N) }
// The first two pairs carry the fields whose gradients every RHS needs; the third (a value-only field and the deposition
// field) is read in a second pass, only by the evaluations that use it. A 4x4 stencil is 4 contiguous runs of 192 B behind
// one address each, read with 16-byte __ldg loads.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace torj {

// reference src/constants.jl:13-26
#define TORJ_C 2.99792458e8
#define TORJ_EPS0 8.8541878128e-12
#define TORJ_E 1.602176634e-19
#define TORJ_ME 9.1093837015e-31

#define TORJ_MAX_GL 64
#ifndef TORJ_ROW_FENCE
#define TORJ_ROW_FENCE 1
#endif
#define TORJ_BESS_K 24
#ifndef TORJ_PAIR
#define TORJ_PAIR 1      // harmonic integrals over node PAIRS (+t, -t): the Bessel series of a pair are evaluated once
#endif
#ifndef TORJ_PAIR_EXP
#define TORJ_PAIR_EXP 0  // 1: the two exponentials of a node pair from ONE exp and one reciprocal (see harmonic_sum). Measured
                         // slower on the headline bundle (129.7 -> 132.0 ms: longer dependent chain per pair); kept for the record, off
#endif
#ifndef TORJ_ROLL_DEPO
#define TORJ_ROLL_DEPO 0  // 1: the three Newton steps of a level crossing as a real loop (smaller hot code)
#endif
#ifndef TORJ_ROLL_STENCIL
#define TORJ_ROLL_STENCIL 0  // 1: the four Z rows of the 4x4 stencil as a real loop (row weights from polynomial tables)
#endif
#ifndef TORJ_PSI_LAZY
#define TORJ_PSI_LAZY 0  // 1: psi_N rides on the stencil only at the stages whose psi is used. Measured SLOWER (134.6 -> 141.0 ms: the
                         // predicate costs more than the 37 DFMA it saves); kept for the record, off
#endif

struct DevTables {
    const double2* __restrict__ A;  // 3 double2 per node: (BR, BZ), (Bphi, ln ne), (ln Te, psi) — one table, so that a stencil row is
                                    // ONE contiguous 192-byte run behind one address (two tables: two address chains per row)
    int nR, nZ, row;                // row = nR + 2 nodes per Z line
    double r0, z0, inv_hr, inv_hz, rlast, zlast;
    double psi_prof_max;
};

struct GLNodes {
    double t[TORJ_MAX_GL];
    double w[TORJ_MAX_GL];
    double sq[TORJ_MAX_GL];  // sqrt(1 - t^2)
    double wp[TORJ_MAX_GL];  // pair weights: w[k] for k < n/2, w[k]/2 for the middle node of an odd rule
    int n;
};

__constant__ GLNodes c_gl;
// Horner coefficients of J_n(z) = (z/2)^n/n! * sum_k c[n][k] y^k, y = z^2/4 : c[n][k] = (-1)^k n! / (k! (n+k)!)
__constant__ double c_bess[5][TORJ_BESS_K];
__constant__ double c_bessd[5][TORJ_BESS_K];  // (n + 2k) c_bess[n][k]: series of z J_n'(z)

struct RayConst {
    double omega;    // 2 pi f
    double cX;       // e^2/(eps0 m_e omega^2)
    double cY;       // e/(m_e omega)
    double w_over_c; // omega / c
    double moded;    // +1 X-mode, -1 O-mode
    double te_min;
    double alpha_floor;
    double icY;       // 1/cY
    double ln_te_min; // log(te_min): the Te gate is applied to the spline value ln Te
    int mode;
    int max_harmonic;
};

__device__ __forceinline__ RayConst make_ray_const(double f, int mode, double te_min, int max_harmonic, double alpha_floor) {
    RayConst rc;
    rc.omega = 2.0 * M_PI * f;
    rc.cX = TORJ_E * TORJ_E / (TORJ_EPS0 * TORJ_ME * rc.omega * rc.omega);
    rc.cY = TORJ_E / (TORJ_ME * rc.omega);
    rc.w_over_c = rc.omega / TORJ_C;
    rc.moded = (double)mode;
    rc.mode = mode;
    rc.te_min = te_min;
    rc.max_harmonic = max_harmonic;
    rc.alpha_floor = alpha_floor;
    rc.icY = 1.0 / rc.cY;
    rc.ln_te_min = log(te_min);
    return rc;
}

struct Counters {
    unsigned int n_acc, n_rej, n_rhs, n_alpha, n_harm, n_prune, n_askip;
};

// ------------------------------------------------------------------------------------------------
// Branch-free FP64 reciprocal / reciprocal square root / square root: MUFU seed (rcp/rsqrt.approx.ftz.f64 work on
// the upper 32 bits of the operand: relative error e0 ~ 2^-20) + ONE third-order correction in FMA arithmetic
// (residual ~e0^3 = 2^-60, i.e. the result is good to the rounding of the last FMA, ~1 ulp). Two Newton steps reach
// the same accuracy with 1 (rcp) / 3 (rsqrt) more instructions and a dependency chain one longer. The library's IEEE
// division and sqrt cost 3-4x the instructions and carry a slow-path branch each; the hot path has ~25 of them per
// RHS. Arguments here are normal, positive (or non-zero for rcp) numbers; zero is handled where it can occur
// (sqrt_fast).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);        // 1/x = r / (1 - e) = r (1 + e + e^2 + ...)
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);    // x^-1/2 = y (1 - e)^-1/2 = y (1 + e/2 + 3 e^2/8 + ...)
    return fma(y, fma(0.375, e, 0.5) * e, y);
}
__device__ __forceinline__ double sqrt_fast(double x) {  // x >= 0
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, 0.5 * y, 0.5);
    y = fma(y, e, y);                        // one Newton step: relative error ~2^-40
    const double s = x * y;
    const double r = fma(fma(-s, s, x), 0.5 * y, s);   // s (1 + d) -> s (1 - d^2)
    return x > 0.0 ? r : 0.0;
}

// ------------------------------------------------------------------------------------------------
// cubic B-spline weights; dw already divided by the grid step
// ------------------------------------------------------------------------------------------------
__constant__ double c_bsk[2] = {1.0 / 6.0, 2.0 / 3.0};  // constant-bank operands: a 64-bit literal costs two MOVs per use
__device__ __forceinline__ void bs_weights(double d, double inv_h, double w[4], double dw[4]) {
    double e = 1.0 - d;
    double d2 = d * d, e2 = e * e;
    w[0] = e2 * e * c_bsk[0];
    w[1] = c_bsk[1] - d2 + 0.5 * d2 * d;
    w[2] = c_bsk[1] - e2 + 0.5 * e2 * e;
    w[3] = d2 * d * c_bsk[0];
    dw[0] = -0.5 * e2 * inv_h;
    dw[1] = (1.5 * d2 - 2.0 * d) * inv_h;
    dw[2] = (2.0 * e - 1.5 * e2) * inv_h;
    dw[3] = 0.5 * d2 * inv_h;
}

__device__ __forceinline__ int bs_locate(double x, double x0, double inv_h, int n, double w[4], double dw[4]) {
    double u = (x - x0) * inv_h + 1.0;
    int i = (int)floor(u);
    i = min(max(i, 1), n - 1);
    bs_weights(u - (double)i, inv_h, w, dw);
    return i - 1;
}

// cubic B-spline weights as polynomials in the cell coordinate d, w_j = sum_k c_bsw[j][k] d^k, and their
// derivatives: for the rolled row loop (TORJ_ROLL_STENCIL), where the row index is a run-time value
__constant__ double c_bsw[4][4] = {{1.0 / 6.0, -0.5, 0.5, -1.0 / 6.0}, {2.0 / 3.0, 0.0, -1.0, 0.5}, {1.0 / 6.0, 0.5, 0.5, -0.5}, {0.0, 0.0, 0.0, 1.0 / 6.0}};
__constant__ double c_bsd[4][3] = {{-0.5, 1.0, -0.5}, {0.0, -2.0, 1.5}, {0.5, 1.0, -1.5}, {0.0, 0.0, 0.5}};

struct Fields {  // value, d/dR, d/dZ
    double BR, BR_R, BR_Z;
    double BZ, BZ_R, BZ_Z;
    double Bp, Bp_R, Bp_Z;
    double L, L_R, L_Z;  // ln n_e
    double lnTe;
    double psi, psi_R, psi_Z;  // psi_N rides along: its coefficients share the (ln T_e, psi_N) loads
};

// all five RHS fields (and optionally psi_N) at a point inside the grid
#ifndef TORJ_TE_PSI_LAZY
#define TORJ_TE_PSI_LAZY 1  // 1: ln T_e and psi_N (the third double2 of a node) in a second pass over the stencil, taken only by the
#endif                      //    evaluations that use one of them — the FSAL / seed / callback stages (psi_N) and those that evaluate
                            //    alpha (T_e): outside the absorbing layer five RHS evaluations of six need neither
#if TORJ_TE_PSI_LAZY && TORJ_ROLL_STENCIL
#error "TORJ_TE_PSI_LAZY needs the row weights of the unrolled stencil"
#endif
template <bool WITH_PSI>
__device__ __forceinline__ void eval_fields_in(const DevTables& T, double R, double Z, Fields& f, bool need_psi = true,
                                               bool need_te = true) {
    double wr[4], dwr[4];
    int br = bs_locate(R, T.r0, T.inv_hr, T.nR, wr, dwr);
#if TORJ_ROLL_STENCIL
    const double uz = (Z - T.z0) * T.inv_hz + 1.0;
    int iz = (int)floor(uz);
    iz = min(max(iz, 1), T.nZ - 1);
    const double dz = uz - (double)iz;
    const int bz = iz - 1;
#else
    double wz[4], dwz[4];
    int bz = bs_locate(Z, T.z0, T.inv_hz, T.nZ, wz, dwz);
#endif
    double v[4] = {0, 0, 0, 0}, vR[4] = {0, 0, 0, 0}, vZ[4] = {0, 0, 0, 0};
    double te = 0.0, ps = 0.0, psR = 0.0, psZ = 0.0;
#if TORJ_ROLL_STENCIL
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) {
#if TORJ_ROLL_STENCIL
        const double wzj = fma(fma(fma(c_bsw[j][3], dz, c_bsw[j][2]), dz, c_bsw[j][1]), dz, c_bsw[j][0]);
        const double dwzj = fma(fma(c_bsd[j][2], dz, c_bsd[j][1]), dz, c_bsd[j][0]) * T.inv_hz;
#else
        const double wzj = wz[j], dwzj = dwz[j];
#endif
        // 32-bit node index (the host refuses grids of 2^31 nodes): one IMAD.WIDE per table and row instead of 64-bit multiplies
        const unsigned node = (unsigned)((bz + j) * T.row + br);
        const double2* pa = T.A + 3 * (size_t)node;
        double a[4] = {0, 0, 0, 0}, aR[4] = {0, 0, 0, 0};
#if !TORJ_TE_PSI_LAZY
        double at = 0.0, ap = 0.0, apR = 0.0;
#endif
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double2 q0 = __ldg(pa + 3 * i);
            double2 q1 = __ldg(pa + 3 * i + 1);
            a[0] = fma(wr[i], q0.x, a[0]); aR[0] = fma(dwr[i], q0.x, aR[0]);
            a[1] = fma(wr[i], q0.y, a[1]); aR[1] = fma(dwr[i], q0.y, aR[1]);
            a[2] = fma(wr[i], q1.x, a[2]); aR[2] = fma(dwr[i], q1.x, aR[2]);
            a[3] = fma(wr[i], q1.y, a[3]); aR[3] = fma(dwr[i], q1.y, aR[3]);
#if !TORJ_TE_PSI_LAZY
            double2 q2 = __ldg(pa + 3 * i + 2);
            at = fma(wr[i], q2.x, at);
            if (WITH_PSI && (!TORJ_PSI_LAZY || need_psi)) { ap = fma(wr[i], q2.y, ap); apR = fma(dwr[i], q2.y, apR); }
#endif
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = fma(wzj, a[q], v[q]);
            vR[q] = fma(wzj, aR[q], vR[q]);
            vZ[q] = fma(dwzj, a[q], vZ[q]);
        }
#if !TORJ_TE_PSI_LAZY
        te = fma(wzj, at, te);
        if (WITH_PSI && (!TORJ_PSI_LAZY || need_psi)) { ps = fma(wzj, ap, ps); psR = fma(wzj, apR, psR); psZ = fma(dwzj, ap, psZ); }
#endif
#if TORJ_ROW_FENCE
        // keep the 12 loads of the next stencil row from being hoisted above this row's arithmetic: the compiler
        // otherwise issues all 32 LDG.128 first and holds 128 registers of loads in flight
        asm volatile("" ::: "memory");
#endif
    }
#if TORJ_TE_PSI_LAZY
    if (need_te || (WITH_PSI && need_psi)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2* pc = T.A + 3 * (size_t)(unsigned)((bz + j) * T.row + br) + 2;
            double at = 0.0, ap = 0.0, apR = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 q2 = __ldg(pc + 3 * i);
                at = fma(wr[i], q2.x, at);
                if (WITH_PSI) { ap = fma(wr[i], q2.y, ap); apR = fma(dwr[i], q2.y, apR); }
            }
            te = fma(wz[j], at, te);
            if (WITH_PSI) { ps = fma(wz[j], ap, ps); psR = fma(wz[j], apR, psR); psZ = fma(dwz[j], ap, psZ); }
        }
    }
#endif
    f.BR = v[0]; f.BR_R = vR[0]; f.BR_Z = vZ[0];
    f.BZ = v[1]; f.BZ_R = vR[1]; f.BZ_Z = vZ[1];
    f.Bp = v[2]; f.Bp_R = vR[2]; f.Bp_Z = vZ[2];
    f.L = v[3]; f.L_R = vR[3]; f.L_Z = vZ[3];
    f.lnTe = te;
    f.psi = ps; f.psi_R = psR; f.psi_Z = psZ;
}

// one field (0..3 from tabA, 4 = lnTe, 5 = psi) with value, gradient and mixed derivative, any position:
// "Line" extrapolation  v = itp(xc) + sum_d (x_d - xc_d) g_d(xc)  and the gradient of that expression
// Out of line and cold (rays are inside the grid except for a few steps at the edge). Table descriptor BY VALUE and
// results BY VALUE: a reference or pointer parameter here forces the caller's copy of the descriptor and its result
// registers into local memory on the HOT path as well (measured: 35 local loads/stores per RHS).
struct Ext3 {
    double v, dR, dZ;
};
__device__ __noinline__ Ext3 eval_field_ext(const DevTables T, int field, double R, double Z) {
    double Rc = fmin(fmax(R, T.r0), T.rlast), Zc = fmin(fmax(Z, T.z0), T.zlast);
    double wr[4], dwr[4], wz[4], dwz[4];
    int br = bs_locate(Rc, T.r0, T.inv_hr, T.nR, wr, dwr);
    int bz = bs_locate(Zc, T.z0, T.inv_hz, T.nZ, wz, dwz);
    double s = 0, sR = 0, sZ = 0, sRZ = 0;
    for (int j = 0; j < 4; ++j) {
        double a = 0, aR = 0;
        for (int i = 0; i < 4; ++i) {
            size_t node = (size_t)(bz + j) * T.row + br + i;
            double c;
            if (field < 4) {
                double2 q = __ldg(T.A + 3 * node + (field >> 1));
                c = (field & 1) ? q.y : q.x;
            } else {
                double2 q = __ldg(T.A + 3 * node + 2);
                c = (field == 4) ? q.x : q.y;
            }
            a = fma(wr[i], c, a);
            aR = fma(dwr[i], c, aR);
        }
        s = fma(wz[j], a, s); sR = fma(wz[j], aR, sR); sZ = fma(dwz[j], a, sZ); sRZ = fma(dwz[j], aR, sRZ);
    }
    bool outR = (Rc != R), outZ = (Zc != Z);
    double v = s, vR = sR, vZ = sZ;
    if (outR) { v += (R - Rc) * sR; if (!outZ) vZ += (R - Rc) * sRZ; }
    if (outZ) { v += (Z - Zc) * sZ; if (!outR) vR += (Z - Zc) * sRZ; }
    Ext3 r;
    r.v = v; r.dR = vR; r.dZ = vZ;
    return r;
}

__device__ __forceinline__ bool inside_grid(const DevTables& T, double R, double Z) {
    return R >= T.r0 && R <= T.rlast && Z >= T.z0 && Z <= T.zlast;
}

template <bool WITH_PSI>
__device__ __forceinline__ void eval_fields(const DevTables& T, double R, double Z, Fields& f, bool need_psi = true,
                                            bool need_te = true) {
    if (inside_grid(T, R, Z)) {
        eval_fields_in<WITH_PSI>(T, R, Z, f, need_psi, need_te);
    } else {
        Ext3 e;
        e = eval_field_ext(T, 0, R, Z); f.BR = e.v; f.BR_R = e.dR; f.BR_Z = e.dZ;
        e = eval_field_ext(T, 1, R, Z); f.BZ = e.v; f.BZ_R = e.dR; f.BZ_Z = e.dZ;
        e = eval_field_ext(T, 2, R, Z); f.Bp = e.v; f.Bp_R = e.dR; f.Bp_Z = e.dZ;
        e = eval_field_ext(T, 3, R, Z); f.L = e.v; f.L_R = e.dR; f.L_Z = e.dZ;
        e = eval_field_ext(T, 4, R, Z); f.lnTe = e.v;
        f.psi = f.psi_R = f.psi_Z = 0.0;
        if (WITH_PSI) { e = eval_field_ext(T, 5, R, Z); f.psi = e.v; f.psi_R = e.dR; f.psi_Z = e.dZ; }
    }
}

// psi_N with gradient (reference src/plasma.jl:61-65 on psi_norm_spline; src/solve.jl:63)
__device__ __forceinline__ void eval_psi(const DevTables& T, double R, double Z, double* psi, double* pR, double* pZ) {
    if (!inside_grid(T, R, Z)) {
        const Ext3 e = eval_field_ext(T, 5, R, Z);
        *psi = e.v; *pR = e.dR; *pZ = e.dZ;
        return;
    }
    double wr[4], dwr[4], wz[4], dwz[4];
    int br = bs_locate(R, T.r0, T.inv_hr, T.nR, wr, dwr);
    int bz = bs_locate(Z, T.z0, T.inv_hz, T.nZ, wz, dwz);
    double s = 0, sR = 0, sZ = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double2* pb = T.A + 3 * ((size_t)(bz + j) * T.row + br) + 2;
        double a = 0, aR = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double c = __ldg(pb + 3 * i).y;
            a = fma(wr[i], c, a);
            aR = fma(dwr[i], c, aR);
        }
        s = fma(wz[j], a, s); sR = fma(wz[j], aR, sR); sZ = fma(dwz[j], a, sZ);
    }
    *psi = s; *pR = sR; *pZ = sZ;
}

__device__ __forceinline__ double psi_at(const DevTables& T, const double x[3]) {
    double p, a, b;
    eval_psi(T, sqrt(x[0] * x[0] + x[1] * x[1]), x[2], &p, &a, &b);
    return p;
}

// ------------------------------------------------------------------------------------------------
// cold dispersion: reference src/dispersion.jl:21-32 with first derivatives
// ------------------------------------------------------------------------------------------------
struct Disp {
    double Ns2, dX, dY, dNp;
};

// iY = 1/Y is passed in (the RHS has 1/|B| already); one rsqrt and one division in total
template <bool DERIV>
__device__ __forceinline__ Disp refractive_index_sq(double X, double Y, double iY, double Np, double moded) {
    Disp o;
    double Np2 = Np * Np, Y2 = Y * Y;
    double om = 1.0 - Np2;
    double iY2 = iY * iY;
    double delta = om * om + 4.0 * Np2 * (1.0 - X) * iY2;
    double rS = rsqrt_fast(delta);
    double S = delta * rS;
    if (delta == 0.0) S = 0.0;
    double D = 2.0 * (-1.0 + X + Y2);
    double iD = rcp_fast(D);
    double G = 1.0 + moded * S + Np2;
    double H = X * Y2;
    o.Ns2 = 1.0 - X + G * iD * H;
    if (DERIV) {
        double hS = 0.5 * rS;  // dS/d. = dDelta/d. * hS
        double S_X = -4.0 * Np2 * iY2 * hS;
        double S_Y = -8.0 * Np2 * (1.0 - X) * iY2 * iY * hS;
        double S_N = (-4.0 * Np * om + 8.0 * Np * (1.0 - X) * iY2) * hS;
        double GH_D2 = G * H * iD * iD;
        o.dX = -1.0 + (moded * S_X * H + G * Y2) * iD - 2.0 * GH_D2;
        o.dY = (moded * S_Y * H + 2.0 * G * X * Y) * iD - 4.0 * Y * GH_D2;
        o.dNp = (moded * S_N + 2.0 * Np) * H * iD;
    } else {
        o.dX = o.dY = o.dNp = 0.0;
    }
    return o;
}

// ------------------------------------------------------------------------------------------------
// Albajar absorption: reference src/absorption.jl:10-64 (polarisation), :132-189 (harmonic integral), :191-226
// ------------------------------------------------------------------------------------------------
struct Pol {  // e = (ex, i*ey, ez) with ex, ey, ez real
    double ex, ey, ez;
};

// reference src/absorption.jl:10-64. iY = 1/Y; st2 = sin^2, ct2 = cos^2 of the propagation angle.
// 3 sqrt/rsqrt + 2 divisions (the reference's sqrt(1/(N sqrt(.))) is one rsqrt here).
__device__ __forceinline__ double abs_Al_N_with_pol_vec(double X, double Y, double iY, double ct, double st, double ct2,
                                                        double st2, int mode, Pol& e) {
    e.ex = e.ey = e.ez = 0.0;
    if (X >= 1.0) return 0.0;
    double Y2 = Y * Y;
    double omX = 1.0 - X;
    double rho = Y2 * st2 * st2 + 4.0 * omX * omX * ct2;
    if (rho < 0.0) return 0.0;
    rho = sqrt_fast(rho);
    double f = (2.0 * omX) * rcp_fast(2.0 * omX - Y2 * st2 - (double)mode * Y * rho);
    double N2 = 1.0 - X * f;
    if (N2 < 0.0) return 0.0;
    double N = sqrt_fast(N2);
    double h = 1.0 - (1.0 - Y2) * f;
    double g = h * iY;
    if (ct2 < 1e-5 || 1.0 - st2 < 1e-5) {
        if (mode > 0) {
            e.ey = rsqrt_fast(N);
            e.ex = -g * e.ey;
        } else {
            e.ez = rsqrt_fast(N);
        }
    } else {
        double den = omX - N2 * st2;
        double iden = rcp_fast(den);
        double hy = g * g;  // (1/Y^2) (1 - (1 - Y^2) f)^2
        double a_in = 1.0 + (omX * N2 * ct2) * (iden * iden) * hy;
        double a_sq = st2 * a_in * a_in;
        double b_in = 1.0 + (omX * iden) * hy;
        double b_sq = ct2 * b_in * b_in;
        double ey = rsqrt_fast(N * sqrt_fast(a_sq + b_sq));
        e.ey = mode > 0 ? ey : -ey;
        e.ex = -g * e.ey;  // i g * (i ey)
        e.ez = -((N2 * st * ct) * iden) * e.ex;
    }
    return N;
}

// exp(x), branch-free, for |x| <= 708 (clamped outside: the callers' arguments are ln n_e ~ 40..46, -ln T_e and
// non-positive exponents whose values below e^-708 cannot matter next to alpha_floor).
// k = rint(x log2 e), r = x - k ln2 (two-part), degree-13 Taylor polynomial on |r| <= ln2/2 (remainder < 4e-18),
// scaled by 2^k built in the exponent field.
__constant__ double c_exp[17];  // log2(e), -ln2_hi, -ln2_lo, 1/13!, 1/12!, ..., 1/2!, 1, 1 (constant-bank operands: a
                                // 64-bit literal costs two extra issue slots per use, a c[][] operand none)
#ifndef TORJ_EXP_TABLE
#define TORJ_EXP_TABLE 1  // 1: exp from a 64-entry table of 2^(j/64) (in shared memory, TORJ_EXP_SMEM) and a degree-5 polynomial instead of degree 13
                          //    (worst 1.3 ulp; 119.6 -> 118.5 ms on the headline bundle, 253 -> 242 ms with alpha_floor = 0)
#endif
__device__ const double d_exp_tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951,
};
#ifndef TORJ_EXP_SMEM
#define TORJ_EXP_SMEM 1  // 1: the 2^(j/64) table is read from shared memory (LDS behind a 32-bit address: 3 instructions instead of
#endif                   //    5 for the 64-bit global address; 99.0 -> 95.6 ms, alpha_floor = 0: 215.9 -> 197.8). EVERY kernel that
                         //    reaches exp_fast must call exp_tab_init() first
__device__ __forceinline__ double* exp_tab_smem() {
    __shared__ double tab[64];
    return tab;
}
__device__ __forceinline__ void exp_tab_init() {  // whole block, before any early return
#if TORJ_EXP_SMEM && TORJ_EXP_TABLE
    for (int j = threadIdx.x; j < 64; j += blockDim.x) exp_tab_smem()[j] = d_exp_tab[j];
    __syncthreads();
#endif
}
__constant__ double c_expt[6] = {92.33248261689366, -0.01083042469326756, -2.9815858269852933e-12, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};
// NEG: the caller's argument cannot exceed +708 (an integrand exp(-mu (gamma - 1) ...)): only the underflow side is clamped
template <bool NEG = false>
__device__ __forceinline__ double exp_fast(double x) {
#if TORJ_EXP_TABLE
    {
        // 22-24 instructions (29 before this arrangement). An FP64 instruction takes ONE operand that is not a register: a
        // constant-bank word or an immediate that fits the high word; everything else costs MOVs — hence mul + add pairs
        // where a fused form would need two constants.
        // |x| > 708 -> +-708 (NaN stays NaN): the high word is replaced, the low word kept (|x| < 708.0005)
        const int hi = __double2hiint(x);
        int hic;
        if (NEG) hic = (x < -708.0) ? (int)0xC0862000 : hi;
        else hic = (fabs(x) > 708.0) ? (0x40862000 | (hi & 0x80000000)) : hi;
        x = __hiloint2double(hic, __double2loint(x));
        // n = rint(64 x / ln2) = 64 k + j in the low word of t (1.5 2^52 trick: no conversion instructions);
        // r = x - n ln2/64 (|r| <= ln2/128) ; exp(x) = 2^k 2^(j/64) (1 + expm1(r))
        const double t = __dadd_rn(__dmul_rn(x, c_expt[0]), 6755399441055744.0);
        const int n = __double2loint(t);
        const double nd = t - 6755399441055744.0;
        double r = fma(nd, c_expt[1], x);
        r = fma(nd, c_expt[2], r);
#if TORJ_EXP_SMEM
        const double tb = exp_tab_smem()[n & 63];
#else
        const double tb = __ldg(d_exp_tab + (n & 63));
#endif
        // expm1(r) = r + r^2 (1/2 + r (1/6 + r (1/24 + r/120)))
        double p = __dadd_rn(__dmul_rn(r, c_expt[3]), c_expt[4]);
        p = fma(p, r, c_expt[5]);
        p = fma(p, r, 0.5);
        const double q = fma(p * r, r, r);
        const double v = fma(tb, q, tb);  // in (0.99, 2): t + t q rounds once (worst 1.3 ulp); the result is normal for |x| <= 708
        return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
    }
#endif
    x = x < -708.0 ? -708.0 : x;  // plain selects: fmin/fmax carry NaN handling that costs ~10 instructions each
    x = x > 708.0 ? 708.0 : x;
    const double kd = rint(x * c_exp[0]);
    double r = fma(kd, c_exp[1], x);
    r = fma(kd, c_exp[2], r);
    double p = c_exp[3];
#pragma unroll
    for (int i = 4; i < 17; ++i) p = fma(p, r, c_exp[i]);
    const int k = (int)kd;  // in [-1022, 1022]
    return p * __hiloint2double((k + 1023) << 20, 0);
}

// J_M(z) and D = z J_M'(z) by K-term Horner series in y = z^2/4 (hz = z/2):
//   J_M = hz^M/M! sum_k c[M][k] y^k ,  z J_M' = hz^M/M! sum_k (M+2k) c[M][k] y^k
// The reference's three Bessel functions enter abs_Al_pol_fact (src/absorption.jl:152-165) only through
//   J_M^2,  J_M (J_{M-1} - J_{M+1}) = 2 J_M J_M',  J_{M-1} J_{M+1} = (M J_M / z)^2 - J_M'^2
// so two series suffice and nothing is divided by z.
// M = 0: order m_rt chosen at run time (harmonics above the third; the one-copy-for-m-2-and-3 variant was slower)
template <int M, int K>
__device__ __forceinline__ void bessel_JD(double hz, double y, double& J, double& D, int m_rt = M) {
    const int m = M ? M : m_rt;
    double sj = c_bess[m][K - 1], sd = c_bessd[m][K - 1];
#pragma unroll
    for (int k = K - 2; k >= 0; --k) {
        sj = fma(sj, y, c_bess[m][k]);
        sd = fma(sd, y, c_bessd[m][k]);
    }
    double p = (m == 2) ? hz * hz * 0.5 : hz * hz * hz * (1.0 / 6.0);
    J = p * sj;
    D = p * sd;
}

struct HarmPre {  // per-call invariants of the harmonic integral
    double mu, Y, iY, N_par, N_perp, spar, ispar;  // spar = sqrt(1 - N_par^2)
    double Axz_sq_ey_sq, Re_Axz_ey, Re_Axz_ez, Re_ey_ez, ey_sq, ez_sq;
    double iNp2;     // 1 / N_perp^2
    double pref;     // mu * a * (mu/2pi)^(3/2), a = 1/(1 + 105/(128 mu^2) + 15/(8 mu))
    double to_alpha; // 2 pi^2/m_0 * X * (omega/c) / Y = 2 pi^2 X (omega/c) / spar : c_abs -> alpha
    double floor_;   // alpha_floor (1/m); harmonics whose rigorous upper bound is below it are skipped
};

struct HarmCoef {  // per-harmonic invariants
    double x_m, e0, e1, k1, k2, k3m2, k3, k4, k5, k6, scale;
    double eb, c2, ae1;  // exp(e0 + |e1|) (the largest exponential on the resonance curve), exp(-2 |e1|), |e1|
};

// reference src/absorption.jl:132-189: sum over the Gauss-Legendre nodes for harmonic M
// LPR > 1 (LPR = 8 or 32 lanes per ray; arguments uniform over the group): lane L of the group takes nodes L, L+LPR, ...;
// butterfly reduction inside the group, so every lane of it returns the same sum.
__device__ __forceinline__ unsigned group_mask(int lpr) {
    return lpr >= 32 ? 0xffffffffu : (((1u << lpr) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(lpr - 1)));
}
template <int M, int K, bool GENERIC, int LPR = 1>
__device__ __forceinline__ double harmonic_sum(const HarmCoef& c, int m_rt = M) {
    double sum = 0.0;
    const int n = c_gl.n;
#if TORJ_PAIR
    // Gauss-Legendre rules are symmetric (torj_abs_init checks it): nodes +-t share sqrt(1 - t^2), hence the Bessel
    // argument, J and D; the polarisation factor splits into a part even in t and a part odd in t.
    const int half = (n + 1) >> 1;
    const double ke = c.k1 - c.k3m2;
#pragma unroll 1
    for (int k = (int)(threadIdx.x & (unsigned)(LPR - 1)); k < half; k += LPR) {
        const double t = c_gl.t[k], sq = c_gl.sq[k];
        const double et = c.e1 * t;
#if TORJ_PAIR_EXP
        // exp(e0 +- et) = eb * exp(+-et - |e1|): the larger factor from one exp, the smaller one as exp(-2|e1|) / larger.
        // (c2 = 0 when it would underflow: the smaller term is then < 1e-150 of the larger one.)
        const double big = exp_fast<true>(fabs(et) - c.ae1);
        const double small = c.c2 * rcp_fast(big);
        const bool pos = et >= 0.0;
        const double exa = c.eb * (pos ? big : small), exb = c.eb * (pos ? small : big);
#else
        const double exa = exp_fast<true>(c.e0 + et), exb = exp_fast<true>(c.e0 - et);
#endif
        const double z = c.x_m * sq;
        double J, D;
        if (GENERIC) {
            J = jn(m_rt, z);
            D = 0.5 * z * (jn(m_rt - 1, z) - jn(m_rt + 1, z));
        } else {
            const double hz = 0.5 * z;
            bessel_JD<M, K>(hz, hz * hz, J, D, m_rt);
        }
        const double J2 = J * J, JD = J * D;
        // (|Axz|^2 + |ey|^2) J^2 + Re(Axz ey*) x_m/m dsq - (z/m)^2 |ey|^2 J_{m-1} J_{m+1} + (xs t)^2 |ez|^2 J^2   [even in t]
        double ev = fma(c.k4 * t, t, ke) * J2;
        ev = fma(c.k2, JD, ev);
        ev = fma(c.k3, D * D, ev);
        const double od = t * fma(c.k5, J2, c.k6 * JD);  // 2 xs Re(Axz ez*) t J^2 + xs Re(ey* ez) t x_m/m dsq           [odd in t]
        sum = fma(c_gl.wp[k], fma(ev + od, exa, (ev - od) * exb), sum);
    }
#else
#pragma unroll 1
    for (int k = (int)(threadIdx.x & (unsigned)(LPR - 1)); k < n; k += LPR) {
        const double t = c_gl.t[k], sq = c_gl.sq[k];
        const double ex = exp_fast(fma(c.e1, t, c.e0));
        const double z = c.x_m * sq;
        double J, D;
        if (GENERIC) {
            J = jn(m_rt, z);
            D = 0.5 * z * (jn(m_rt - 1, z) - jn(m_rt + 1, z));
        } else {
            const double hz = 0.5 * z;
            bessel_JD<M, K>(hz, hz * hz, J, D, m_rt);
        }
        const double J2 = J * J, JD = J * D;
        double pf = c.k1 * J2;                 // (|Axz|^2 + |ey|^2) J^2
        pf = fma(c.k2, JD, pf);                // Re(Axz ey*) x_m/m * dsq,   dsq = (2/x_m) J D
        pf = fma(-c.k3m2, J2, pf);             // -(z/m)^2 |ey|^2 J_{m-1} J_{m+1} = -|ey|^2 (J^2 - D^2/m^2)
        pf = fma(c.k3, D * D, pf);
        const double tJ = t * J2;
        pf = fma(c.k4 * t, tJ, pf);
        pf = fma(c.k5, tJ, pf);
        pf = fma(c.k6 * t, JD, pf);
        sum = fma(c_gl.w[k] * pf, ex, sum);
    }
#endif
    if (LPR > 1) {
        const unsigned gm = group_mask(LPR);
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(gm, sum, off);
    }
    return sum * c.scale;
}

// rarely taken variants kept out of line so the hot code stays small (instruction cache)
template <int M, int LPR = 1>
__device__ __noinline__ double harmonic_sum_large(const HarmCoef c, int m_rt = M) {  // by value: see eval_field_ext
    if (c.x_m <= 6.5) return harmonic_sum<M, 24, false, LPR>(c, m_rt);
    return harmonic_sum<(M ? M : 2), 1, true, LPR>(c, m_rt);
}

// One harmonic's contribution to alpha [1/m] (sign included), or 0 when a rigorous bound shows it is < floor.
// Series length: |term_K| = y^K M!/(K!(M+K)!) (x (M+2K) for D) with y = x_m^2/4: K=12 leaves < 5e-14 for
// x_m <= 3.2 (m=2 layer: x_m ~ 1; m=3 next to it: x_m ~ 2); K=24 for x_m <= 6.5; libm jn() beyond.
// One sqrt per harmonic; everything else is products of the reciprocals prepared in abs_albajar.
// `safe` is cleared unless the bound is below floor * TORJ_SKIP_MARGIN (see abs_albajar).
#ifndef TORJ_SKIP_MARGIN
#define TORJ_SKIP_MARGIN 1e-10
#endif
// M = 0: order given at run time (harmonics above the reference's third; libm jn() throughout).
template <int M, int LPR = 1>
__device__ __forceinline__ double harmonic_alpha(const HarmPre& h, Counters& cnt, bool& safe, int m_rt = M) {
    const int m = M ? M : m_rt;
    const double fm = (double)m, ifm = 1.0 / (double)m;
    const double A = fm * h.Y * h.ispar;  // m / m_0, m_0 = spar / Y
    double q2 = A * A - 1.0;
    q2 = q2 > 0.0 ? q2 : 0.0;
    const double q = sqrt_fast(q2);
    HarmCoef c;
    c.x_m = h.N_perp * h.iY * q;
    // gamma is linear in t on the resonance curve: gamma = (A + N_par q t)/spar  (== sqrt(1 + u_par^2 + u_perp^2))
    const double g0 = A * h.ispar, g1 = h.N_par * q * h.ispar;
    c.e0 = h.mu * (1.0 - g0);
    c.e1 = -h.mu * g1;
    const double xs = c.x_m * h.ispar * ifm;
    c.k1 = h.Axz_sq_ey_sq;
    c.k2 = 2.0 * ifm * h.Re_Axz_ey;
    c.k3m2 = h.ey_sq;
    c.k3 = h.ey_sq * (ifm * ifm);
    c.k4 = xs * xs * h.ez_sq;
    c.k5 = 2.0 * xs * h.Re_Axz_ez;
    c.k6 = 2.0 * ifm * xs * h.Re_ey_ez;
    // (m/(N_perp omega_bar))^2 * c_abs_m * sqrt((m/m0)^2-1) and the conversion to alpha (reference
    // src/absorption.jl:165,186,218-223); the reference's overall minus sign cancels the (-mu) of the integrand
    c.scale = h.pref * (fm * fm * h.Y * h.Y * h.iNp2) * q * h.to_alpha;
    c.ae1 = fabs(c.e1);
    const double emax = c.e0 + c.ae1;
    c.eb = 1.0; c.c2 = 0.0;
    if (TORJ_PAIR_EXP || h.floor_ > 0.0) c.eb = exp_fast(emax);
    if (h.floor_ > 0.0) {
        // |J_n| <= 1, |D| = |z (J_{m-1} - J_{m+1})/2| <= x_m, sum of weights = 2, exponent <= e0 + |e1|
        const double xm = c.x_m;
        const double pmax = fabs(c.k1) + fabs(c.k2) * xm + fabs(c.k3m2) + fabs(c.k3) * xm * xm + fabs(c.k4) + fabs(c.k5)
                            + fabs(c.k6) * xm;
        const double bound = 2.0 * pmax * fabs(c.scale) * (emax < 0.0 ? c.eb : 1.0);
        if (bound < h.floor_) {
            cnt.n_prune++;
            if (!(bound < h.floor_ * TORJ_SKIP_MARGIN)) safe = false;
            return 0.0;
        }
    }
    safe = false;
    cnt.n_harm++;
#if TORJ_PAIR_EXP && TORJ_PAIR
    c.c2 = c.ae1 > 350.0 ? 0.0 : exp_fast(-2.0 * c.ae1);
#endif
    if (M == 0 && m > 3) return harmonic_sum<2, 1, true, LPR>(c, m);
    if (c.x_m <= 3.2) return harmonic_sum<M, 12, false, LPR>(c, m);
    return harmonic_sum_large<M, LPR>(c, m);
}

// Harmonics 4..max_harmonic (torj_options.max_harmonic > 3; the reference stops at 3, src/absorption.jl:199). Cold and
// out of line: the hot loop pays one uniform compare for it.
struct HighHarm {
    double alpha;
    int n_harm, n_prune;
    bool safe;
};
template <int LPR = 1>
__device__ __noinline__ HighHarm harmonics_above_3(const HarmPre& h, int max_harmonic, double m_0, bool safe) {
    HighHarm r;
    Counters cnt = {};
    r.alpha = 0.0;
    for (int m = 4; m <= max_harmonic; ++m) {
        if ((double)m >= m_0) r.alpha += harmonic_alpha<0, LPR>(h, cnt, safe, m);
        else if (!(m_0 - (double)m > 0.02 * m_0)) safe = false;
    }
    r.n_harm = (int)cnt.n_harm; r.n_prune = (int)cnt.n_prune; r.safe = safe;
    return r;
}

// reference src/absorption.jl:191-226 (+ the T_e evaluation of :233). omega enters only through omega/c.
// N2 = |N|^2, iY = 1/Y, lnTe = spline value (T_e = exp(lnTe), reference src/plasma.jl:87-89).
// skip_ok (out): true when every harmonic is either pruned with a 1e10 margin below alpha_floor or more than 2 % away
// (in m_0) from becoming eligible. The integrator then takes alpha = 0 at the inner Runge-Kutta stages of the NEXT
// step, which lie within dtmax (0.1 mm) of this point: the bound would have to grow by 1e10 (its exponent by 23)
// and m_0 = sqrt(1-N_par^2)/Y to move by 2 % over a distance far below the cell size of the spline tables.
// HIGH: the instantiation that also sums harmonics 4..max_harmonic (kept out of the default kernels: even an
// out-of-line call on a never-taken branch cost 6 % there, through the stack copy of HarmPre and its spills).
template <bool HIGH, int LPR = 1>
__device__ __forceinline__ double abs_albajar(const RayConst& rc, double X, double Y, double iY, double N2, double N_par,
                                              double lnTe, Counters& cnt, bool& skip_ok) {
    skip_ok = false;
    if (lnTe < rc.ln_te_min) return 0.0;  // Te < te_min
    cnt.n_alpha++;
    HarmPre h;
    const double Te = exp_fast(lnTe);
    h.mu = (TORJ_ME * TORJ_C * TORJ_C / TORJ_E) * rcp_fast(Te);
    h.Y = Y; h.iY = iY;
    const double iNabs = rsqrt_fast(N2);
    const double ct = N_par * iNabs;
    const double ct2 = ct * ct;
    double st2 = (1.0 - ct) * (1.0 + ct);
    st2 = st2 > 0.0 ? st2 : 0.0;
    const double st = sqrt_fast(st2);                       // sin(acos(ct))
    h.N_perp = (N2 * iNabs) * st;                           // sqrt(N_abs^2 - N_par^2)
    h.N_par = N_par;
    Pol e;
    double N_test = abs_Al_N_with_pol_vec(X, Y, iY, ct, st, ct2, st2, rc.mode, e);
    if (!(N_test > 0.0) || N_test > 1.0) return 0.0;  // also catches NaN
    const double sp2 = 1.0 - N_par * N_par;
    h.ispar = rsqrt_fast(sp2);
    h.spar = sp2 * h.ispar;
    const double m_0 = h.spar * iY;
    const double N_eff = (h.N_perp * N_par) * (h.ispar * h.ispar);
    const double Axz = e.ex + N_eff * e.ez;
    h.ey_sq = e.ey * e.ey;
    h.ez_sq = e.ez * e.ez;
    h.Axz_sq_ey_sq = Axz * Axz + h.ey_sq;
    h.Re_Axz_ey = Axz * e.ey;
    h.Re_Axz_ez = Axz * e.ez;
    h.Re_ey_ez = e.ey * e.ez;
    h.iNp2 = rcp_fast(h.N_perp * h.N_perp);
    const double mu2 = h.mu * h.mu;
    const double a = mu2 * rcp_fast(mu2 + 105.0 / 128.0 + (15.0 / 8.0) * h.mu);  // 1/(1 + 105/(128 mu^2) + 15/(8 mu))
    const double sm = sqrt_fast(h.mu * (0.5 / M_PI));
    h.pref = h.mu * a * (sm * sm * sm);
    h.to_alpha = 2.0 * M_PI * M_PI * X * rc.w_over_c * h.ispar;
    h.floor_ = rc.alpha_floor;
    double alpha = 0.0;
    bool safe = rc.alpha_floor > 0.0;
    if (rc.max_harmonic >= 2) {
        if (2.0 >= m_0) alpha += harmonic_alpha<2, LPR>(h, cnt, safe);
        else if (!(m_0 - 2.0 > 0.02 * m_0)) safe = false;
    }
    if (rc.max_harmonic >= 3) {
        if (3.0 >= m_0) alpha += harmonic_alpha<3, LPR>(h, cnt, safe);
        else if (!(m_0 - 3.0 > 0.02 * m_0)) safe = false;
    }
    if (HIGH && rc.max_harmonic >= 4) {
        const HighHarm r = harmonics_above_3<LPR>(h, rc.max_harmonic, m_0, safe);
        alpha += r.alpha; cnt.n_harm += r.n_harm; cnt.n_prune += r.n_prune; safe = r.safe;
    }
    skip_ok = safe;
    return alpha;
}

// ------------------------------------------------------------------------------------------------
// Hamiltonian RHS: reference src/solve.jl:85-95 (gradΛ!) with analytic gradients
// 4 rsqrt + 1 division + 2 exp outside the absorption coefficient.
// ------------------------------------------------------------------------------------------------
struct PointVals {
    double X, Y, N_par, b[3], Te, Lambda;
};

// What an absorption model evaluated OUTSIDE rhs<> (the warm-plasma model, torj_warm.cuh) needs from the point
struct AlphaIn {
    double X, Y, N2, Np, lnTe, inorm;  // inorm = 1/|dΛ/dN|
};

// WITH_PSI: du[7] = psi_N at the point and du[8] = grad(psi_N) . dx/ds (inputs of the streaming deposition)
// alpha_skip (in/out, optional): on entry true = take alpha = 0 without evaluating it (see abs_albajar); on exit, when
// alpha was evaluated, whether the next step's inner stages may skip it.
// LPR: lanes per ray. LPR > 1: the call is made by the LPR lanes of a group on ONE ray (uniform arguments); the harmonic
// integrals are split over them.
// ain (optional): filled for an absorption model the caller evaluates itself (then WITH_ALPHA = false, du[6] = 0).
template <bool WITH_ALPHA, bool WITH_PSI = false, bool HIGH = false, int LPR = 1>
__device__ __forceinline__ void rhs(const DevTables& T, const RayConst& rc, const double* u, double* du, Counters& cnt,
                                    PointVals* pv = nullptr, bool skip_alpha = false, bool* skip_ok = nullptr,
                                    AlphaIn* ain = nullptr, bool need_psi = true) {
    const double x = u[0], y = u[1], z = u[2], Nx = u[3], Ny = u[4], Nz = u[5];
    const double R2 = fma(x, x, y * y);
    const double iR = rsqrt_fast(R2);
    const double R = R2 * iR;
    const double c = x * iR, s = y * iR;
    Fields f;
    // ln T_e feeds alpha only (and the probes): not needed where alpha is skipped
    eval_fields<WITH_PSI>(T, R, z, f, need_psi, (WITH_ALPHA && !skip_alpha) || ain != nullptr || pv != nullptr);
    const double B2 = f.BR * f.BR + f.Bp * f.Bp + f.BZ * f.BZ;
    const double iB = rsqrt_fast(B2);
    const double Babs = B2 * iB;
    const double NR = Nx * c + Ny * s, Nph = -Nx * s + Ny * c;
    const double NB = NR * f.BR + Nph * f.Bp + Nz * f.BZ;
    const double Np = NB * iB;
    const double bx = (f.BR * c - f.Bp * s) * iB, by = (f.BR * s + f.Bp * c) * iB, bz = f.BZ * iB;
    const double X = rc.cX * exp_fast(f.L);
    const double Y = rc.cY * Babs;
    const double iY = rc.icY * iB;
    Disp d = refractive_index_sq<true>(X, Y, iY, Np, rc.moded);
    // cylindrical gradients of X, Y, N_par at fixed N
    const double X_R = X * f.L_R, X_Z = X * f.L_Z;
    const double Bab_R = (f.BR * f.BR_R + f.Bp * f.Bp_R + f.BZ * f.BZ_R) * iB;
    const double Bab_Z = (f.BR * f.BR_Z + f.Bp * f.Bp_Z + f.BZ * f.BZ_Z) * iB;
    const double Y_R = rc.cY * Bab_R, Y_Z = rc.cY * Bab_Z;
    const double NB_R = NR * f.BR_R + Nph * f.Bp_R + Nz * f.BZ_R;
    const double NB_Z = NR * f.BR_Z + Nph * f.Bp_Z + Nz * f.BZ_Z;
    const double Np_R = (NB_R - Np * Bab_R) * iB;
    const double Np_Z = (NB_Z - Np * Bab_Z) * iB;
    const double Np_ph = (Nph * f.BR - NR * f.Bp) * iB;
    const double gR = -(d.dX * X_R + d.dY * Y_R + d.dNp * Np_R);
    const double gZ = -(d.dX * X_Z + d.dY * Y_Z + d.dNp * Np_Z);
    const double gph = -(d.dNp * Np_ph) * iR;
    const double gx = c * gR - s * gph, gy = s * gR + c * gph, gz = gZ;
    const double hx = 2.0 * Nx - bx * d.dNp, hy = 2.0 * Ny - by * d.dNp, hz = 2.0 * Nz - bz * d.dNp;
    const double inorm = rsqrt_fast(hx * hx + hy * hy + hz * hz);
    du[0] = hx * inorm; du[1] = hy * inorm; du[2] = hz * inorm;
    du[3] = -gx * inorm; du[4] = -gy * inorm; du[5] = -gz * inorm;
    if (WITH_PSI) {
        du[7] = f.psi;
        du[8] = f.psi_R * (c * du[0] + s * du[1]) + f.psi_Z * du[2];
    }
    const double N2 = Nx * Nx + Ny * Ny + Nz * Nz;
    if (WITH_ALPHA) {
        if (skip_alpha) {
            du[6] = 0.0;
            cnt.n_askip++;
        } else {
            bool ok;
            const double alpha = abs_albajar<HIGH, LPR>(rc, X, Y, iY, N2, Np, f.lnTe, cnt, ok);
            du[6] = -u[6] * alpha;
            if (skip_ok) *skip_ok = ok;
        }
    } else {
        du[6] = 0.0;
    }
    if (ain) { ain->X = X; ain->Y = Y; ain->N2 = N2; ain->Np = Np; ain->lnTe = f.lnTe; ain->inorm = inorm; }
    cnt.n_rhs++;
    if (pv) {
        pv->X = X; pv->Y = Y; pv->N_par = Np; pv->b[0] = bx; pv->b[1] = by; pv->b[2] = bz;
        pv->Te = exp(f.lnTe);
        pv->Lambda = N2 - d.Ns2;
    }
}

// ------------------------------------------------------------------------------------------------
// streaming psi-shell deposition (algorithm shared with oracle/deposition.hpp: deposition_streaming)
// ------------------------------------------------------------------------------------------------
struct DepoState {
    int shell;      // psi_grid[shell] <= psi < psi_grid[shell+1]; -1 below all levels, n_psi-1 above all
    int valid;      // a crossing has been seen
    double P_last;  // P at the last crossing
};

__device__ __forceinline__ int locate_shell(const double* __restrict__ g, int n, double psi) {  // upper_bound(psi) - 1
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(g + mid) <= psi) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

__device__ __forceinline__ double hermite(double f0, double f1, double d0, double d1, double h, double th) {
    double t2 = th * th, t3 = t2 * th;
    return (2 * t3 - 3 * t2 + 1) * f0 + (t3 - 2 * t2 + th) * h * d0 + (-2 * t3 + 3 * t2) * f1 + (t3 - t2) * h * d1;
}
__device__ __forceinline__ double hermite_d(double f0, double f1, double d0, double d1, double h, double th) {
    double t2 = th * th;
    return (6 * t2 - 6 * th) * f0 + (3 * t2 - 4 * th + 1) * h * d0 + (-6 * t2 + 6 * th) * f1 + (3 * t2 - 2 * th) * h * d1;
}

// Deposit callback: sink(shell, dP)
template <class Sink>
__device__ __forceinline__ void depo_step(DepoState& st, const double* __restrict__ g, int npsi, double h, double psi_a, double psi_b,
                                          double dpsi_a, double dpsi_b, double P_a, double P_b, double dP_a, double dP_b,
                                          Sink&& sink) {
    while (true) {
        int lvl, nshell;
        if (psi_b < psi_a) {
            lvl = st.shell;
            if (lvl < 0 || !(psi_b < __ldg(g + lvl))) return;
            nshell = lvl - 1;
        } else {
            lvl = st.shell + 1;
            if (lvl > npsi - 1 || !(psi_b >= __ldg(g + lvl))) return;
            nshell = lvl;
        }
        double gl = __ldg(g + lvl);
        double th = (gl - psi_a) / (psi_b - psi_a);
#if TORJ_ROLL_DEPO
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int it = 0; it < 3; ++it) {
            double f = hermite(psi_a, psi_b, dpsi_a, dpsi_b, h, th) - gl;
            double d = hermite_d(psi_a, psi_b, dpsi_a, dpsi_b, h, th);
            if (d != 0.0) th -= f * rcp_fast(d);
            th = fmin(1.0, fmax(0.0, th));
        }
        double Pc = hermite(P_a, P_b, dP_a, dP_b, h, th);
        if (st.valid && st.shell >= 0 && st.shell < npsi - 1) sink(st.shell, fabs(st.P_last - Pc));
        st.P_last = Pc;
        st.valid = 1;
        st.shell = nshell;
    }
}

}  // namespace torj

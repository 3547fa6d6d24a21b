// Warm-plasma (fully relativistic, Larmor-radius expansion) EC damping for sm_100a, FP64:
//   reference src/general_absorption.jl:1328-1337 α  ->  larmornumber :1285, warmdisp :1158, dieltens_maxw_fr :1056,
//   hermitian :646 (the 501-node quadrature with expei :29), antihermitian :951 (ssbi :291, gammln :265, fact :240),
//   tables of set_extv! :8-13 and src/constants.jl:1-4.
// The reference file is dead code (never include()d); its arithmetic is defined, its wiring into gradΛ! is not. The
// wiring here is build-defined: alpha = α(ω, X, Y, |N|, acos(N∥/|N|), Te, 1/|∂Λ/∂N|, mode)[2].
// Dropped: ssbi's debug assertion :314-316 (calls the un-imported `sphericalbesselj`; the series is the modified one).
//
// Mapping. One α call is 501 nodes x (2 llm + 1 <= 7) resonance indices n, one expei each (~0.4 Mflop): far too much
// for the lane that owns the ray, and embarrassingly parallel. So the quadrature is WARP-COOPERATIVE: the 32 lanes of a
// warp take nodes lane, lane+32, ... of ONE ray's integral (arguments broadcast by shuffle), accumulate the 19 sums the
// dielectric tensor needs and butterfly-reduce them; the short serial rest (anti-Hermitian part, tensor assembly,
// fixed-point iteration of the biquadratic) runs per lane for all rays of the warp at once.
// Folding: the reference integrates rr(n,k,m) (46 numbers for llm = 3) and then forms, per order l, six combinations
// ca11..ca33 with constant weights (:1079-1108). The weights do not depend on the node, so the combinations are
// accumulated directly: 6 x 3 sums + rr(0,2,0) = 19 accumulators per lane instead of 46 (registers, shuffles).
#pragma once
#include "torj_device.cuh"

namespace torj {

#define TORJ_WARM_NTV 501
#define TORJ_WARM_H 19  // folded Hermitian sums: [l-1][6] for l = 1..3, then rr(0,2,0)

// expei coefficient tables (reference src/general_absorption.jl:50-142)
__constant__ double cw_a[7] = {1.1669552669734461083368e2, 2.1500672908092918123209e3, 1.5924175980637303639884e4,
                               8.9904972007457256553251e4, 1.5026059476436982420737e5, -1.4815102102575750838086e5,
                               5.0196785185439843791020e0};
__constant__ double cw_b[6] = {4.0205465640027706061433e1, 7.5043163907103936624165e2, 8.1258035174768735759855e3,
                               5.2440529172056355429883e4, 1.8434070063353677359298e5, 2.5666493484897117319268e5};
__constant__ double cw_c[9] = {3.828573121022477169108e-1, 1.107326627786831743809e+1, 7.246689782858597021199e+1,
                               1.700632978311516129328e+2, 1.698106763764238382705e+2, 7.633628843705946890896e+1,
                               1.487967702840464066613e+1, 9.999989642347613068437e-1, 1.737331760720576030932e-8};
__constant__ double cw_d[9] = {8.258160008564488034698e-2, 4.344836335509282083360e+0, 4.662179610356861756812e+1,
                               1.775728186717289799677e+2, 2.953136335677908517423e+2, 2.342573504717625153053e+2,
                               9.021658450529372642314e+1, 1.587964570758947927903e+1, 1.000000000000000000000e+0};
__constant__ double cw_e[10] = {1.3276881505637444622987e+2, 3.5846198743996904308695e+4, 1.7283375773777593926828e+5,
                                2.6181454937205639647381e+5, 1.7503273087497081314708e+5, 5.9346841538837119172356e+4,
                                1.0816852399095915622498e+4, 1.0611777263550331766871e+3, 5.2199632588522572481039e+1,
                                9.9999999999999999087819e-1};
__constant__ double cw_f[10] = {3.9147856245556345627078e+4, 2.5989762083608489777411e+5, 5.5903756210022864003380e+5,
                                5.4616842050691155735758e+5, 2.7858134710520842139357e+5, 7.9231787945279043698718e+4,
                                1.2842808586627297365998e+4, 1.1635769915320848035459e+3, 5.4199632588522559414924e+1,
                                1.0e0};
__constant__ double cw_plg[4] = {-2.4562334077563243311e+01, 2.3642701335621505212e+02, -5.4989956895857911039e+02,
                                 3.5687548468071500413e+02};
__constant__ double cw_qlg[4] = {-3.5553900764052419184e+01, 1.9400230218539473193e+02, -3.3442903192607538956e+02,
                                 1.7843774234035750207e+02};
__constant__ double cw_p[10] = {-1.2963702602474830028590e+01, -1.2831220659262000678155e+03, -1.4287072500197005777376e+04,
                                -1.4299841572091610380064e+06, -3.1398660864247265862050e+05, -3.5377809694431133484800e+08,
                                3.1984354235237738511048e+08,  -2.5301823984599019348858e+10, 1.2177698136199594677580e+10,
                                -2.0829040666802497120940e+11};
__constant__ double cw_q[10] = {7.6886718750000000000000e+01, -5.5648470543369082846819e+03, 1.9418469440759880361415e+05,
                                -4.2648434812177161405483e+06, 6.4698830956576428587653e+07, -7.0108568774215954065376e+08,
                                5.4229617984472955011862e+09, -2.8986272696554495342658e+10, 9.8900934262481749439886e+10,
                                -8.9673749185755048616855e+10};
__constant__ double cw_r[10] = {-2.645677793077147237806e+00, -2.378372882815725244124e+00, -2.421106956980653511550e+01,
                                1.052976392459015155422e+01,  1.945603779539281810439e+01,  -3.015761863840593359165e+01,
                                1.120011024227297451523e+01,  -3.988850730390541057912e+00, 9.565134591978630774217e+00,
                                9.981193787537396413219e-1};
__constant__ double cw_s[9] = {1.598517957704779356479e-4, 4.644185932583286942650e+00, 3.697412299772985940785e+02,
                               -8.791401054875438925029e+00, 7.608194509086645763123e+02, 2.852397548119248700147e+01,
                               4.731097187816050252967e+02, -2.369210235636181001661e+02, 1.249884822712447891440e+00};
__constant__ double cw_p1[10] = {-1.647721172463463140042e+00, -1.860092121726437582253e+01, -1.000641913989284829961e+01,
                                 -2.105740799548040450394e+01, -9.134835699998742552432e-1,  -3.323612579343962284333e+01,
                                 2.495487730402059440626e+01,  2.652575818452799819855e+01,  -1.845086232391278674524e+00,
                                 9.999933106160568739091e-1};
__constant__ double cw_q1[9] = {9.792403599217290296840e+01, 6.403800405352415551324e+01, 5.994932325667407355255e+01,
                                2.538819315630708031713e+02, 4.429413178337928401161e+01, 1.192832423968601006985e+03,
                                1.991004470817742470726e+02, -1.093556195391091143924e+01, 1.001533852045342697818e+00};
__constant__ double cw_p2[10] = {1.75338801265465972390e+02,  -2.23127670777632409550e+02, -1.81949664929868906455e+01,
                                 -2.79798528624305389340e+01, -7.63147701620253630855e+00, -1.52856623636929636839e+01,
                                 -7.06810977895029358836e+00, -5.00006640413131002475e+00, -3.00000000320981265753e+00,
                                 1.00000000000000485503e+00};
__constant__ double cw_q2[9] = {3.97845977167414720840e+04, 3.97277109100414518365e+00, 1.37790390235747998793e+02,
                                1.17179220502086455287e+02, 7.04831847180424675988e+01, -1.20187763547154743238e+01,
                                -7.99243595776339741065e+00, -2.99999894040324959612e+00, 1.99999999999048104167e+00};

// Combination weights of dieltens_maxw_fr (:1079-1108) for the pair (is, l), 0 <= is <= l <= 5, uploaded by the host:
//   [0] is^2 asl   [1] is l asl   [2] bsl   [3] is asl   [4] l asl   [5] asl
//   asl = (-1)^(l-is) / ((is+l)! (l-is)!),  bsl = asl (is^2 + 2 (l-is)(l-1)(l+is)/(2l-1))
// cw_fal[l] = -0.25^l (2l)!/(l!)^2 (the yg^(2(l-1)) divisor is applied at run time); cw_fact[m] = m!;
// cw_igam[m] = 1/exp(gammln(m + 1.5)) with the reference's 6-term Lanczos gammln (:265-283), m = 0..7.
__constant__ double cw_comb[6][6][6];
__constant__ double cw_combs[7][4][6];  // the same for the quadrature, indexed [n + 3][l], with the sign of n folded into [1] and [3]
__constant__ double cw_fal[6];
__constant__ double cw_fact[8];
__constant__ double cw_igam[8];

// reference src/general_absorption.jl:29-232: exp(-x) Ei(x). Divisions through rcp_fast (<= 1 ulp), exp through
// exp_fast; log is CUDA's (<= 1 ulp).
__device__ __forceinline__ double expei_dev(double x) {
    if (x == 0.0) return -1.79e+308;
    if (x < 0.0) {
        const double y = -x;
        if (y <= 1.0) {
            double sump = fma(cw_a[6], y, cw_a[0]), sumq = y + cw_b[0];
#pragma unroll
            for (int i = 1; i < 6; ++i) { sump = fma(sump, y, cw_a[i]); sumq = fma(sumq, y, cw_b[i]); }
            return (log(y) - sump * rcp_fast(sumq)) * exp_fast(y);
        } else if (y <= 4.0) {
            const double w = rcp_fast(y);
            double sump = cw_c[0], sumq = cw_d[0];
#pragma unroll
            for (int i = 1; i < 9; ++i) { sump = fma(sump, w, cw_c[i]); sumq = fma(sumq, w, cw_d[i]); }
            return -sump * rcp_fast(sumq);
        } else {
            const double w = rcp_fast(y);
            double sump = cw_e[0], sumq = cw_f[0];
#pragma unroll
            for (int i = 1; i < 10; ++i) { sump = fma(sump, w, cw_e[i]); sumq = fma(sumq, w, cw_f[i]); }
            return w * (w * sump * rcp_fast(sumq) - 1.0);
        }
    } else if (x < 6.0) {
        const double t = (x + x) / 3.0 - 2.0;
        double pm = 0.0, pc = cw_p[0], qm = 0.0, qc = cw_q[0];  // px[i-1], px[i]
#pragma unroll
        for (int i = 1; i < 9; ++i) {
            const double pn = fma(t, pc, -pm) + cw_p[i], qn = fma(t, qc, -qm) + cw_q[i];
            pm = pc; pc = pn; qm = qc; qc = qn;
        }
        const double sump = 0.5 * t * pc - pm + cw_p[9];
        const double sumq = 0.5 * t * qc - qm + cw_q[9];
        const double frac = sump * rcp_fast(sumq);
        const double x0 = 0.37250741078136663466;
        const double xmx0 = (x - 381.5 / 1024.0) - (-5.1182968633365538008e-5);
        if (fabs(xmx0) >= 0.037) return exp_fast(-x) * (log(x / x0) + xmx0 * frac);
        const double y = xmx0 * rcp_fast(x + x0), ysq = y * y;
        double sp = cw_plg[0], sq = ysq + cw_qlg[0];
#pragma unroll
        for (int i = 1; i < 4; ++i) { sp = fma(sp, ysq, cw_plg[i]); sq = fma(sq, ysq, cw_qlg[i]); }
        return exp_fast(-x) * (sp * rcp_fast(sq * (x + x0)) + frac) * xmx0;
    } else if (x < 12.0) {
        double frac = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) frac = cw_s[i] * rcp_fast(cw_r[i] + x + frac);
        return (cw_r[9] + frac) * rcp_fast(x);
    } else if (x <= 24.0) {
        double frac = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) frac = cw_q1[i] * rcp_fast(cw_p1[i] + x + frac);
        return (cw_p1[9] + frac) * rcp_fast(x);
    } else {
        const double y = rcp_fast(x);
        double frac = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) frac = cw_q2[i] * rcp_fast(cw_p2[i] + x + frac);
        frac = cw_p2[9] + frac;
        return y + y * y * frac;
    }
}

// reference src/general_absorption.jl:1285-1326
__device__ __forceinline__ int larmornumber_dev(double yg, double npl, double mu) {
    const double dnl = 1.0 - npl * npl;
    int imax = 1;
    int nharm = (int)floor(1.0 / yg);
    if (nharm * yg < 1.0) nharm = nharm + 1;
    for (;;) {
        const double ygn = nharm * yg;
        const double rdu2 = ygn * ygn - dnl;
        const double gg = (ygn - sqrt(npl * npl * rdu2)) / dnl;
        const double argexp = mu * (gg - 1.0);
        if (argexp > 15.0) break;
        nharm = nharm + 1;
        imax = imax + 1;
        if (imax > 100) { nharm = (int)floor(yg); break; }
    }
    return nharm;
}

// reference src/general_absorption.jl:646-714 (iwarm = 3), executed by a FULL WARP with warp-uniform arguments: lane L
// integrates nodes L, L+32, ...; on return every lane holds the complete folded sums
//   H[6 (l-1) + q], q = 0..5: sum_n comb[|n|][l][q] sgn_q(n) rr(n, k_q, l)   (k = 0,0,0,1,1,2; sgn only for q = 1, 3)
//   H[18] = rr(0,2,0)
// tab[i] = (t_i, exp(-t_i^2) dt): the tables of set_extv! (:8-13).
// Out of line on purpose: inside the trace kernel the quadrature loop would otherwise be register-allocated together
// with the whole integrator state (1.1 KB of spills through every one of its ~13 000 instructions); as a call the
// caller's state is saved once around it.
__device__ __noinline__ void warm_hermitian_warp(const double2* __restrict__ tab, double yg, double anpl, double amu, int llm,
                                                 double* Hout) {
    const unsigned lane = threadIdx.x & 31u;
    double H[TORJ_WARM_H];
#pragma unroll
    for (int q = 0; q < TORJ_WARM_H; ++q) H[q] = 0.0;
    const double iamu = rcp_fast(amu);
    const double cmxw = 1.0 + 15.0 / 8.0 * iamu + 105.0 / 128.0 * iamu * iamu;
    const double cr = -amu * amu * rcp_fast(1.7724538509055160272981674833411 * cmxw);
    const double bth2 = 2.0 * iamu, bth = sqrt_fast(bth2);
    const double iamu2 = iamu * iamu, iamu4 = iamu2 * iamu2, iamu6 = iamu4 * iamu2;
#pragma unroll 1
    for (int i = lane; i < TORJ_WARM_NTV; i += 32) {
        const double2 te = __ldg(tab + i);
        const double t = te.x;
        const double g2 = fma(t * t, 0.5 * iamu, 1.0);
        const double irxt = rsqrt_fast(g2);
        const double rxt = g2 * irxt;
        const double x = t * rxt;
        const double upl2 = bth2 * (x * x);
        const double upl = bth * x;
        const double gx = fma(t * t, iamu, 1.0);
        const double exdx = cr * te.y * gx * irxt;
#pragma unroll 1
        for (int n = -llm; n <= llm; ++n) {
            const int nn = n < 0 ? -n : n;
            const double gr = fma(anpl, upl, (double)n * yg);
            const double zm = -amu * (gx - gr);
            const double s = amu * (gx + gr);
            const double zm2 = zm * zm;
            const double fe0m = expei_dev(zm);
            const double zf = zm * fe0m;
            if (nn == 0) H[18] = fma(-exdx * fe0m, upl2, H[18]);
            // ffe for m = max(|n|, 1) .. llm (:691-697); n is warp-uniform, so these branches do not diverge
#pragma unroll
            for (int l = 1; l <= 3; ++l) {
                if (l >= nn && l <= llm) {
                    double ffe;
                    if (l == 1) ffe = fma(s, 1.0 - zf, 1.0) * iamu2;
                    else if (l == 2) ffe = (6.0 - 2.0 * zm + 4.0 * s + s * s * (1.0 + zm - zm * zf)) * iamu4;
                    else ffe = (18.0 * s * (s + 4.0 - zm) + 6.0 * (20.0 - 8.0 * zm + zm2) + s * s * s * (2.0 + zm + zm2 - zm2 * zf)) * iamu6;
                    const double E = exdx * ffe;
                    const double E1 = E * upl, E2 = E * upl2;
                    const double* cb = cw_combs[n + 3][l];
                    double* h = H + 6 * (l - 1);
                    h[0] = fma(cb[0], E, h[0]);
                    h[1] = fma(cb[1], E, h[1]);
                    h[2] = fma(cb[2], E, h[2]);
                    h[3] = fma(cb[3], E1, h[3]);
                    h[4] = fma(cb[4], E1, h[4]);
                    h[5] = fma(cb[5], E2, h[5]);
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int q = 0; q < TORJ_WARM_H; ++q) H[q] += __shfl_xor_sync(0xffffffffu, H[q], off);
#pragma unroll
    for (int q = 0; q < TORJ_WARM_H; ++q) Hout[q] = H[q];
}

struct Cx {
    double re, im;
};
__device__ __forceinline__ Cx cx(double a, double b) { Cx z; z.re = a; z.im = b; return z; }
__device__ __forceinline__ Cx operator+(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ Cx operator-(Cx a, Cx b) { return cx(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ Cx operator-(Cx a) { return cx(-a.re, -a.im); }
__device__ __forceinline__ Cx operator*(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ Cx operator*(double s, Cx a) { return cx(s * a.re, s * a.im); }
__device__ __forceinline__ Cx operator+(Cx a, double s) { return cx(a.re + s, a.im); }
__device__ __forceinline__ Cx operator-(Cx a, double s) { return cx(a.re - s, a.im); }
__device__ __forceinline__ Cx operator-(double s, Cx a) { return cx(s - a.re, -a.im); }
__device__ __forceinline__ Cx cdiv(Cx a, Cx b) {
    const double d = 1.0 / (b.re * b.re + b.im * b.im);
    return cx((a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d);
}
__device__ __forceinline__ double cabs(Cx a) { return hypot(a.re, a.im); }
__device__ __forceinline__ Cx csqrt(Cx z) {  // principal branch
    const double r = hypot(z.re, z.im);
    if (r == 0.0) return cx(0.0, z.im);
    if (z.re >= 0.0) {
        const double t = sqrt(0.5 * (r + z.re));
        return cx(t, z.im / (2.0 * t));
    }
    const double t = sqrt(0.5 * (r - z.re));
    return cx(fabs(z.im) / (2.0 * t), z.im < 0.0 ? -t : t);
}

struct WarmOut {
    Cx anpr;  // complex N_perp
    int ierr, iters;
};

// Anti-Hermitian part (:951-1043), tensor assembly (:1056-1134) and the fixed-point iteration of the biquadratic for
// N_perp^2 (:1158-1234), per lane. H = folded Hermitian sums of warm_hermitian_warp.
__device__ __noinline__ WarmOut warm_solve(const double* H, double xg, double yg, double anpl, double amu, double anprc,
                                           int sox, int lrm) {
    // ---- folded anti-Hermitian sums I[l-1][q], q as in H (imaginary parts of ca11, ca12, ca22, ca13, ca23, ca33)
    double I[5][6];
#pragma unroll
    for (int l = 0; l < 5; ++l)
#pragma unroll
        for (int q = 0; q < 6; ++q) I[l][q] = 0.0;
    const double dnl = 1.0 - anpl * anpl;
    const double cmu = anpl * amu;
    const double cmxw = 1.0 + 15.0 / (8.0 * amu) + 105.0 / (128.0 * (amu * amu));
    const double ci = sqrt(2.0 * M_PI * amu) * (amu * amu) / cmxw;
    for (int n = 1; n <= lrm; ++n) {
        const double ygn = n * yg;
        const double rdu2 = ygn * ygn - dnl;
        if (!(rdu2 > 0.0)) continue;
        const double rdu = sqrt(rdu2);
        const double du = rdu / dnl;
        const double ub = anpl * ygn / dnl;
        const double aa = amu * anpl * du;
        double dm = 1.0;  // dnl^m
        if (fabs(aa) > 5.0) {
            const double up = ub + du, um = ub - du;
            const double gp = anpl * up + ygn, gm = anpl * um + ygn;
            const double xp = up + 1.0 / cmu, xm = um + 1.0 / cmu;
            const double eem = exp(-amu * (gm - 1.0)), eep = exp(-amu * (gp - 1.0));
            double fi0p0 = -1.0 / cmu, fi1p0 = -xp / cmu, fi2p0 = -(1.0 / (cmu * cmu) + xp * xp) / cmu;
            double fi0m0 = -1.0 / cmu, fi1m0 = -xm / cmu, fi2m0 = -(1.0 / (cmu * cmu) + xm * xm) / cmu;
            for (int m = 1; m <= lrm; ++m) {
                dm *= dnl;
                const double fi0p1 = -2.0 * m * (fi1p0 - ub * fi0p0) / cmu;
                const double fi0m1 = -2.0 * m * (fi1m0 - ub * fi0m0) / cmu;
                const double fi1p1 = -((1.0 + 2 * m) * fi2p0 - 2.0 * (m + 1) * ub * fi1p0 + up * um * fi0p0) / cmu;
                const double fi1m1 = -((1.0 + 2 * m) * fi2m0 - 2.0 * (m + 1) * ub * fi1m0 + up * um * fi0m0) / cmu;
                const double fi2p1 = (2.0 * (1 + m) * fi1p1 - 2.0 * m * (ub * fi2p0 - up * um * fi1p0)) / cmu;
                const double fi2m1 = (2.0 * (1 + m) * fi1m1 - 2.0 * m * (ub * fi2m0 - up * um * fi1m0)) / cmu;
                if (m >= n) {
                    const double r0 = 0.5 * ci * dm * (fi0p1 * eep - fi0m1 * eem);
                    const double r1 = 0.5 * ci * dm * (fi1p1 * eep - fi1m1 * eem);
                    const double r2 = 0.5 * ci * dm * (fi2p1 * eep - fi2m1 * eem);
                    const double* cb = cw_comb[n][m];
                    I[m - 1][0] += cb[0] * r0; I[m - 1][1] += cb[1] * r0; I[m - 1][2] += cb[2] * r0;
                    I[m - 1][3] += cb[3] * r1; I[m - 1][4] += cb[4] * r1; I[m - 1][5] += cb[5] * r2;
                }
                fi0p0 = fi0p1; fi1p0 = fi1p1; fi2p0 = fi2p1;
                fi0m0 = fi0m1; fi1m0 = fi1m1; fi2m0 = fi2m1;
            }
        } else {
            const double ee = exp(-amu * (ygn - 1.0 + anpl * ub));
            // ssbi(aa, n, lrm) (:291-320): orders n..lrm+2, series sum_k (aa^2/4)^k / (k! Gamma(m+k+3/2))
            double fs[8];
            const double z2q = 0.25 * (aa * aa);
            for (int m = n; m <= lrm + 2; ++m) {
                double c0 = cw_igam[m];
                double sbi = c0;
                for (int k = 1; k <= 50; ++k) {  // c1 = c0 z2q / (m + k + 1/2) / k ; stop when c1 / sbi < 1e-10
                    const double c1 = c0 * z2q * rcp_fast(((m + k) + 0.5) * (double)k);
                    sbi = sbi + c1;
                    if (c1 < 1.0e-10 * sbi) break;
                    c0 = c1;
                }
                fs[m - n] = sbi;
            }
            double dup = du;  // du^(2m+1), built up from m = 0
            for (int m = 1; m <= lrm; ++m) {
                dm *= dnl;
                dup *= du * du;
                if (m < n) continue;
                const double cm = 1.7724538509055160272981674833411 * cw_fact[m] * dup;
                const double cim = 0.5 * ci * dm;
                const int mm = m - n;
                const double fi0m = cm * fs[mm];
                const double fi1m = -0.5 * aa * cm * fs[mm + 1];
                const double fi2m = 0.5 * cm * (fs[mm + 1] + 0.5 * aa * aa * fs[mm + 2]);
                const double r0 = cim * ee * fi0m;
                const double r1 = cim * ee * (du * fi1m + ub * fi0m);
                const double r2 = cim * ee * (du * du * fi2m + 2.0 * du * ub * fi1m + ub * ub * fi0m);
                const double* cb = cw_comb[n][m];
                I[m - 1][0] += cb[0] * r0; I[m - 1][1] += cb[1] * r0; I[m - 1][2] += cb[2] * r0;
                I[m - 1][3] += cb[3] * r1; I[m - 1][4] += cb[4] * r1; I[m - 1][5] += cb[5] * r2;
            }
        }
    }
    // ---- epsl(l) (:1110-1131): e11, e12, e22, a13, a23, a33 per order
    Cx e11[5], e12[5], e22[5], a13[5], a23[5], a33[5];
    const double iyg = 1.0 / yg, iyg2 = iyg * iyg;
    double ypw = 1.0;  // yg^(2 (l-1))
    for (int l = 1; l <= lrm; ++l) {
        const double fal = cw_fal[l] / ypw;
        ypw *= yg * yg;
        double h[6] = {0, 0, 0, 0, 0, 0};
        if (l <= 3) {
#pragma unroll
            for (int q = 0; q < 6; ++q) h[q] = H[6 * (l - 1) + q];
        }
        const double* im = I[l - 1];
        const double xf = xg * fal;
        e11[l - 1] = cx(-xf * h[0], -xf * im[0]);
        e12[l - 1] = cx(-xf * im[1], xf * h[1]);                          //  i xg ca12 fal
        e22[l - 1] = cx(-xf * h[2], -xf * im[2]);
        a13[l - 1] = cx(-xf * h[3] * iyg, -xf * im[3] * iyg);
        a23[l - 1] = cx(xf * im[4] * iyg, -xf * h[4] * iyg);              // -i xg ca23 fal
        a33[l - 1] = cx(-xf * h[5] * iyg2, -xf * im[5] * iyg2);
    }
    e11[0].re += 1.0;
    e22[0].re += 1.0;
    const double e330 = 1.0 + xg * H[18];
    // ---- fixed-point iteration (:1172-1226)
    WarmOut out;
    out.ierr = 0; out.iters = 0;
    double errnpr = 1.0;
    Cx anpr2a = cx(anprc * anprc, 0.0), anpr2 = cx(0.0, 0.0);
    const double anpl2 = anpl * anpl;
    Cx s11, s12, s22, s13, s23, s33;
    for (int i = 1; i <= 100; ++i) {
        // sum_l epsl(l) anpr2a^(l-1)
        s11 = e11[lrm - 1]; s12 = e12[lrm - 1]; s22 = e22[lrm - 1]; s13 = a13[lrm - 1]; s23 = a23[lrm - 1]; s33 = a33[lrm - 1];
        for (int l = lrm - 1; l >= 1; --l) {
            s11 = s11 * anpr2a + e11[l - 1]; s12 = s12 * anpr2a + e12[l - 1]; s22 = s22 * anpr2a + e22[l - 1];
            s13 = s13 * anpr2a + a13[l - 1]; s23 = s23 * anpr2a + a23[l - 1]; s33 = s33 * anpr2a + a33[l - 1];
        }
        if (i > 2 && errnpr < 1.0e-4) break;
        out.iters = i;
        const Cx a31 = s13, a32 = -s23;
        const Cx e11m = s11 - anpl2, e22m = s22 - anpl2, om33 = 1.0 - s33;
        const Cx p13 = s13 + anpl, p31 = a31 + anpl;
        const Cx cc4 = e11m * om33 + p13 * p31;
        const Cx cc2 = -(s12 * s12) * om33 - a32 * s12 * p13 + s23 * s12 * p31 - (s23 * a32 + e330 + e22m * om33) * e11m -
                       p13 * p31 * e22m;
        const Cx cc0 = e330 * (e11m * e22m + s12 * s12);
        const Cx rr = cc2 * cc2 - 4.0 * (cc0 * cc4);
        double s;
        if (yg > 1.0) {
            s = (double)sox;
            if (rr.im <= 0.0) s = -s;
        } else {
            s = (double)(-sox);
            if (rr.re <= 0.0 && rr.im >= 0.0) s = -s;
        }
        anpr2 = cdiv(-cc2 + s * csqrt(rr), 2.0 * cc4);
        errnpr = fabs(1.0 - cabs(anpr2) / cabs(anpr2a));
        anpr2a = anpr2;
    }
    if (anpr2.re < 0.0 && anpr2.im < 0.0) {
        anpr2 = cx(0.0, 0.0);
        out.ierr = 99;
    }
    out.anpr = csqrt(anpr2);
    return out;
}

// Per-lane preliminaries of α (:1329-1333) and the gate of the quadrature.
struct WarmPrep {
    double yg, anpl, amu, anprc, sinth;
    int lrm;
};
// returns false when alpha = 0 without any work (Te gate) or alpha is NaN by construction (lrm < 1)
__device__ __noinline__ bool warm_prepare(const RayConst& rc, const AlphaIn& a, WarmPrep& p, double& alpha_out, bool& gated,
                                          bool& safe) {
    gated = false; safe = false;
    alpha_out = 0.0;
    if (a.lnTe < rc.ln_te_min) { safe = rc.alpha_floor > 0.0; return false; }
    const double te = exp(a.lnTe);
    const double N_r = sqrt(a.N2);
    const double theta = acos(a.Np / N_r);
    const double N_par = N_r * cos(theta);
    p.anprc = sqrt(N_r * N_r - N_par * N_par);
    p.sinth = sin(theta);
    p.amu = TORJ_ME * (TORJ_C * TORJ_C) / (te * TORJ_E);
    p.yg = a.Y; p.anpl = N_par;
    const int nharm = larmornumber_dev(a.Y, N_par, p.amu);
    p.lrm = nharm < 5 ? nharm : 5;
    if (p.lrm < 1) { alpha_out = nan(""); return false; }  // the reference would throw (BoundsError on a 3x3x0 tensor)
    if (rc.alpha_floor > 0.0) {
        // Gate: the damping comes from the anti-Hermitian part, every term of which carries exp(-mu (gamma_n - 1)) with
        // gamma_n >= gg_n = (n Y - |N_par| sqrt((n Y)^2 - 1 + N_par^2)) / (1 - N_par^2), the smallest resonant energy of
        // harmonic n (the quantity larmornumber thresholds at 15). With ci = sqrt(2 pi mu) mu^2 as the common prefactor and
        // 1e6 for the polynomial factors and the 1/|dD/dN_perp^2| amplification, alpha < 2 (w/c) v_g X ci 1e6 exp(-argmin).
        // The quadrature is skipped when that is below alpha_floor; `safe` (inner-stage skip) asks 1e10 more.
        const double dnl = 1.0 - N_par * N_par;
        double argmin = 1e300;
        for (int n = 1; n <= p.lrm; ++n) {
            const double ygn = n * a.Y, rdu2 = ygn * ygn - dnl;
            if (rdu2 > 0.0) {
                const double gg = (ygn - sqrt(N_par * N_par * rdu2)) / dnl;
                argmin = fmin(argmin, p.amu * (gg - 1.0));
            }
        }
        const double pref = 2.0 * rc.w_over_c * a.inorm * a.X * sqrt(2.0 * M_PI * p.amu) * p.amu * p.amu * 1e6;
        const double lim = log(pref / rc.alpha_floor);
        if (argmin > lim) {
            gated = true;
            safe = argmin > lim + 23.0;
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ double warm_alpha_finish(const RayConst& rc, const AlphaIn& a, const WarmPrep& p, const double* H,
                                                    double* N_warm = nullptr, int* ierr = nullptr, int* iters = nullptr) {
    const WarmOut w = warm_solve(H, a.X, p.yg, p.anpl, p.amu, p.anprc, rc.mode, p.lrm);
    if (N_warm) *N_warm = w.anpr.re / p.sinth;
    if (ierr) *ierr = w.ierr;
    if (iters) *iters = w.iters;
    const Cx n2 = w.anpr * w.anpr;
    return 2.0 * n2.im * rc.w_over_c * a.inorm;
}

// alpha of the warm model for every lane of a warp that `want`s it. Must be called by all 32 lanes, converged.
//   COOP = false (one ray per lane): the quadratures of the wanting lanes are done one after the other by the whole warp
//     (arguments broadcast from the owner, sums parked in the owner's shared-memory slots stash[q * stride]); then all
//     owners run the serial rest at once.
//   COOP = true (the warp holds ONE ray, arguments warp-uniform): one quadrature, every lane keeps the sums.
// safe (out): this evaluation allows alpha = 0 at the inner stages of the next step (see abs_albajar's skip_ok).
template <bool COOP>
__device__ __forceinline__ double warm_alpha_warp(const double2* __restrict__ tab, const RayConst& rc, const AlphaIn& ain,
                                                  bool want, double* stash, int stride, Counters& cnt, bool& safe) {
    const unsigned FULLM = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    WarmPrep wp;
    wp.yg = 0.5; wp.anpl = 0.0; wp.amu = 100.0; wp.anprc = 1.0; wp.sinth = 1.0; wp.lrm = 1;
    double alpha = 0.0;
    bool gated = false, needq = false;
    safe = false;
    if (want) {
        needq = warm_prepare(rc, ain, wp, alpha, gated, safe);
        if (!(ain.lnTe < rc.ln_te_min)) cnt.n_alpha++;
        if (gated) cnt.n_prune += 2 * (wp.lrm < 3 ? wp.lrm : 3) + 1;
    }
    const int llm = wp.lrm < 3 ? wp.lrm : 3;
    if (COOP) {
        if (needq) {
            double H[TORJ_WARM_H];
            warm_hermitian_warp(tab, wp.yg, wp.anpl, wp.amu, llm, H);
            alpha = warm_alpha_finish(rc, ain, wp, H);
        }
    } else {
        unsigned m = __ballot_sync(FULLM, needq);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double yg = __shfl_sync(FULLM, wp.yg, src), anpl = __shfl_sync(FULLM, wp.anpl, src);
            const double amu = __shfl_sync(FULLM, wp.amu, src);
            const int l3 = __shfl_sync(FULLM, llm, src);
            double H[TORJ_WARM_H];
            warm_hermitian_warp(tab, yg, anpl, amu, l3, H);
            if ((int)lane == src) {
#pragma unroll
                for (int q = 0; q < TORJ_WARM_H; ++q) stash[q * stride] = H[q];
            }
        }
        if (needq) {
            double H[TORJ_WARM_H];
#pragma unroll
            for (int q = 0; q < TORJ_WARM_H; ++q) H[q] = stash[q * stride];
            alpha = warm_alpha_finish(rc, ain, wp, H);
        }
    }
    if (needq) cnt.n_harm += 2 * llm + 1;
    return alpha;
}

}  // namespace torj

// TEST INFRASTRUCTURE — CPU oracle (see dual.hpp header). Not part of the product.
//
// Warm-plasma (fully / weakly relativistic) dispersion and EC absorption, restating line by line
//   reference src/general_absorption.jl:8-13      set_extv!           (tables of src/constants.jl:1-4,10-11)
//   reference src/general_absorption.jl:29-232    expei               exp(-x) Ei(x)  (W. J. Cody's CALCEI, int = 3)
//   reference src/general_absorption.jl:240-257   fact
//   reference src/general_absorption.jl:265-283   gammln              (6-term Lanczos: accurate to ~2e-10 only)
//   reference src/general_absorption.jl:291-320   ssbi                series sum_k (z^2/4)^k / (k! Gamma(m+k+3/2))
//   reference src/general_absorption.jl:345-465   zetac               plasma dispersion function (ACM Algorithm 680)
//   reference src/general_absorption.jl:473-561   fsup
//   reference src/general_absorption.jl:573-638   dieltens_maxw_wr    (iwarm = 1)
//   reference src/general_absorption.jl:646-714   hermitian           (the quadrature, iwarm > 2)
//   reference src/general_absorption.jl:951-1043  antihermitian
//   reference src/general_absorption.jl:1056-1134 dieltens_maxw_fr
//   reference src/general_absorption.jl:1158-1267 warmdisp
//   reference src/general_absorption.jl:1285-1326 larmornumber
//   reference src/general_absorption.jl:1328-1337 α
//
// What is deliberately different from the file as written, and why:
//   * ssbi's debug assertion (:314-316) is dropped: it calls `sphericalbesselj`, which the module never imports
//     (src/TorJ.jl:11 imports only besselj), and compares the series with it, although the series is the MODIFIED
//     spherical Bessel function: (2/z)^m (2/sqrt(pi)) i_m(z). The arithmetic of the series itself is kept.
//   * set_extv!() is never called by the reference; here the tables are filled on first use.
//   * hermitian's closed-form expansions for iwarm <= 2 (:716-941) are not restated: α hard-codes iwarm = 3 (:1334) and
//     warmdisp reaches hermitian only through dieltens_maxw_fr, so they are unreachable from the path. hermitian()
//     here serves iwarm > 2 only.
// The file is dead code in the reference (never include()d, src/TorJ.jl:19-29) and has no tests or golden vectors:
// PARITY UNPINNED. The special functions are pinned independently against scipy in tests/test_oracle_warm.py
// (expi, gammaln, spherical_in, wofz).
#pragma once
#include <cmath>
#include <complex>
#include <vector>

namespace torj_oracle {
namespace warm {

typedef std::complex<double> cplx;

// reference src/constants.jl:1-4
constexpr int ntv = 501;
constexpr double tmax = 5.0;
constexpr double dt = 2.0 * tmax / (ntv - 1);
constexpr int i_max = 5;
constexpr double PI_SQRT = 1.7724538509055160272981674833411;  // constants.π_sqrt, src/constants.jl:25
// src/constants.jl:13-26
constexpr double C_LIGHT = 2.99792458e8, E_CHARGE = 1.602176634e-19, M_ELECTRON = 9.1093837015e-31;

struct ExtV {  // reference src/general_absorption.jl:8-13
    double ttv[ntv], extdtv[ntv];
    ExtV() {
        for (int i = 1; i <= ntv; ++i) {
            ttv[i - 1] = -tmax + (i - 1) * dt;
            extdtv[i - 1] = std::exp(-(ttv[i - 1] * ttv[i - 1])) * dt;
        }
    }
};
inline const ExtV& extv() {
    static const ExtV e;
    return e;
}

// reference src/general_absorption.jl:29-232
inline double expei(double x) {
    const double zero = 0.0, p037 = 0.037, half = 0.5, one = 1.0, two = 2.0, three = 3.0, four = 4.0, six = 6.0,
                 twelve = 12.0, two4 = 24.0, x01 = 381.5, x11 = 1024.0, x02 = -5.1182968633365538008e-5,
                 x0 = 0.37250741078136663466;
    const double xinf = 1.79e+308;
    static const double a[7] = {1.1669552669734461083368e2, 2.1500672908092918123209e3, 1.5924175980637303639884e4,
                                8.9904972007457256553251e4, 1.5026059476436982420737e5, -1.4815102102575750838086e5,
                                5.0196785185439843791020e0};
    static const double b[6] = {4.0205465640027706061433e1, 7.5043163907103936624165e2, 8.1258035174768735759855e3,
                                5.2440529172056355429883e4, 1.8434070063353677359298e5, 2.5666493484897117319268e5};
    static const double c[9] = {3.828573121022477169108e-1, 1.107326627786831743809e+1, 7.246689782858597021199e+1,
                                1.700632978311516129328e+2, 1.698106763764238382705e+2, 7.633628843705946890896e+1,
                                1.487967702840464066613e+1, 9.999989642347613068437e-1, 1.737331760720576030932e-8};
    static const double d[9] = {8.258160008564488034698e-2, 4.344836335509282083360e+0, 4.662179610356861756812e+1,
                                1.775728186717289799677e+2, 2.953136335677908517423e+2, 2.342573504717625153053e+2,
                                9.021658450529372642314e+1, 1.587964570758947927903e+1, 1.000000000000000000000e+0};
    static const double e[10] = {1.3276881505637444622987e+2, 3.5846198743996904308695e+4, 1.7283375773777593926828e+5,
                                 2.6181454937205639647381e+5, 1.7503273087497081314708e+5, 5.9346841538837119172356e+4,
                                 1.0816852399095915622498e+4, 1.0611777263550331766871e+3, 5.2199632588522572481039e+1,
                                 9.9999999999999999087819e-1};
    static const double f[10] = {3.9147856245556345627078e+4, 2.5989762083608489777411e+5, 5.5903756210022864003380e+5,
                                 5.4616842050691155735758e+5, 2.7858134710520842139357e+5, 7.9231787945279043698718e+4,
                                 1.2842808586627297365998e+4, 1.1635769915320848035459e+3, 5.4199632588522559414924e+1,
                                 1.0e0};
    static const double plg[4] = {-2.4562334077563243311e+01, 2.3642701335621505212e+02, -5.4989956895857911039e+02,
                                  3.5687548468071500413e+02};
    static const double qlg[4] = {-3.5553900764052419184e+01, 1.9400230218539473193e+02, -3.3442903192607538956e+02,
                                  1.7843774234035750207e+02};
    static const double p[10] = {-1.2963702602474830028590e+01, -1.2831220659262000678155e+03, -1.4287072500197005777376e+04,
                                 -1.4299841572091610380064e+06, -3.1398660864247265862050e+05, -3.5377809694431133484800e+08,
                                 3.1984354235237738511048e+08,  -2.5301823984599019348858e+10, 1.2177698136199594677580e+10,
                                 -2.0829040666802497120940e+11};
    static const double q[10] = {7.6886718750000000000000e+01, -5.5648470543369082846819e+03, 1.9418469440759880361415e+05,
                                 -4.2648434812177161405483e+06, 6.4698830956576428587653e+07, -7.0108568774215954065376e+08,
                                 5.4229617984472955011862e+09, -2.8986272696554495342658e+10, 9.8900934262481749439886e+10,
                                 -8.9673749185755048616855e+10};
    static const double r[10] = {-2.645677793077147237806e+00, -2.378372882815725244124e+00, -2.421106956980653511550e+01,
                                 1.052976392459015155422e+01,  1.945603779539281810439e+01,  -3.015761863840593359165e+01,
                                 1.120011024227297451523e+01,  -3.988850730390541057912e+00, 9.565134591978630774217e+00,
                                 9.981193787537396413219e-1};
    static const double s[9] = {1.598517957704779356479e-4, 4.644185932583286942650e+00, 3.697412299772985940785e+02,
                                -8.791401054875438925029e+00, 7.608194509086645763123e+02, 2.852397548119248700147e+01,
                                4.731097187816050252967e+02, -2.369210235636181001661e+02, 1.249884822712447891440e+00};
    static const double p1[10] = {-1.647721172463463140042e+00, -1.860092121726437582253e+01, -1.000641913989284829961e+01,
                                  -2.105740799548040450394e+01, -9.134835699998742552432e-1,  -3.323612579343962284333e+01,
                                  2.495487730402059440626e+01,  2.652575818452799819855e+01,  -1.845086232391278674524e+00,
                                  9.999933106160568739091e-1};
    static const double q1[9] = {9.792403599217290296840e+01, 6.403800405352415551324e+01, 5.994932325667407355255e+01,
                                 2.538819315630708031713e+02, 4.429413178337928401161e+01, 1.192832423968601006985e+03,
                                 1.991004470817742470726e+02, -1.093556195391091143924e+01, 1.001533852045342697818e+00};
    static const double p2[10] = {1.75338801265465972390e+02,  -2.23127670777632409550e+02, -1.81949664929868906455e+01,
                                  -2.79798528624305389340e+01, -7.63147701620253630855e+00, -1.52856623636929636839e+01,
                                  -7.06810977895029358836e+00, -5.00006640413131002475e+00, -3.00000000320981265753e+00,
                                  1.00000000000000485503e+00};
    static const double q2[9] = {3.97845977167414720840e+04, 3.97277109100414518365e+00, 1.37790390235747998793e+02,
                                 1.17179220502086455287e+02, 7.04831847180424675988e+01, -1.20187763547154743238e+01,
                                 -7.99243595776339741065e+00, -2.99999894040324959612e+00, 1.99999999999048104167e+00};
    if (x == zero) return -xinf;
    if (x < zero) {
        double y = std::fabs(x);
        if (y <= one) {
            double sump = a[6] * y + a[0], sumq = y + b[0];
            for (int i = 1; i < 6; ++i) { sump = sump * y + a[i]; sumq = sumq * y + b[i]; }
            return (std::log(y) - sump / sumq) * std::exp(y);
        } else if (y <= four) {
            double w = one / y, sump = c[0], sumq = d[0];
            for (int i = 1; i < 9; ++i) { sump = sump * w + c[i]; sumq = sumq * w + d[i]; }
            return -sump / sumq;
        } else {
            double w = one / y, sump = e[0], sumq = f[0];
            for (int i = 1; i < 10; ++i) { sump = sump * w + e[i]; sumq = sumq * w + f[i]; }
            return w * (w * sump / sumq - one);
        }
    } else if (x < six) {
        double t = (x + x) / three - two;
        double px[10], qx[10];
        px[0] = zero; qx[0] = zero; px[1] = p[0]; qx[1] = q[0];
        for (int i = 1; i < 9; ++i) {  // Julia i = 2:9 -> px[i+1] = t px[i] - px[i-1] + p[i]
            px[i + 1] = t * px[i] - px[i - 1] + p[i];
            qx[i + 1] = t * qx[i] - qx[i - 1] + q[i];
        }
        double sump = half * t * px[9] - px[8] + p[9];
        double sumq = half * t * qx[9] - qx[8] + q[9];
        double frac = sump / sumq;
        double xmx0 = (x - x01 / x11) - x02;
        if (std::fabs(xmx0) >= p037) return std::exp(-x) * (std::log(x / x0) + xmx0 * frac);
        double y = xmx0 / (x + x0), ysq = y * y;
        sump = plg[0]; sumq = ysq + qlg[0];
        for (int i = 1; i < 4; ++i) { sump = sump * ysq + plg[i]; sumq = sumq * ysq + qlg[i]; }
        return std::exp(-x) * (sump / (sumq * (x + x0)) + frac) * xmx0;
    } else if (x < twelve) {
        double frac = zero;
        for (int i = 0; i < 9; ++i) frac = s[i] / (r[i] + x + frac);
        return (r[9] + frac) / x;
    } else if (x <= two4) {
        double frac = zero;
        for (int i = 0; i < 9; ++i) frac = q1[i] / (p1[i] + x + frac);
        return (p1[9] + frac) / x;
    } else {
        double y = one / x, frac = zero;
        for (int i = 0; i < 9; ++i) frac = q2[i] / (p2[i] + x + frac);
        frac = p2[9] + frac;
        return y + y * y * frac;
    }
}

// reference src/general_absorption.jl:240-257
inline double fact(int k) {
    static const double precomp[16] = {1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 720.0, 5040.0, 40320.0, 362880.0, 3628800.0,
                                       39916800.0, 479001600.0, 6227020800.0, 87178291200.0, 1307674368000.0};
    if (k < 0) return 0.0;
    if (k < 16) return precomp[k];
    double result = precomp[15];
    for (int i = 16; i <= k; ++i) result = result * i;
    return result;
}

// reference src/general_absorption.jl:265-283
inline double gammln(double x) {
    const double stp = 2.5066282746310005;
    static const double cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155,
                                  0.1208650973866179e-2, -0.5395239384953e-5};
    double y = x, tmp = x + 5.5;
    tmp = (x + 0.5) * std::log(tmp) - tmp;
    double ser = 1.000000000190015;
    for (int j = 0; j < 6; ++j) { y = y + 1.0; ser = ser + cof[j] / y; }
    return tmp + std::log(stp * ser / x);
}

// reference src/general_absorption.jl:291-320 (assertion :314-316 dropped, see header). Entry mm-1 holds order m = n+mm-1.
inline std::vector<double> ssbi(double zz, int n, int l) {
    const double eps = 1.0e-10;
    const double z2q = 0.25 * (zz * zz);
    std::vector<double> out(l + 2, 0.0);
    for (int m = n; m <= l + 2; ++m) {
        double c0 = 1.0 / std::exp(gammln(m + 1.5));
        double sbi = c0;
        for (int k = 1; k <= 50; ++k) {
            double c1 = c0 * z2q / ((m + k) + 0.5) / k;
            sbi = sbi + c1;
            if (c1 / sbi < eps) break;
            c0 = c1;
        }
        int mm = m - n + 1;
        if (mm - 1 < (int)out.size()) out[mm - 1] = sbi;  // Julia would throw BoundsError for n < 1; never reached (n >= 1)
    }
    return out;
}

// reference src/general_absorption.jl:345-465. Julia's round() is round-half-to-even = nearbyint in the default mode.
inline cplx zetac(double xi, double yi) {
    const double factor = 1.12837916709551257388;
    const double rpi = 2.0 / factor;
    double xabs = std::fabs(xi), yabs = std::fabs(yi);
    double x = xabs / 6.3, y = yabs / 4.4;
    double qrho = x * x + y * y;
    double xabsq = xabs * xabs;
    double xquad = xabsq - yabs * yabs;
    double yquad = 2 * xabs * yabs;
    double u = 0, v = 0, u2 = 0, v2 = 0;
    const bool series = qrho < 0.085264;
    if (series) {
        qrho = (1 - 0.85 * y) * std::sqrt(qrho);
        int n = (int)std::nearbyint(6 + 72 * qrho);
        int j = 2 * n + 1;
        double xsum = 1.0 / j, ysum = 0.0;
        for (int i = n; i >= 1; --i) {
            j = j - 2;
            double xaux = (xsum * xquad - ysum * yquad) / i;
            ysum = (xsum * yquad + ysum * xquad) / i;
            xsum = xaux + 1.0 / j;
        }
        double u1 = -factor * (xsum * yabs + ysum * xabs) + 1.0;
        double v1 = factor * (xsum * xabs - ysum * yabs);
        double daux = std::exp(-xquad);
        u2 = daux * std::cos(yquad);
        v2 = -daux * std::sin(yquad);
        u = u1 * u2 - v1 * v2;
        v = u1 * v2 + v1 * u2;
    } else {
        double h = 0.0, h2 = 0.0;
        int kapn = 0, nu;
        if (qrho > 1.0) {
            qrho = std::sqrt(qrho);
            nu = (int)(3 + std::floor(1442 / (26 * qrho + 77)));  // Int(3 + (1442 ÷ (26 qrho + 77))): ÷ is floored division
        } else {
            qrho = (1 - y) * std::sqrt(1 - qrho);
            h = 1.88 * qrho;
            h2 = 2 * h;
            kapn = (int)std::nearbyint(7 + 34 * qrho);
            nu = (int)std::nearbyint(16 + 26 * qrho);
        }
        double qlambda = h > 0.0 ? std::pow(h2, kapn) : 0.0;
        double rx = 0, ry = 0, sx = 0, sy = 0;
        for (int n = nu; n >= 0; --n) {
            int np1 = n + 1;
            double tx = yabs + h + np1 * rx;
            double ty = xabs - np1 * ry;
            double c = 0.5 / (tx * tx + ty * ty);
            rx = c * tx;
            ry = c * ty;
            if (h > 0.0 && n <= kapn) {
                tx = qlambda + sx;
                sx = rx * tx - ry * sy;
                sy = ry * tx + rx * sy;
                qlambda = qlambda / h2;
            }
        }
        if (h == 0.0) { u = factor * rx; v = factor * ry; }
        else { u = factor * sx; v = factor * sy; }
        if (yabs == 0.0) u = std::exp(-(xabs * xabs));
    }
    if (yi < 0.0) {
        // NOTE: the reference tests `qrho < 0.085264` again here, but qrho has been overwritten above; restated as written
        if (qrho < 0.085264) {
            u2 = 2 * u2;
            v2 = 2 * v2;
        } else {
            xquad = -xquad;
            double w1 = 2.0 * std::exp(xquad);
            u2 = w1 * std::cos(yquad);
            v2 = -w1 * std::sin(yquad);
        }
        u = u2 - u;
        v = v2 - v;
        if (xi > 0.0) v = -v;
    } else {
        if (xi < 0.0) v = -v;
    }
    return cplx(-v * rpi, u * rpi);
}

// 0-based (is, ir) matrices of (lrm+1) x 3
struct Cef {
    int lrm;
    std::vector<cplx> p, m;
    explicit Cef(int l) : lrm(l), p((l + 1) * 3, cplx(0, 0)), m((l + 1) * 3, cplx(0, 0)) {}
    cplx& P(int is, int ir) { return p[is * 3 + ir]; }
    cplx& M(int is, int ir) { return m[is * 3 + ir]; }
};

// reference src/general_absorption.jl:473-561
inline Cef fsup(double yg, double anpl, double amu, int lrm) {
    const double soglia = 0.7;
    Cef cef(lrm);
    const double anpl2hm1 = anpl * anpl / 2.0 - 1.0;
    const double psi = std::sqrt(0.5 * amu) * anpl;
    const double apsi = std::fabs(psi);
    const cplx I(0.0, 1.0);
    for (int is = -lrm; is <= lrm; ++is) {
        double alpha = anpl2hm1 + is * yg;
        double phi2 = amu * alpha;
        double phim = std::sqrt(std::fabs(phi2));
        double xp, yp, xm, ym, x0, y0;
        if (alpha >= 0) { xp = psi - phim; yp = 0.0; xm = -psi - phim; ym = 0.0; x0 = -phim; y0 = 0.0; }
        else { xp = psi; yp = phim; xm = -psi; ym = phim; x0 = 0.0; y0 = phim; }
        cplx czp = zetac(xp, yp), czm = zetac(xm, ym);
        cplx cf12;
        if (alpha > 0) cf12 = -(czp + czm) / (2.0 * phim);
        else if (alpha < 0) cf12 = -I * (czp + czm) / (2.0 * phim);
        else cf12 = cplx(0.0, 0.0);
        cplx cf32;
        if (apsi > soglia) {
            cf32 = -(czp - czm) / (2.0 * psi);
        } else {
            cplx cphi = phim;
            if (alpha < 0) cphi = -I * phim;
            cplx cz0 = zetac(x0, y0);
            cf32 = 2.0 * (1.0 - cphi * cz0);
        }
        cplx cf0 = cf12, cf1 = cf32;
        if (is == 0) { cef.P(0, 0) = cf32; cef.M(0, 0) = cf32; }
        int isa = is < 0 ? -is : is;
        for (int l = 1; l <= isa + 2; ++l) {
            cplx cf2;
            if (apsi > soglia) cf2 = (1.0 + phi2 * cf0 - (l - 0.5) * cf1) / (psi * psi);
            else cf2 = (1.0 + phi2 * cf1) / (l + 0.5);
            int ir = l - isa;
            if (ir >= 0) {
                cef.P(isa, ir) += cf2;
                if (is > 0) cef.M(isa, ir) += cf2;
                else cef.M(isa, ir) -= cf2;
            }
            cf0 = cf1;
            cf1 = cf2;
        }
    }
    return cef;
}

struct Eps {  // epsl[a][b][l-1], 1 <= l <= lrm
    int lrm;
    std::vector<cplx> e;
    explicit Eps(int l) : lrm(l), e(9 * l, cplx(0, 0)) {}
    cplx& at(int a, int b, int l) { return e[((l - 1) * 3 + (a - 1)) * 3 + (b - 1)]; }
};

inline double ipow_m1(int k) { return (k % 2 == 0) ? 1.0 : -1.0; }  // Float64((-1)^k)

// reference src/general_absorption.jl:573-638
inline void dieltens_maxw_wr(double xg, double yg, double anpl, double amu, int lrm, cplx& e330, Eps& epsl) {
    const double anpl2 = anpl * anpl;
    Cef cef = fsup(yg, anpl, amu, lrm);
    const cplx I(0.0, 1.0);
    for (int l = 1; l <= lrm; ++l) {
        int lm = l - 1;
        double fcl = std::pow(0.5, l) * std::pow((1.0 / yg) * (1.0 / yg) / amu, lm) * fact(2 * l) / fact(l);
        cplx ca11 = 0, ca12 = 0, ca13 = 0, ca22 = 0, ca23 = 0, ca33 = 0;
        for (int is = 0; is <= l; ++is) {
            int k = l - is;
            double asl = ipow_m1(k) / (fact(is + l) * fact(l - is));
            double bsl = asl * (is * is + (double)(2 * k * lm * (l + is)) / (2 * l - 1));
            cplx cq0p = amu * cef.P(is, 0), cq0m = amu * cef.M(is, 0);
            cplx cq1p = amu * anpl * (cef.P(is, 0) - cef.P(is, 1));
            cplx cq1m = amu * anpl * (cef.M(is, 0) - cef.M(is, 1));
            cplx cq2p = cef.P(is, 1) + amu * anpl2 * (cef.P(is, 2) + cef.P(is, 0) - 2.0 * cef.P(is, 1));
            ca11 += (double)(is * is) * asl * cq0p;
            ca12 += (double)(is * l) * asl * cq0m;
            ca22 += bsl * cq0p;
            ca13 += (double)is * asl * cq1m / yg;
            ca23 += (double)l * asl * cq1p / yg;
            ca33 += asl * cq2p / (yg * yg);
        }
        epsl.at(1, 1, l) = -xg * ca11 * fcl;
        epsl.at(1, 2, l) = I * xg * ca12 * fcl;
        epsl.at(2, 2, l) = -xg * ca22 * fcl;
        epsl.at(1, 3, l) = -xg * ca13 * fcl;
        epsl.at(2, 3, l) = -I * xg * ca23 * fcl;
        epsl.at(3, 3, l) = -xg * ca33 * fcl;
    }
    cplx cq2p = cef.P(0, 1) + amu * anpl2 * (cef.P(0, 2) + cef.P(0, 0) - 2.0 * cef.P(0, 1));
    e330 = 1.0 - xg * amu * cq2p;
    epsl.at(1, 1, 1) = 1.0 + epsl.at(1, 1, 1);
    epsl.at(2, 2, 1) = 1.0 + epsl.at(2, 2, 1);
    for (int l = 1; l <= lrm; ++l) {
        epsl.at(2, 1, l) = -epsl.at(1, 2, l);
        epsl.at(3, 1, l) = epsl.at(1, 3, l);
        epsl.at(3, 2, l) = -epsl.at(2, 3, l);
    }
}

// rr(n, k, m): n in -lrm..lrm, k in 0..2, m in 0..lrm
struct RR {
    int lrm;
    std::vector<double> v;
    explicit RR(int l) : lrm(l), v((2 * l + 1) * 3 * (l + 1), 0.0) {}
    double& at(int n, int k, int m) { return v[((n + lrm) * 3 + k) * (lrm + 1) + m]; }
};
// ri(n, k, m): n in 1..lrm, k in 0..2, m in 1..lrm
struct RI {
    int lrm;
    std::vector<double> v;
    explicit RI(int l) : lrm(l), v(l * 3 * l, 0.0) {}
    double& at(int n, int k, int m) { return v[((n - 1) * 3 + k) * lrm + (m - 1)]; }
};

// reference src/general_absorption.jl:646-714 (iwarm > 2: the quadrature for every n in -llm..llm, returned as is)
inline RR hermitian(double yg, double anpl, double amu, int lrm, int iwarm) {
    RR rr(lrm);
    const ExtV& E = extv();
    const double cmxw = 1.0 + 15.0 / (8.0 * amu) + 105.0 / (128.0 * (amu * amu));
    const double cr = -amu * amu / (PI_SQRT * cmxw);
    const int llm = std::min(3, lrm);
    const double bth2 = 2.0 / amu, bth = std::sqrt(bth2);
    const double amu2 = amu * amu, amu4 = amu2 * amu2, amu6 = amu4 * amu2, amu8 = amu4 * amu4;
    const int n1 = iwarm > 2 ? -llm : 1;
    for (int i = 0; i < ntv; ++i) {
        double t = E.ttv[i];
        double rxt = std::sqrt(1.0 + t * t / (2.0 * amu));
        double x = t * rxt;
        double upl2 = bth2 * (x * x);
        double upl = bth * x;
        double gx = 1.0 + t * t / amu;
        double exdx = cr * E.extdtv[i] * gx / rxt;
        for (int n = n1; n <= llm; ++n) {
            int nn = n < 0 ? -n : n;
            double gr = anpl * upl + n * yg;
            double zm = -amu * (gx - gr);
            double s = amu * (gx + gr);
            double zm2 = zm * zm, zm3 = zm2 * zm;
            double fe0m = expei(zm);
            for (int m = nn; m <= llm; ++m) {
                if (m == 0) {
                    rr.at(0, 2, 0) += -exdx * fe0m * upl2;
                } else {
                    double ffe;
                    if (m == 1) ffe = (1.0 + s * (1.0 - zm * fe0m)) / amu2;
                    else if (m == 2) ffe = (6.0 - 2.0 * zm + 4.0 * s + s * s * (1.0 + zm - zm2 * fe0m)) / amu4;
                    else if (m == 3)
                        ffe = (18.0 * s * (s + 4.0 - zm) + 6.0 * (20.0 - 8.0 * zm + zm2) + s * s * s * (2.0 + zm + zm2 - zm3 * fe0m)) / amu6;
                    else
                        ffe = (96.0 * s * (30.0 + s * s - 10.0 * zm + zm * zm) +
                               24.0 * (210.0 - 6.0 * (s * s) * (-5.0 + zm) - 90.0 * zm + 15.0 * (zm * zm) - zm * zm * zm) +
                               s * s * s * s * (6.0 + 2.0 * zm + zm * zm + zm * zm * zm - fe0m * (zm2 * zm2))) / amu8;
                    rr.at(n, 0, m) += exdx * ffe;
                    rr.at(n, 1, m) += exdx * ffe * upl;
                    rr.at(n, 2, m) += exdx * ffe * upl2;
                }
            }
        }
    }
    return rr;  // iwarm > 2 (see header for iwarm <= 2)
}

// reference src/general_absorption.jl:951-1043
inline RI antihermitian(double yg, double anpl, double amu, int lrm) {
    RI ri(lrm);
    const double dnl = 1.0 - anpl * anpl;
    const double cmu = anpl * amu;
    const double cmxw = 1.0 + 15.0 / (8.0 * amu) + 105.0 / (128.0 * (amu * amu));
    const double ci = std::sqrt(2.0 * M_PI * amu) * (amu * amu) / cmxw;
    for (int n = 1; n <= lrm; ++n) {
        double ygn = n * yg;
        double rdu2 = ygn * ygn - dnl;
        if (!(rdu2 > 0.0)) continue;
        double rdu = std::sqrt(rdu2);
        double du = rdu / dnl;
        double ub = anpl * ygn / dnl;
        double aa = amu * anpl * du;
        if (std::fabs(aa) > 5.0) {
            double up = ub + du, um = ub - du;
            double gp = anpl * up + ygn, gm = anpl * um + ygn;
            double xp = up + 1.0 / cmu, xm = um + 1.0 / cmu;
            double eem = std::exp(-amu * (gm - 1.0)), eep = std::exp(-amu * (gp - 1.0));
            double fi0p0 = -1.0 / cmu, fi1p0 = -xp / cmu, fi2p0 = -(1.0 / (cmu * cmu) + xp * xp) / cmu;
            double fi0m0 = -1.0 / cmu, fi1m0 = -xm / cmu, fi2m0 = -(1.0 / (cmu * cmu) + xm * xm) / cmu;
            for (int m = 1; m <= lrm; ++m) {
                double fi0p1 = -2.0 * m * (fi1p0 - ub * fi0p0) / cmu;
                double fi0m1 = -2.0 * m * (fi1m0 - ub * fi0m0) / cmu;
                double fi1p1 = -((1.0 + 2 * m) * fi2p0 - 2.0 * (m + 1) * ub * fi1p0 + up * um * fi0p0) / cmu;
                double fi1m1 = -((1.0 + 2 * m) * fi2m0 - 2.0 * (m + 1) * ub * fi1m0 + up * um * fi0m0) / cmu;
                double fi2p1 = (2.0 * (1 + m) * fi1p1 - 2.0 * m * (ub * fi2p0 - up * um * fi1p0)) / cmu;
                double fi2m1 = (2.0 * (1 + m) * fi1m1 - 2.0 * m * (ub * fi2m0 - up * um * fi1m0)) / cmu;
                if (m >= n) {
                    double dm = std::pow(dnl, m);
                    ri.at(n, 0, m) = 0.5 * ci * dm * (fi0p1 * eep - fi0m1 * eem);
                    ri.at(n, 1, m) = 0.5 * ci * dm * (fi1p1 * eep - fi1m1 * eem);
                    ri.at(n, 2, m) = 0.5 * ci * dm * (fi2p1 * eep - fi2m1 * eem);
                }
                fi0p0 = fi0p1; fi1p0 = fi1p1; fi2p0 = fi2p1;
                fi0m0 = fi0m1; fi1m0 = fi1m1; fi2m0 = fi2m1;
            }
        } else {
            double ee = std::exp(-amu * (ygn - 1.0 + anpl * ub));
            std::vector<double> fsbi = ssbi(aa, n, lrm);
            for (int m = n; m <= lrm; ++m) {
                double cm = PI_SQRT * fact(m) * std::pow(du, 2 * m + 1);
                double cim = 0.5 * ci * std::pow(dnl, m);
                int mm = m - n + 1;
                double fi0m = cm * fsbi[mm - 1];
                double fi1m = -0.5 * aa * cm * fsbi[mm];
                double fi2m = 0.5 * cm * (fsbi[mm] + 0.5 * aa * aa * fsbi[mm + 1]);
                ri.at(n, 0, m) = cim * ee * fi0m;
                ri.at(n, 1, m) = cim * ee * (du * fi1m + ub * fi0m);
                ri.at(n, 2, m) = cim * ee * (du * du * fi2m + 2.0 * du * ub * fi1m + ub * ub * fi0m);
            }
        }
    }
    return ri;
}

// reference src/general_absorption.jl:1056-1134
inline void dieltens_maxw_fr(double xg, double yg, double anpl, double amu, int lrm, int iwarm, cplx& e330, Eps& epsl) {
    RR rr = hermitian(yg, anpl, amu, lrm, iwarm);
    RI ri = antihermitian(yg, anpl, amu, lrm);
    const cplx I(0.0, 1.0);
    for (int l = 1; l <= lrm; ++l) {
        int lm = l - 1;
        double fal = -std::pow(0.25, l) * fact(2 * l) / (fact(l) * fact(l) * std::pow(yg, 2 * lm));
        cplx ca11 = 0, ca12 = 0, ca13 = 0, ca22 = 0, ca23 = 0, ca33 = 0;
        for (int is = 0; is <= l; ++is) {
            int k = l - is;
            double asl = ipow_m1(k) / (fact(is + l) * fact(l - is));
            double bsl = asl * (is * is + (double)(2 * k * lm * (l + is)) / (2 * l - 1));
            cplx cq0p, cq0m, cq1p, cq1m, cq2p;
            if (is > 0) {
                cq0p = cplx(rr.at(is, 0, l) + rr.at(-is, 0, l), ri.at(is, 0, l));
                cq0m = cplx(rr.at(is, 0, l) - rr.at(-is, 0, l), ri.at(is, 0, l));
                cq1p = cplx(rr.at(is, 1, l) + rr.at(-is, 1, l), ri.at(is, 1, l));
                cq1m = cplx(rr.at(is, 1, l) - rr.at(-is, 1, l), ri.at(is, 1, l));
                cq2p = cplx(rr.at(is, 2, l) + rr.at(-is, 2, l), ri.at(is, 2, l));
            } else {
                cq0p = cplx(rr.at(0, 0, l), 0.0);
                cq0m = cplx(rr.at(0, 0, l), 0.0);
                cq1p = cplx(rr.at(0, 1, l), 0.0);
                cq1m = cplx(rr.at(0, 1, l), 0.0);
                cq2p = cplx(rr.at(0, 2, l), 0.0);
            }
            ca11 += (double)(is * is) * asl * cq0p;
            ca12 += (double)(is * l) * asl * cq0m;
            ca22 += bsl * cq0p;
            ca13 += (double)is * asl * cq1m / yg;
            ca23 += (double)l * asl * cq1p / yg;
            ca33 += asl * cq2p / (yg * yg);
        }
        epsl.at(1, 1, l) = -xg * ca11 * fal;
        epsl.at(1, 2, l) = I * xg * ca12 * fal;
        epsl.at(2, 2, l) = -xg * ca22 * fal;
        epsl.at(1, 3, l) = -xg * ca13 * fal;
        epsl.at(2, 3, l) = -I * xg * ca23 * fal;
        epsl.at(3, 3, l) = -xg * ca33 * fal;
    }
    e330 = 1.0 + xg * cplx(rr.at(0, 2, 0), 0.0);
    epsl.at(1, 1, 1) = 1.0 + epsl.at(1, 1, 1);
    epsl.at(2, 2, 1) = 1.0 + epsl.at(2, 2, 1);
    for (int l = 1; l <= lrm; ++l) {
        epsl.at(2, 1, l) = -epsl.at(1, 2, l);
        epsl.at(3, 1, l) = epsl.at(1, 3, l);
        epsl.at(3, 2, l) = -epsl.at(2, 3, l);
    }
}

// Julia z^n for a small non-negative integer n (Base.power_by_squaring)
inline cplx cpow_int(cplx z, int n) {
    if (n == 0) return cplx(1.0, 0.0);
    if (n == 1) return z;
    if (n == 2) return z * z;
    if (n == 3) return z * z * z;
    cplx z2 = z * z;
    return z2 * z2;  // n == 4 (lrm <= i_max = 5)
}

struct WarmDisp {
    cplx anpr, ex, ey, ez;
    int ierr = 0;
    int iterations = 0;
};

// reference src/general_absorption.jl:1158-1267
inline WarmDisp warmdisp(double xg, double yg, double anpl, double amu, double anprc, int sox, int iwarm, int lrm) {
    const int imx = 100;
    WarmDisp out;
    double errnpr = 1.0;
    cplx anpr2a(anprc * anprc, 0.0);
    const double anpl2 = anpl * anpl;
    cplx e330;
    Eps epsl(lrm);
    if (iwarm == 1) dieltens_maxw_wr(xg, yg, anpl, amu, lrm, e330, epsl);
    else dieltens_maxw_fr(xg, yg, anpl, amu, lrm, iwarm, e330, epsl);
    cplx e11, e12, e13, e22, e23, anpr2(0, 0);
    for (int i = 1; i <= imx; ++i) {
        cplx sepsl[3][3];
        for (int j = 1; j <= 3; ++j)
            for (int k = 1; k <= 3; ++k) {
                cplx acc(0, 0);
                for (int il = 1; il <= lrm; ++il) acc += epsl.at(k, j, il) * cpow_int(anpr2a, il - 1);
                sepsl[k - 1][j - 1] = acc;
            }
        cplx anpra = std::sqrt(anpr2a);
        e11 = sepsl[0][0];
        e22 = sepsl[1][1];
        e12 = sepsl[0][1];
        cplx a33 = sepsl[2][2], a13 = sepsl[0][2], a23 = sepsl[1][2];
        cplx a31 = a13, a32 = -a23;
        e13 = anpra * a13;
        e23 = anpra * a23;
        if (i > 2 && errnpr < 1.0e-4) break;
        out.iterations = i;
        cplx cc4 = (e11 - anpl2) * (1.0 - a33) + (a13 + anpl) * (a31 + anpl);
        cplx cc2 = -e12 * e12 * (1.0 - a33) - a32 * e12 * (a13 + anpl) + a23 * e12 * (a31 + anpl) -
                   (a23 * a32 + e330 + (e22 - anpl2) * (1.0 - a33)) * (e11 - anpl2) - (a13 + anpl) * (a31 + anpl) * (e22 - anpl2);
        cplx cc0 = e330 * ((e11 - anpl2) * (e22 - anpl2) + e12 * e12);
        cplx rr = cc2 * cc2 - 4.0 * cc0 * cc4;
        double s;
        if (yg > 1.0) {
            s = (double)sox;
            if (rr.imag() <= 0.0) s = -s;
        } else {
            s = (double)(-sox);
            if (rr.real() <= 0.0 && rr.imag() >= 0.0) s = -s;
        }
        anpr2 = (-cc2 + s * std::sqrt(rr)) / (2.0 * cc4);
        errnpr = std::fabs(1.0 - std::abs(anpr2) / std::abs(anpr2a));
        anpr2a = anpr2;
    }
    if (anpr2.real() < 0.0 && anpr2.imag() < 0.0) {
        anpr2 = cplx(0.0, 0.0);
        out.ierr = 99;
    }
    cplx anpr = std::sqrt(anpr2);
    cplx ex(0, 0), ey(0, 0), ez(0, 0);
    if (std::fabs(anpl) > 1.0e-6) {
        cplx den = e12 * e23 - (e13 + anpr * anpl) * (e22 - anpr2 - anpl2);
        ey = -(e12 * (e13 + anpr * anpl) + (e11 - anpl2) * e23) / den;
        ez = (e12 * e12 + (e22 - anpr2 - anpl2) * (e11 - anpl2)) / den;
        double ez2 = std::norm(ez), ey2 = std::norm(ey);  // abs(z)^2
        double enx2 = 1.0 / (1.0 + ey2 + ez2);
        ex = cplx(std::sqrt(enx2), 0.0);
        ez = ez * ex;
        ey = ey * ex;
    } else {
        if (sox < 0) {
            ez = cplx(1.0, 0.0);
        } else {
            double r = std::abs(-e11 / e12);
            double ex2 = 1.0 / (1.0 + r * r);
            ex = cplx(std::sqrt(ex2), 0.0);
            ey = -ex * e11 / e12;
        }
    }
    out.anpr = anpr; out.ex = ex; out.ey = ey; out.ez = ez;
    return out;
}

// reference src/general_absorption.jl:1285-1326
inline int larmornumber(double yg, double npl, double mu) {
    const double expcr = 15.0;
    const double dnl = 1.0 - npl * npl;
    int imax = 1;
    int nharm = (int)std::floor(1.0 / yg);
    if (nharm * yg < 1.0) nharm = nharm + 1;
    while (true) {
        double ygn = nharm * yg;
        double rdu2 = ygn * ygn - dnl;
        double gg = (ygn - std::sqrt(npl * npl * rdu2)) / dnl;
        double argexp = mu * (gg - 1.0);
        if (argexp > expcr) break;
        nharm = nharm + 1;
        imax = imax + 1;
        if (imax > 100) {
            nharm = (int)std::floor(yg);
            break;
        }
    }
    return nharm;
}

struct AlphaOut {
    double N_warm = 0.0, alpha = 0.0;
    int lrm = 0, ierr = 0, iterations = 0;
    cplx N_perp;
};

// reference src/general_absorption.jl:1328-1337
inline AlphaOut alpha(double omega, double X, double Y, double N_r, double theta, double te, double v_g_perp, int imod) {
    AlphaOut o;
    double N_par = N_r * std::cos(theta);
    double Npr = std::sqrt(N_r * N_r - N_par * N_par);
    double mu = M_ELECTRON * (C_LIGHT * C_LIGHT) / (te * E_CHARGE);
    int nharm = larmornumber(Y, N_par, mu);
    int lrm = std::min(i_max, nharm);
    o.lrm = lrm;
    if (lrm < 1) {  // Julia would build 3x3x0 arrays and fail on epsl[1,1,1] (BoundsError); reported as ierr
        o.ierr = 98;
        o.alpha = std::nan("");
        return o;
    }
    WarmDisp w = warmdisp(X, Y, N_par, mu, Npr, imod, 3, lrm);
    o.N_perp = w.anpr;
    o.ierr = w.ierr;
    o.iterations = w.iterations;
    o.N_warm = w.anpr.real() / std::sin(theta);
    cplx n2 = w.anpr * w.anpr;
    o.alpha = 2.e0 * n2.imag() * omega / C_LIGHT * v_g_perp;
    return o;
}

}  // namespace warm
}  // namespace torj_oracle

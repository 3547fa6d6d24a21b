// Host side of libtorj_cuda.so: the C ABI declared in include/torj_cuda.h.
// Plain CUDA runtime; no torch types, no CPU fallback (every entry point needs a CUDA device).
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <memory>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/torj_cuda.h"
#include "torj_kernels.cuh"

using namespace torj;

static thread_local std::string g_err;

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            char buf_[512];                                                                               \
            snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            g_err = buf_;                                                                                 \
            return 1;                                                                                     \
        }                                                                                                 \
    } while (0)

#define FAIL(msg) do { g_err = (msg); return 2; } while (0)

struct torj_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool gl_set = false;
    int num_sms = 0;
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // bracket the last k_trace launch
    bool ev_valid = false;
    struct torj_bundle* ws = nullptr;  // workspace of the one-shot torj_trace call, kept between calls
};

static std::atomic<uint64_t> g_plasma_serial{0};  // torj_multi_* creates plasmas from one host thread per device

struct torj_plasma {
    torj_ctx* ctx = nullptr;
    uint64_t id = 0;  // unique per created plasma (addresses can be recycled)
    DevTables T{};
    double2* dA = nullptr;
    double2* dB = nullptr;
    DevTables* dT = nullptr;  // T mirrored in global memory
    // host copy of the V(psi_N) spline (reference src/plasma.jl:42-44)
    std::vector<double> vol_c;
    int n_vol = 0;
    double vol_x0 = 0, vol_h = 1;
};

struct torj_bundle {
    torj_ctx* ctx = nullptr;
    int64_t n = 0;
    BundleDev B{};
    int per_ray_fm = 0;
    // owned device buffers
    double *d_pos = nullptr, *d_dir = nullptr, *d_w = nullptr, *d_freq = nullptr, *d_u0 = nullptr, *d_s0 = nullptr,
           *d_psil = nullptr, *d_Pf = nullptr, *d_Pdep = nullptr;
    int *d_mode = nullptr, *d_status = nullptr, *d_npts = nullptr;
    // deposition
    int n_psi = 0;
    double *d_edges = nullptr, *d_bins = nullptr, *d_dV = nullptr, *d_profile = nullptr;
    unsigned long long *d_queue = nullptr, *d_counters = nullptr;
    // segment hand-off (allocated on first use)
    double* d_hand = nullptr;
    int *d_segdone = nullptr, *d_left = nullptr;
    int h_left = 0;  // source of the copy into d_left (outlives the call)
    // trajectory window
    int64_t traj_first = 0, traj_count = 0;
    int traj_max = 0;
    double *d_ts = nullptr, *d_txyz = nullptr, *d_tP = nullptr, *d_tdP = nullptr, *d_tprof = nullptr;
    int traj_prof_npsi = 0;
    // what the device copies of psi_edges / dV were computed from (skips the re-upload when unchanged)
    std::vector<double> h_edges;
    uint64_t edges_plasma = 0;
    // beams
    int n_beams = 1, prof_rows = 0;
    int* d_beam = nullptr;
};

// Error paths: temporaries are owned by dev_ptr, objects under construction by a guard that destroys them unless the
// function reaches its end (every CK / FAIL returns early).
struct DevFree {
    void operator()(void* p) const { cudaFree(p); }
};
template <class T>
using dev_ptr = std::unique_ptr<T, DevFree>;
template <class T>
static cudaError_t dev_alloc(dev_ptr<T>& p, size_t count) {
    T* raw = nullptr;
    cudaError_t e = cudaMalloc(&raw, count * sizeof(T));
    p.reset(raw);
    return e;
}
template <class T, void (*Destroy)(T*)>
struct Guard {
    T* p;
    explicit Guard(T* q) : p(q) {}
    ~Guard() { if (p) Destroy(p); }
    T* release() { T* q = p; p = nullptr; return q; }
    Guard(const Guard&) = delete;
    Guard& operator=(const Guard&) = delete;
};

static int set_device(const torj_ctx* c) {
    CK(cudaSetDevice(c->device));
    return 0;
}

extern "C" {

const char* torj_last_error(void) { return g_err.c_str(); }
int torj_abi_version(void) { return TORJ_ABI_VERSION; }

void torj_options_default(torj_options* o) {
    o->scheme = 0;
    o->n_segments = 100;
    o->dtmax = 1e-4;
    o->abstol = 1e-6;
    o->reltol = 1e-6;
    o->psi_stop = 1.0;
    o->p_stop = 1e-6;
    o->te_min = 20.0;
    o->max_harmonic = 3;
    o->schedule = 0;
    o->reserved_ = 0;
    o->max_steps_per_segment = 100000;
    o->alpha_floor = 1e-14;
}

static void fill_tableaux(Tableau t[2]) {
    memset(t, 0, 2 * sizeof(Tableau));
    double(*a)[7] = t[0].a;  // Tsit5
    a[1][0] = 0.161;
    a[2][0] = -0.008480655492356989; a[2][1] = 0.335480655492357;
    a[3][0] = 2.8971530571054935; a[3][1] = -6.359448489975075; a[3][2] = 4.3622954328695815;
    a[4][0] = 5.325864828439257; a[4][1] = -11.748883564062828; a[4][2] = 7.4955393428898365; a[4][3] = -0.09249506636175525;
    a[5][0] = 5.86145544294642; a[5][1] = -12.92096931784711; a[5][2] = 8.159367898576159; a[5][3] = -0.071584973281401;
    a[5][4] = -0.028269050394068383;
    a[6][0] = 0.09646076681806523; a[6][1] = 0.01; a[6][2] = 0.4798896504144996; a[6][3] = 1.379008574103742;
    a[6][4] = -3.290069515436081; a[6][5] = 2.324710524099774;
    const double bt[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                          0.5823571654525552, -0.45808210592918697, 0.015151515151515152};
    for (int i = 0; i < 7; ++i) t[0].bt[i] = bt[i];
    a = t[1].a;  // OwrenZen3
    a[1][0] = 12.0 / 23.0;
    a[2][0] = -68.0 / 375.0; a[2][1] = 368.0 / 375.0;
    a[3][0] = 31.0 / 144.0; a[3][1] = 529.0 / 1152.0; a[3][2] = 125.0 / 384.0;
    t[1].bt[0] = 25.0 / 144.0; t[1].bt[1] = -575.0 / 1152.0; t[1].bt[2] = 125.0 / 384.0;
}

int torj_ctx_create(int device, void* cuda_stream, torj_ctx** out) {
    if (!out) FAIL("torj_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_err = std::string("torj_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); libtorj_cuda has no CPU path";
        return 1;
    }
    if (device < 0 || device >= ndev) FAIL("torj_ctx_create: bad device index");
    CK(cudaSetDevice(device));
    torj_ctx* c = new torj_ctx();
    Guard<torj_ctx, torj_ctx_destroy> guard(c);
    c->device = device;
    if (cuda_stream) {
        c->stream = (cudaStream_t)cuda_stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    Tableau t[2];
    fill_tableaux(t);
    CK(cudaMemcpyToSymbol(c_tab, t, sizeof t));
    // Bessel series coefficients c[n][k] = (-1)^k n!/(k!(n+k)!)
    double bc[5][TORJ_BESS_K];
    for (int n = 0; n < 5; ++n) {
        bc[n][0] = 1.0;
        for (int k = 1; k < TORJ_BESS_K; ++k) bc[n][k] = -bc[n][k - 1] / ((double)k * (double)(n + k));
    }
    CK(cudaMemcpyToSymbol(c_bess, bc, sizeof bc));
    double bd[5][TORJ_BESS_K];
    for (int n = 0; n < 5; ++n)
        for (int k = 0; k < TORJ_BESS_K; ++k) bd[n][k] = (double)(n + 2 * k) * bc[n][k];
    CK(cudaMemcpyToSymbol(c_bessd, bd, sizeof bd));
    {   // exp_fast: range reduction constants and Taylor coefficients 1/13! ... 1/2!, 1/1!, 1/0!
        double ce[17] = {1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10};
        double fact = 1.0, inv[14];
        for (int n = 0; n <= 13; ++n) { if (n > 0) fact *= (double)n; inv[n] = 1.0 / fact; }
        for (int j = 0; j <= 13; ++j) ce[3 + j] = inv[13 - j];
        CK(cudaMemcpyToSymbol(c_exp, ce, sizeof ce));
    }
    *out = guard.release();
    return 0;
}

void torj_ctx_destroy(torj_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->ws) torj_bundle_destroy(c->ws);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
}

int torj_ctx_sync(torj_ctx* c) {
    if (set_device(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int64_t torj_ctx_launch_count(const torj_ctx* c) { return c->launches; }

int torj_ctx_last_trace_ms(torj_ctx* c, double* ms) {
    if (!c->ev_valid) FAIL("torj_ctx_last_trace_ms: no trace launched yet");
    if (set_device(c)) return 1;
    CK(cudaEventSynchronize(c->ev1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, c->ev0, c->ev1));
    *ms = (double)t;
    return 0;
}

int torj_abs_init(torj_ctx* c, int32_t n, const double* nodes, const double* weights) {
    if (n < 1 || n > TORJ_MAX_GL) FAIL("torj_abs_init: need 1 <= n <= 64");
    if (set_device(c)) return 1;
    GLNodes g;
    memset(&g, 0, sizeof g);
    g.n = n;
    for (int i = 0; i < n; ++i) {
        g.t[i] = nodes[i];
        g.w[i] = weights[i];
        g.sq[i] = std::sqrt(1.0 - nodes[i] * nodes[i]);
    }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpyToSymbol(c_gl, &g, sizeof g));
    c->gl_set = true;
    return 0;
}

// ---- B-spline prefilter (host): rows (1/6, 2/3, 1/6), natural ends c0-2c1+c2 = 0 -> c[1]=y[0], c[n]=y[n-1]
static void prefilter_line(const double* y, size_t ys, int n, double* c, size_t cs) {
    std::vector<double> sol(n), cp(n), dp(n);
    sol[0] = y[0];
    sol[n - 1] = y[(size_t)(n - 1) * ys];
    if (n > 2) {
        const double a = 1.0 / 6.0, b = 2.0 / 3.0;
        int m = n - 2;
        cp[0] = a / b;
        dp[0] = (y[ys] - a * sol[0]) / b;
        for (int k = 1; k < m; ++k) {
            double r = y[(size_t)(k + 1) * ys];
            if (k == m - 1) r -= a * sol[n - 1];
            double den = b - a * cp[k - 1];
            cp[k] = a / den;
            dp[k] = (r - a * dp[k - 1]) / den;
        }
        if (m == 1) dp[0] = (y[ys] - a * sol[0] - a * sol[n - 1]) / b;
        sol[m] = dp[m - 1];
        for (int k = m - 2; k >= 0; --k) sol[k + 1] = dp[k] - cp[k] * sol[k + 2];
    }
    for (int k = 0; k < n; ++k) c[(size_t)(k + 1) * cs] = sol[k];
    c[0] = 2.0 * sol[0] - sol[1];
    c[(size_t)(n + 1) * cs] = 2.0 * sol[n - 1] - sol[n - 2];
}

int torj_bspline_prefilter_1d(int32_t n, const double* data, double* coef) {
    if (n < 2) FAIL("torj_bspline_prefilter_1d: n < 2");
    prefilter_line(data, 1, n, coef, 1);
    return 0;
}

int torj_bspline_prefilter_2d(int32_t nR, int32_t nZ, const double* data, double* coef) {
    if (nR < 2 || nZ < 2) FAIL("torj_bspline_prefilter_2d: need at least 2 points per axis");
    size_t sr = (size_t)nR + 2;
    std::vector<double> tmp(sr * nZ);
    for (int j = 0; j < nZ; ++j) prefilter_line(data + (size_t)j * nR, 1, nR, tmp.data() + (size_t)j * sr, 1);
    for (size_t i = 0; i < sr; ++i) prefilter_line(tmp.data() + i, sr, nZ, coef + i, sr);
    return 0;
}

int torj_plasma_create(torj_ctx* c, const torj_grid* g, const double* coef_psi, const double* coef_lnne,
                       const double* coef_lnTe, const double* coef_BR, const double* coef_BZ, const double* coef_Bphi,
                       const double* vol_coef, int32_t n_vol, double vol_psi0, double vol_dpsi, double psi_prof_max,
                       torj_plasma** out) {
    if (!c || !g || !out) FAIL("torj_plasma_create: NULL argument");
    if (g->nR < 2 || g->nZ < 2 || n_vol < 2) FAIL("torj_plasma_create: grid too small");
    if (set_device(c)) return 1;
    torj_plasma* p = new torj_plasma();
    Guard<torj_plasma, torj_plasma_destroy> guard(p);
    p->ctx = c;
    p->id = ++g_plasma_serial;
    size_t nodes = (size_t)(g->nR + 2) * (g->nZ + 2);
    std::vector<double2> hA(2 * nodes), hB(nodes);
    for (size_t k = 0; k < nodes; ++k) {
        hA[2 * k] = make_double2(coef_BR[k], coef_BZ[k]);
        hA[2 * k + 1] = make_double2(coef_Bphi[k], coef_lnne[k]);
        hB[k] = make_double2(coef_lnTe[k], coef_psi[k]);
    }
    CK(cudaMalloc(&p->dA, 2 * nodes * sizeof(double2)));
    CK(cudaMalloc(&p->dB, nodes * sizeof(double2)));
    CK(cudaMemcpyAsync(p->dA, hA.data(), 2 * nodes * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(p->dB, hB.data(), nodes * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    DevTables& T = p->T;
    T.A = p->dA; T.B = p->dB;
    T.nR = g->nR; T.nZ = g->nZ; T.row = g->nR + 2;
    T.r0 = g->R_first; T.z0 = g->Z_first;
    double hr = (g->R_last - g->R_first) / (double)(g->nR - 1), hz = (g->Z_last - g->Z_first) / (double)(g->nZ - 1);
    T.inv_hr = 1.0 / hr; T.inv_hz = 1.0 / hz;
    T.rlast = g->R_first + hr * (g->nR - 1); T.zlast = g->Z_first + hz * (g->nZ - 1);
    T.psi_prof_max = psi_prof_max;
    CK(cudaMalloc(&p->dT, sizeof(DevTables)));
    CK(cudaMemcpy(p->dT, &T, sizeof(DevTables), cudaMemcpyHostToDevice));
    p->vol_c.assign(vol_coef, vol_coef + n_vol + 2);
    p->n_vol = n_vol; p->vol_x0 = vol_psi0; p->vol_h = vol_dpsi;
    *out = guard.release();
    return 0;
}

// natural cubic spline through (x, y), x strictly increasing, resampled on the uniform range x[0]..x[n-1] with n
// points: IMAS.interp1d(x, y, :cubic).(range(x[1], x[end], length(x))) of reference src/plasma.jl:17-18,42-43
static void resample_uniform(const double* x, const double* y, int n, std::vector<double>& xr, std::vector<double>& yr) {
    std::vector<double> m2(n, 0.0);  // second derivatives, natural ends
    if (n > 2) {
        std::vector<double> cp(n), dp(n);
        for (int i = 1; i < n - 1; ++i) {
            double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
            double diag = 2.0 * (h0 + h1), rhs = 6.0 * ((y[i + 1] - y[i]) / h1 - (y[i] - y[i - 1]) / h0);
            if (i == 1) { cp[i] = h1 / diag; dp[i] = rhs / diag; }
            else { double den = diag - h0 * cp[i - 1]; cp[i] = h1 / den; dp[i] = (rhs - h0 * dp[i - 1]) / den; }
        }
        m2[n - 2] = dp[n - 2];
        for (int i = n - 3; i >= 1; --i) m2[i] = dp[i] - cp[i] * m2[i + 1];
    }
    xr.resize(n); yr.resize(n);
    int seg = 0;
    for (int q = 0; q < n; ++q) {
        double t = (q == n - 1) ? x[n - 1] : x[0] + (x[n - 1] - x[0]) * (double)q / (double)(n - 1);
        xr[q] = t;
        while (seg < n - 2 && t >= x[seg + 1]) ++seg;
        double h = x[seg + 1] - x[seg], A = (x[seg + 1] - t) / h, B = (t - x[seg]) / h;
        yr[q] = A * y[seg] + B * y[seg + 1] + ((A * A * A - A) * m2[seg] + (B * B * B - B) * m2[seg + 1]) * h * h / 6.0;
    }
}

static std::vector<double> thomas_pivots(int n) {  // cp[k] of the (1/6, 2/3, 1/6) system with n-2 unknowns
    std::vector<double> cp(std::max(n, 2));
    const double a = 1.0 / 6.0, b = 2.0 / 3.0;
    cp[0] = a / b;
    for (int k = 1; k < n; ++k) cp[k] = a / (b - a * cp[k - 1]);
    return cp;
}

int torj_plasma_create_from_data(torj_ctx* c, const torj_grid* g, const double* psi_norm, const double* psi_prof,
                                 const double* ne_prof, const double* Te_prof, int32_t n_prof, const double* BR,
                                 const double* BZ, const double* Bphi, const double* psi_1d, const double* vol_1d,
                                 int32_t n_1d, torj_plasma** out) {
    if (!c || !g || !out) FAIL("torj_plasma_create_from_data: NULL argument");
    if (g->nR < 2 || g->nZ < 2 || n_prof < 4 || n_1d < 2) FAIL("torj_plasma_create_from_data: grid or profile too small");
    for (int i = 1; i < n_prof; ++i) if (!(psi_prof[i] > psi_prof[i - 1])) FAIL("torj_plasma_create_from_data: psi_prof must increase");
    if (set_device(c)) return 1;
    cudaStream_t st = c->stream;
    const int nR = g->nR, nZ = g->nZ, sr = nR + 2;
    const size_t npts = (size_t)nR * nZ, nodes = (size_t)sr * (nZ + 2);
    // ---- 1-D host work (tiny): resample, log, prefilter (src/plasma.jl:16-19,42-44)
    std::vector<double> xr, yr, c1[2];
    double x0[2], hh[2];
    const double* profs[2] = {ne_prof, Te_prof};
    for (int q = 0; q < 2; ++q) {
        resample_uniform(psi_prof, profs[q], n_prof, xr, yr);
        for (auto& v : yr) v = std::log(v);
        c1[q].resize(n_prof + 2);
        prefilter_line(yr.data(), 1, n_prof, c1[q].data(), 1);
        x0[q] = xr[0]; hh[q] = (xr[n_prof - 1] - xr[0]) / (double)(n_prof - 1);
    }
    std::vector<double> vr, vv, vc(n_1d + 2);
    resample_uniform(psi_1d, vol_1d, n_1d, vr, vv);
    prefilter_line(vv.data(), 1, n_1d, vc.data(), 1);
    double psi_prof_max = psi_prof[0];
    for (int i = 1; i < n_prof; ++i) psi_prof_max = std::max(psi_prof_max, psi_prof[i]);
    // ---- device: node data of six fields, R pass, Z pass, pack
    dev_ptr<double> o_raw, o_tmp, o_coef, o_c1, o_cpR, o_cpZ;
    CK(dev_alloc(o_raw, 6 * npts));
    CK(dev_alloc(o_tmp, (size_t)sr * nZ));
    CK(dev_alloc(o_coef, 6 * nodes));
    CK(dev_alloc(o_c1, 2 * (size_t)(n_prof + 2)));
    std::vector<double> cpR = thomas_pivots(nR), cpZ = thomas_pivots(nZ);
    CK(dev_alloc(o_cpR, cpR.size()));
    CK(dev_alloc(o_cpZ, cpZ.size()));
    double *d_raw = o_raw.get(), *d_tmp = o_tmp.get(), *d_coef = o_coef.get(), *d_c1 = o_c1.get(), *d_cpR = o_cpR.get(),
           *d_cpZ = o_cpZ.get();
    const double* src[6] = {psi_norm, nullptr, nullptr, BR, BZ, Bphi};  // field order: psi, lnne, lnTe, BR, BZ, Bphi
    for (int f = 0; f < 6; ++f)
        if (src[f]) CK(cudaMemcpyAsync(d_raw + f * npts, src[f], npts * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int q = 0; q < 2; ++q)
        CK(cudaMemcpyAsync(d_c1 + q * (n_prof + 2), c1[q].data(), (n_prof + 2) * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cpR, cpR.data(), cpR.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cpZ, cpZ.data(), cpZ.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int q = 0; q < 2; ++q) {
        k_profile_to_grid<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(d_raw, (long long)npts, d_c1 + q * (n_prof + 2), n_prof,
                                                                       x0[q], hh[q], d_raw + (1 + q) * npts);
        c->launches++;
    }
    for (int f = 0; f < 6; ++f) {
        // R pass: nZ lines of nR points (contiguous) -> tmp[(nR+2) x nZ]; Z pass: nR+2 columns of nZ points -> coef
        k_prefilter_lines<<<(nZ + 127) / 128, 128, 0, st>>>(d_raw + f * npts, d_tmp, nR, nZ, nR, 1, sr, 1, d_cpR);
        k_prefilter_lines<<<(sr + 127) / 128, 128, 0, st>>>(d_tmp, d_coef + f * nodes, nZ, sr, 1, sr, 1, sr, d_cpZ);
        c->launches += 2;
    }
    CK(cudaGetLastError());
    torj_plasma* p = new torj_plasma();
    Guard<torj_plasma, torj_plasma_destroy> guard(p);
    p->ctx = c;
    p->id = ++g_plasma_serial;
    CK(cudaMalloc(&p->dA, 2 * nodes * sizeof(double2)));
    CK(cudaMalloc(&p->dB, nodes * sizeof(double2)));
    k_pack_tables<<<(unsigned)((nodes + 255) / 256), 256, 0, st>>>(d_coef, d_coef + nodes, d_coef + 2 * nodes, d_coef + 3 * nodes,
                                                                 d_coef + 4 * nodes, d_coef + 5 * nodes, (long long)nodes, p->dA, p->dB);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    DevTables& T = p->T;
    T.A = p->dA; T.B = p->dB;
    T.nR = nR; T.nZ = nZ; T.row = sr;
    T.r0 = g->R_first; T.z0 = g->Z_first;
    double hr = (g->R_last - g->R_first) / (double)(nR - 1), hz = (g->Z_last - g->Z_first) / (double)(nZ - 1);
    T.inv_hr = 1.0 / hr; T.inv_hz = 1.0 / hz;
    T.rlast = g->R_first + hr * (nR - 1); T.zlast = g->Z_first + hz * (nZ - 1);
    T.psi_prof_max = psi_prof_max;
    CK(cudaMalloc(&p->dT, sizeof(DevTables)));
    CK(cudaMemcpy(p->dT, &T, sizeof(DevTables), cudaMemcpyHostToDevice));
    p->vol_c = vc;
    p->n_vol = n_1d; p->vol_x0 = vr[0]; p->vol_h = (vr[n_1d - 1] - vr[0]) / (double)(n_1d - 1);
    *out = guard.release();
    return 0;
}

void torj_plasma_destroy(torj_plasma* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaFree(p->dA);
    cudaFree(p->dB);
    cudaFree(p->dT);
    delete p;
}

// V(psi_N): 1-D cubic B-spline with Line extrapolation (reference src/plasma.jl:44,109,113)
static double volume_at(const torj_plasma* p, double x) {
    const int n = p->n_vol;
    double xl = p->vol_x0 + p->vol_h * (n - 1);
    double xc = std::min(std::max(x, p->vol_x0), xl);
    double u = (xc - p->vol_x0) / p->vol_h + 1.0;
    int i = (int)std::floor(u);
    i = std::min(std::max(i, 1), n - 1);
    double d = u - i, e = 1.0 - d;
    double w[4] = {e * e * e / 6.0, 2.0 / 3.0 - d * d + d * d * d / 2.0, 2.0 / 3.0 - e * e + e * e * e / 2.0, d * d * d / 6.0};
    double dw[4] = {-e * e / 2.0 / p->vol_h, (-2.0 * d + 1.5 * d * d) / p->vol_h, (2.0 * e - 1.5 * e * e) / p->vol_h,
                    d * d / 2.0 / p->vol_h};
    double v = 0, dv = 0;
    for (int k = 0; k < 4; ++k) { v += w[k] * p->vol_c[i - 1 + k]; dv += dw[k] * p->vol_c[i - 1 + k]; }
    if (xc != x) v += (x - xc) * dv;
    return v;
}

static SolverOpts to_sopts(const torj_options* o, double s_max) {
    torj_options d;
    torj_options_default(&d);
    if (o) d = *o;
    SolverOpts s;
    s.n_segments = d.n_segments; s.max_steps = d.max_steps_per_segment; s.s_max = s_max; s.dtmax = d.dtmax;
    s.abstol = d.abstol; s.reltol = d.reltol; s.psi_stop = d.psi_stop; s.p_stop = d.p_stop; s.te_min = d.te_min;
    s.max_harmonic = d.max_harmonic;
    s.alpha_floor = d.alpha_floor;
    return s;
}

int torj_probe(torj_ctx* c, const torj_plasma* p, const torj_options* opt, int64_t n, const double* x, const double* N,
               double freq_hz, int32_t mode, double* out) {
    if (!c->gl_set) FAIL("The weights and abscissae for the absorption were never initialized. Call `abs_Al_init` before using the absorption.");
    if (set_device(c)) return 1;
    SolverOpts so = to_sopts(opt, 1.0);
    dev_ptr<double> o_x, o_N, o_out;
    CK(dev_alloc(o_x, 3 * (size_t)n));
    CK(dev_alloc(o_N, 3 * (size_t)n));
    CK(dev_alloc(o_out, 11 * (size_t)n));
    double *dx = o_x.get(), *dN = o_N.get(), *dout = o_out.get();
    CK(cudaMemcpyAsync(dx, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dN, N, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_probe<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(p->T, n, dx, dN, freq_hz, mode, so.te_min, so.max_harmonic, so.alpha_floor, dout);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, 11 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int torj_rhs(torj_ctx* c, const torj_plasma* p, const torj_options* opt, int64_t n, const double* u, double freq_hz,
             int32_t mode, double* du) {
    if (!c->gl_set) FAIL("The weights and abscissae for the absorption were never initialized. Call `abs_Al_init` before using the absorption.");
    if (set_device(c)) return 1;
    SolverOpts so = to_sopts(opt, 1.0);
    dev_ptr<double> o_u, o_du;
    CK(dev_alloc(o_u, 7 * (size_t)n));
    CK(dev_alloc(o_du, 7 * (size_t)n));
    double *d_u = o_u.get(), *d_du = o_du.get();
    CK(cudaMemcpyAsync(d_u, u, 7 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_rhs<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(p->T, n, d_u, freq_hz, mode, so.te_min, so.max_harmonic, so.alpha_floor, d_du);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(du, d_du, 7 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- bundle ---------------------------------------------------------------------------------------------------
static void free_traj(torj_bundle* b) {
    cudaFree(b->d_ts); cudaFree(b->d_txyz); cudaFree(b->d_tP); cudaFree(b->d_tdP); cudaFree(b->d_tprof);
    b->d_ts = b->d_txyz = b->d_tP = b->d_tdP = b->d_tprof = nullptr;
    b->traj_prof_npsi = 0;
}

static int bundle_upload(torj_bundle* b, const double* pos, const double* dir, const double* weight,
                         const double* freq_hz, const int32_t* mode) {
    torj_ctx* c = b->ctx;
    const int64_t n = b->n;
    size_t nf = b->per_ray_fm ? (size_t)n : 1;
    CK(cudaMemcpyAsync(b->d_pos, pos, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(b->d_dir, dir, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(b->d_w, weight, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(b->d_freq, freq_hz, nf * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(b->d_mode, mode, nf * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // host buffers may be released by the caller on return
    return 0;
}

int torj_bundle_create(torj_ctx* c, int64_t n, const double* pos, const double* dir, const double* weight,
                       const double* freq_hz, const int32_t* mode, int32_t per_ray_fm, torj_bundle** out) {
    if (!c || !out || n < 1) FAIL("torj_bundle_create: bad argument");
    if (set_device(c)) return 1;
    torj_bundle* b = new torj_bundle();
    Guard<torj_bundle, torj_bundle_destroy> guard(b);
    b->ctx = c; b->n = n; b->per_ray_fm = per_ray_fm;
    size_t nf = per_ray_fm ? (size_t)n : 1;
    CK(cudaMalloc(&b->d_pos, 3 * n * sizeof(double)));
    CK(cudaMalloc(&b->d_dir, 3 * n * sizeof(double)));
    CK(cudaMalloc(&b->d_w, n * sizeof(double)));
    CK(cudaMalloc(&b->d_freq, nf * sizeof(double)));
    CK(cudaMalloc(&b->d_mode, nf * sizeof(int)));
    CK(cudaMalloc(&b->d_u0, 7 * n * sizeof(double)));
    CK(cudaMalloc(&b->d_s0, n * sizeof(double)));
    CK(cudaMalloc(&b->d_psil, n * sizeof(double)));
    CK(cudaMalloc(&b->d_Pf, n * sizeof(double)));
    CK(cudaMalloc(&b->d_Pdep, n * sizeof(double)));
    CK(cudaMalloc(&b->d_status, n * sizeof(int)));
    CK(cudaMalloc(&b->d_npts, n * sizeof(int)));
    CK(cudaMalloc(&b->d_queue, sizeof(unsigned long long)));
    CK(cudaMalloc(&b->d_counters, 8 * sizeof(unsigned long long)));
    BundleDev& B = b->B;
    B.n_rays = n; B.pos = b->d_pos; B.dir = b->d_dir; B.weight = b->d_w; B.freq = b->d_freq; B.mode = b->d_mode;
    B.per_ray_fm = per_ray_fm; B.u0 = b->d_u0; B.s0 = b->d_s0; B.psi_launch = b->d_psil; B.status = b->d_status;
    B.P_final = b->d_Pf; B.P_dep = b->d_Pdep; B.n_points = b->d_npts;
    if (bundle_upload(b, pos, dir, weight, freq_hz, mode)) return 1;
    *out = guard.release();
    return 0;
}

// Bundle generated on the device from launcher parameters: launch_peripheral_rays (reference src/launch.jl:24-132) for
// n_launchers beams at once; ray k of launcher L is ray L*n_per+k, its beam index is L, frequency and mode are per ray.
// gh_nodes/gh_weights: Gauss-Hermite rule of order 2*N_rings+2 (FastGaussQuadrature.gausshermite, src/launch.jl:72).
int torj_bundle_create_from_launchers(torj_ctx* c, int32_t n_launchers, const double* x0, const double* N0, const double* w,
                                      const double* inv_Rc, const double* f, const int32_t* mode, int32_t N_rings,
                                      int32_t min_azimuthal_points, int32_t normalize_weight_sum, const double* gh_nodes,
                                      const double* gh_weights, torj_bundle** out, int64_t* n_rays_out) {
    if (!c || !out || n_launchers < 1) FAIL("torj_bundle_create_from_launchers: bad argument");
    if (N_rings < 2) FAIL("N_rings < 2 which is the minimum");  // ArgumentError of src/launch.jl:27-29
    if (set_device(c)) return 1;
    // ring radii in units of w/sqrt(2): v[N_rings+2:end] of the ascending rule, first N_rings used (src/launch.jl:72-83)
    std::vector<double> ru(N_rings), rwu(N_rings);
    std::vector<int> start(N_rings + 1, 0);
    double wsum = 0.0;
    for (int i = 0; i < N_rings; ++i) {
        ru[i] = gh_nodes[N_rings + 1 + i];
        rwu[i] = gh_weights[N_rings + 1 + i];
        long nth = std::max(1L, (long)std::nearbyint((double)min_azimuthal_points * ru[i] / ru[0]));
        start[i + 1] = start[i] + (int)nth;
    }
    // sum of the launcher's weights, accumulated ray by ray in launch order like `sum(ray_weights)` does
    for (int i = 0; i < N_rings; ++i) {
        int nth = start[i + 1] - start[i];
        double wt = ru[i] * rwu[i] * (2.0 * M_PI / (double)nth);
        for (int j = 0; j < nth; ++j) wsum += wt;
    }
    const int n_per = start[N_rings];
    const int64_t n = (int64_t)n_launchers * n_per;
    // an empty bundle shell with device arrays, then the generator kernel fills pos/dir/weight/freq/mode/beam
    std::vector<double> zero3(3, 0.0);
    torj_bundle* b = new torj_bundle();
    Guard<torj_bundle, torj_bundle_destroy> guard(b);
    b->ctx = c; b->n = n; b->per_ray_fm = 1;
    CK(cudaMalloc(&b->d_pos, 3 * n * sizeof(double)));
    CK(cudaMalloc(&b->d_dir, 3 * n * sizeof(double)));
    CK(cudaMalloc(&b->d_w, n * sizeof(double)));
    CK(cudaMalloc(&b->d_freq, n * sizeof(double)));
    CK(cudaMalloc(&b->d_mode, n * sizeof(int)));
    CK(cudaMalloc(&b->d_u0, 7 * n * sizeof(double)));
    CK(cudaMalloc(&b->d_s0, n * sizeof(double)));
    CK(cudaMalloc(&b->d_psil, n * sizeof(double)));
    CK(cudaMalloc(&b->d_Pf, n * sizeof(double)));
    CK(cudaMalloc(&b->d_Pdep, n * sizeof(double)));
    CK(cudaMalloc(&b->d_status, n * sizeof(int)));
    CK(cudaMalloc(&b->d_npts, n * sizeof(int)));
    CK(cudaMalloc(&b->d_queue, sizeof(unsigned long long)));
    CK(cudaMalloc(&b->d_counters, 8 * sizeof(unsigned long long)));
    CK(cudaMalloc(&b->d_beam, n * sizeof(int)));
    b->n_beams = n_launchers;
    BundleDev& B = b->B;
    B.n_rays = n; B.pos = b->d_pos; B.dir = b->d_dir; B.weight = b->d_w; B.freq = b->d_freq; B.mode = b->d_mode;
    B.per_ray_fm = 1; B.u0 = b->d_u0; B.s0 = b->d_s0; B.psi_launch = b->d_psil; B.status = b->d_status;
    B.P_final = b->d_Pf; B.P_dep = b->d_Pdep; B.n_points = b->d_npts;
    // launcher parameters to the device
    const size_t nl = n_launchers;
    dev_ptr<double> o_par; dev_ptr<int> o_ipar;
    CK(dev_alloc(o_par, 9 * nl + 2 * N_rings));
    CK(dev_alloc(o_ipar, nl + N_rings + 1));
    double* d_par = o_par.get(); int* d_ipar = o_ipar.get();
    cudaStream_t st = c->stream;
    CK(cudaMemcpyAsync(d_par, x0, 3 * nl * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_par + 3 * nl, N0, 3 * nl * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_par + 6 * nl, w, nl * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_par + 7 * nl, inv_Rc, nl * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_par + 8 * nl, f, nl * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_par + 9 * nl, ru.data(), N_rings * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_par + 9 * nl + N_rings, rwu.data(), N_rings * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_ipar, mode, nl * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_ipar + nl, start.data(), (N_rings + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    LaunchArgs A;
    A.n_launchers = n_launchers; A.n_rings = N_rings; A.n_per = n_per;
    A.x0 = d_par; A.N0 = d_par + 3 * nl; A.w = d_par + 6 * nl; A.inv_Rc = d_par + 7 * nl; A.f = d_par + 8 * nl;
    A.r_unit = d_par + 9 * nl; A.rw_unit = d_par + 9 * nl + N_rings;
    A.mode = d_ipar; A.ring_start = d_ipar + nl;
    A.wsum_unit = wsum; A.normalize = normalize_weight_sum;
    A.pos = b->d_pos; A.dir = b->d_dir; A.weight = b->d_w; A.freq = b->d_freq; A.mode_out = b->d_mode; A.beam = b->d_beam;
    k_launch_rays<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(A);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    *out = guard.release();
    if (n_rays_out) *n_rays_out = n;
    return 0;
}

// copy the generated (or uploaded) launch arrays back: pos/dir [3][n], weight [n] (any pointer may be NULL)
int torj_bundle_rays(torj_bundle* b, double* pos, double* dir, double* weight) {
    torj_ctx* c = b->ctx;
    if (set_device(c)) return 1;
    if (pos) CK(cudaMemcpyAsync(pos, b->d_pos, 3 * b->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (dir) CK(cudaMemcpyAsync(dir, b->d_dir, 3 * b->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (weight) CK(cudaMemcpyAsync(weight, b->d_w, b->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

void torj_bundle_destroy(torj_bundle* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->d_pos); cudaFree(b->d_dir); cudaFree(b->d_w); cudaFree(b->d_freq); cudaFree(b->d_mode); cudaFree(b->d_u0);
    cudaFree(b->d_s0); cudaFree(b->d_psil); cudaFree(b->d_Pf); cudaFree(b->d_Pdep); cudaFree(b->d_status); cudaFree(b->d_npts);
    cudaFree(b->d_hand); cudaFree(b->d_segdone); cudaFree(b->d_left);
    cudaFree(b->d_queue); cudaFree(b->d_counters); cudaFree(b->d_edges); cudaFree(b->d_bins); cudaFree(b->d_dV);
    cudaFree(b->d_profile);
    cudaFree(b->d_beam);
    free_traj(b);
    delete b;
}

int torj_bundle_set_window(torj_bundle* b, int64_t first, int64_t count, int32_t max_pts) {
    if (set_device(b->ctx)) return 1;
    CK(cudaStreamSynchronize(b->ctx->stream));
    free_traj(b);
    if (count <= 0 || max_pts <= 0) { b->traj_first = 0; b->traj_count = 0; b->traj_max = 0; return 0; }
    if (first < 0 || first + count > b->n) FAIL("torj_bundle_set_window: window outside the bundle");
    b->traj_first = first; b->traj_count = count; b->traj_max = max_pts;
    size_t m = (size_t)count * max_pts;
    CK(cudaMalloc(&b->d_ts, m * sizeof(double)));
    CK(cudaMalloc(&b->d_txyz, 3 * m * sizeof(double)));
    CK(cudaMalloc(&b->d_tP, m * sizeof(double)));
    CK(cudaMalloc(&b->d_tdP, m * sizeof(double)));
    return 0;
}

int torj_bundle_set_beams(torj_bundle* b, int32_t n_beams, const int32_t* beam_id) {
    torj_ctx* c = b->ctx;
    if (n_beams < 1) FAIL("torj_bundle_set_beams: n_beams < 1");
    if (set_device(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    if (n_beams == 1 || !beam_id) {
        b->n_beams = 1;
        return 0;
    }
    for (int64_t i = 0; i < b->n; ++i)
        if (beam_id[i] < 0 || beam_id[i] >= n_beams) FAIL("torj_bundle_set_beams: beam_id out of range");
    if (!b->d_beam) CK(cudaMalloc(&b->d_beam, b->n * sizeof(int)));
    CK(cudaMemcpy(b->d_beam, beam_id, b->n * sizeof(int), cudaMemcpyHostToDevice));
    b->n_beams = n_beams;
    return 0;
}

int torj_bundle_trace(torj_bundle* b, const torj_plasma* p, const torj_options* opt, double s_max, int32_t n_psi,
                      const double* psi_edges) {
    torj_ctx* c = b->ctx;
    if (!c->gl_set) FAIL("The weights and abscissae for the absorption were never initialized. Call `abs_Al_init` before using the absorption.");
    if (n_psi < 2) FAIL("torj_bundle_trace: need at least 2 psi levels");
    for (int j = 1; j < n_psi; ++j) if (!(psi_edges[j] > psi_edges[j - 1])) FAIL("torj_bundle_trace: psi_dP_dV must be strictly increasing");
    torj_options od;
    torj_options_default(&od);
    if (opt) od = *opt;
    if (od.scheme != 0 && od.scheme != 1) FAIL("torj_bundle_trace: scheme must be 0 (Tsit5) or 1 (OwrenZen3)");
    if (od.n_segments < 1) FAIL("torj_bundle_trace: n_segments < 1");
    if (!(od.alpha_floor >= 0.0)) FAIL("torj_bundle_trace: alpha_floor must be >= 0");
    if (od.max_harmonic < 1 || od.max_harmonic > 16) FAIL("torj_bundle_trace: max_harmonic must be in 1..16 (1 = no absorption)");
    if (!(od.dtmax > 0.0) || !(od.abstol > 0.0) || !(od.reltol > 0.0)) FAIL("torj_bundle_trace: dtmax, abstol and reltol must be > 0");
    if (od.schedule < 0 || od.schedule > 2) FAIL("torj_bundle_trace: schedule must be 0 (automatic), 1 (whole rays) or 2 (segment hand-off)");
    if (od.max_steps_per_segment < 1) FAIL("torj_bundle_trace: max_steps_per_segment < 1");
    if (!(s_max > 0.0)) FAIL("torj_bundle_trace: s_max must be > 0");
    if (set_device(c)) return 1;
    cudaStream_t st = c->stream;
    if (b->n_psi != n_psi || b->prof_rows != b->n_beams) {
        CK(cudaStreamSynchronize(st));
        cudaFree(b->d_edges); cudaFree(b->d_bins); cudaFree(b->d_dV); cudaFree(b->d_profile);
        b->d_edges = b->d_bins = b->d_dV = b->d_profile = nullptr;  // a failing allocation below must not leave them dangling
        b->n_psi = 0; b->prof_rows = 0;
        b->h_edges.clear();
        CK(cudaMalloc(&b->d_edges, n_psi * sizeof(double)));
        CK(cudaMalloc(&b->d_bins, (size_t)b->n_beams * (n_psi + 2) * sizeof(double)));
        CK(cudaMalloc(&b->d_dV, n_psi * sizeof(double)));
        CK(cudaMalloc(&b->d_profile, (size_t)b->n_beams * (n_psi + 2) * sizeof(double)));
        b->n_psi = n_psi;
        b->prof_rows = b->n_beams;
    }
    if (b->traj_count > 0 && b->traj_prof_npsi != n_psi) {
        CK(cudaStreamSynchronize(st));
        cudaFree(b->d_tprof);
        b->d_tprof = nullptr; b->traj_prof_npsi = 0;
        CK(cudaMalloc(&b->d_tprof, (size_t)b->traj_count * n_psi * sizeof(double)));
        b->traj_prof_npsi = n_psi;
    }
    if (b->edges_plasma != p->id || (int)b->h_edges.size() != n_psi ||
        memcmp(b->h_edges.data(), psi_edges, n_psi * sizeof(double)) != 0) {
        std::vector<double> dV(n_psi, 1.0);
        for (int j = 0; j + 1 < n_psi; ++j) dV[j] = volume_at(p, psi_edges[j + 1]) - volume_at(p, psi_edges[j]);
        CK(cudaMemcpyAsync(b->d_edges, psi_edges, n_psi * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b->d_dV, dV.data(), n_psi * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));  // dV is a local; psi_edges belongs to the caller
        b->h_edges.assign(psi_edges, psi_edges + n_psi);
        b->edges_plasma = p->id;
    }
    CK(cudaMemsetAsync(b->d_bins, 0, (size_t)b->n_beams * (n_psi + 2) * sizeof(double), st));
    CK(cudaMemsetAsync(b->d_queue, 0, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(b->d_counters, 0, 8 * sizeof(unsigned long long), st));
    if (b->traj_count > 0) CK(cudaMemsetAsync(b->d_tprof, 0, (size_t)b->traj_count * n_psi * sizeof(double), st));

    SolverOpts so = to_sopts(&od, s_max);
    k_ray_init<<<(unsigned)((b->n + 127) / 128), 128, 0, st>>>(p->T, b->B, so);
    c->launches++;
    CK(cudaGetLastError());

    TraceArgs a;
    a.T = p->T; a.B = b->B; a.O = so;
    a.J.first = b->traj_first; a.J.count = b->traj_count; a.J.max_pts = b->traj_max;
    a.J.s = b->d_ts; a.J.xyz = b->d_txyz; a.J.P = b->d_tP; a.J.dP = b->d_tdP; a.J.prof = b->d_tprof;
    a.n_psi = n_psi; a.n_beams = b->n_beams; a.beam_id = b->d_beam; a.psi_edges = b->d_edges; a.bins = b->d_bins; a.next_ray = b->d_queue; a.counters = b->d_counters;
    size_t smem = (size_t)n_psi * sizeof(double);
#if TORJ_K_SMEM
    smem += (size_t)7 * 7 * TORJ_TPB * sizeof(double);
#endif
#if TORJ_PARK
    smem += (size_t)TORJ_PARK_SLOTS * TORJ_TPB * sizeof(double);
#endif
    int bps = 0;
    int64_t warps = (b->n + 31) / 32;
    int64_t blocks_needed = (warps + (TORJ_TPB / 32) - 1) / (TORJ_TPB / 32);
    // harmonics above the third run in their own instantiations, so the default kernels do not carry the code
    void (*kern)(TraceArgs) = od.max_harmonic > 3 ? (od.scheme == 0 ? k_trace<0, true> : k_trace<1, true>)
                                                  : (od.scheme == 0 ? k_trace<0, false> : k_trace<1, false>);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TORJ_TPB, smem));
    if (bps < 1) FAIL("torj_bundle_trace: trace kernel does not fit on an SM (n_psi too large for shared memory)");
    int64_t grid = std::min<int64_t>((int64_t)c->num_sms * bps, blocks_needed);  // persistent: resident CTAs only
    // Segment hand-off whenever the bundle exceeds the resident lanes. It removes the partly filled last wave (65 543
    // rays on 37 888 lanes: 1.73 ray-times instead of 2) and keeps the lanes of a warp in step: with whole rays per lane
    // a bundle of rays of very different lengths (the 1 M-ray angle sweep) drifts apart and some lane pays for the full
    // absorption coefficient on every trip (measured 3.1 s -> 2.4 s). Below one wave there is nothing to balance.
    const int64_t lanes = (int64_t)c->num_sms * bps * TORJ_TPB;
    int interleave = od.schedule == 2 || (od.schedule == 0 && b->n > lanes);
    if (od.n_segments < 2 || b->n > 0x7fffffff) interleave = 0;
    a.interleave = interleave; a.hand = nullptr; a.seg_done = nullptr; a.rays_left = nullptr;
    if (interleave) {
        if (!b->d_hand) CK(cudaMalloc(&b->d_hand, (size_t)b->n * TORJ_HAND_D * sizeof(double)));
        if (!b->d_segdone) CK(cudaMalloc(&b->d_segdone, (size_t)b->n * sizeof(int)));
        if (!b->d_left) CK(cudaMalloc(&b->d_left, sizeof(int)));
        b->h_left = (int)b->n;
        CK(cudaMemsetAsync(b->d_segdone, 0, (size_t)b->n * sizeof(int), st));
        CK(cudaMemcpyAsync(b->d_left, &b->h_left, sizeof(int), cudaMemcpyHostToDevice, st));
        a.hand = b->d_hand; a.seg_done = b->d_segdone; a.rays_left = b->d_left;
    }
    CK(cudaEventRecord(c->ev0, st));
    kern<<<(unsigned)grid, TORJ_TPB, smem, st>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
    k_finalize<<<(unsigned)(((size_t)b->n_beams * (n_psi + 2) + 127) / 128), 128, 0, st>>>(b->d_bins, b->d_dV, n_psi, b->n_beams,
                                                                                         b->d_profile);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

void* torj_bundle_device_profile(torj_bundle* b) { return b->d_profile; }

int torj_bundle_results(torj_bundle* b, double* dP_dV, double* deposited_power, double* P_final, double* P_dep,
                        int32_t* n_points, int32_t* status, torj_counters* counters) {
    torj_ctx* c = b->ctx;
    if (set_device(c)) return 1;
    cudaStream_t st = c->stream;
    if (b->n_psi == 0) FAIL("torj_bundle_results: nothing traced yet");
    const size_t row = (size_t)b->n_psi + 2;
    std::vector<double> prof(row * b->n_beams);
    CK(cudaMemcpyAsync(prof.data(), b->d_profile, prof.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (P_final) CK(cudaMemcpyAsync(P_final, b->d_Pf, b->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (P_dep) CK(cudaMemcpyAsync(P_dep, b->d_Pdep, b->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (n_points) CK(cudaMemcpyAsync(n_points, b->d_npts, b->n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (status) CK(cudaMemcpyAsync(status, b->d_status, b->n * sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned long long cn[8];
    CK(cudaMemcpyAsync(cn, b->d_counters, sizeof cn, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int q = 0; q < b->n_beams; ++q) {
        if (dP_dV) memcpy(dP_dV + (size_t)q * b->n_psi, prof.data() + q * row, b->n_psi * sizeof(double));
        if (deposited_power) deposited_power[q] = prof[q * row + b->n_psi];
    }
    if (counters) {
        counters->n_acc = (int64_t)cn[0]; counters->n_rej = (int64_t)cn[1]; counters->n_rhs = (int64_t)cn[2];
        counters->n_alpha = (int64_t)cn[3]; counters->n_harm = (int64_t)cn[4]; counters->n_rays_ok = (int64_t)cn[5];
        counters->n_harm_pruned = (int64_t)cn[6];
        counters->n_alpha_skipped = (int64_t)cn[7];
    }
    return 0;
}

int torj_bundle_trajectories(torj_bundle* b, double* s, double* xyz, double* P, double* dP_ds, double* dP_dV_ray) {
    torj_ctx* c = b->ctx;
    if (b->traj_count <= 0) FAIL("torj_bundle_trajectories: no trajectory window set");
    if (set_device(c)) return 1;
    cudaStream_t st = c->stream;
    size_t m = (size_t)b->traj_count * b->traj_max;
    if (s) CK(cudaMemcpyAsync(s, b->d_ts, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (xyz) CK(cudaMemcpyAsync(xyz, b->d_txyz, 3 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (P) CK(cudaMemcpyAsync(P, b->d_tP, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dP_ds) CK(cudaMemcpyAsync(dP_ds, b->d_tdP, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dP_dV_ray && b->d_tprof)
        CK(cudaMemcpyAsync(dP_dV_ray, b->d_tprof, (size_t)b->traj_count * b->n_psi * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (dP_dV_ray) {  // shell power -> dP/dV per ray (reference src/plasma.jl:141)
        std::vector<double> dV(b->n_psi);
        CK(cudaMemcpy(dV.data(), b->d_dV, b->n_psi * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t r = 0; r < b->traj_count; ++r) {
            double* row = dP_dV_ray + (size_t)r * b->n_psi;
            for (int j = 0; j + 1 < b->n_psi; ++j) row[j] /= dV[j];
            row[b->n_psi - 1] = 0.0;
        }
    }
    return 0;
}

int torj_trace(torj_ctx* c, const torj_plasma* p, const torj_options* opt, int64_t n_rays, const double* pos,
               const double* dir, const double* weight, const double* freq_hz, const int32_t* mode, int32_t per_ray_fm,
               double s_max, int32_t n_psi, const double* psi_edges, int32_t n_beams, const int32_t* beam_id, double* dP_dV,
               double* deposited_power, double* P_final, double* P_dep, int32_t* n_points, int32_t* status, int64_t traj_first,
               int64_t traj_count,
               int32_t traj_max_pts, double* traj_s, double* traj_xyz, double* traj_P, double* traj_dP_ds,
               double* traj_dP_dV_ray, torj_counters* counters) {
    // the device workspace of the previous call is reused when the bundle shape is unchanged (no cudaMalloc/cudaFree
    // on the repeated-call path of make_beam scans)
    int rc = 0;
    torj_bundle* b = c->ws;
    if (b && (b->n != n_rays || b->per_ray_fm != per_ray_fm)) {
        torj_bundle_destroy(b);
        b = c->ws = nullptr;
    }
    if (!b) {
        rc = torj_bundle_create(c, n_rays, pos, dir, weight, freq_hz, mode, per_ray_fm, &b);
        if (rc) return rc;
        c->ws = b;
    } else {
        if (set_device(c)) return 1;
        rc = bundle_upload(b, pos, dir, weight, freq_hz, mode);
        if (rc) return rc;
    }
    if (traj_count != b->traj_count || traj_first != b->traj_first || (traj_count > 0 && traj_max_pts != b->traj_max))
        rc = torj_bundle_set_window(b, traj_first, traj_count, traj_max_pts);
    if (!rc) rc = torj_bundle_set_beams(b, n_beams, beam_id);
    if (!rc) rc = torj_bundle_trace(b, p, opt, s_max, n_psi, psi_edges);
    if (!rc) rc = torj_bundle_results(b, dP_dV, deposited_power, P_final, P_dep, n_points, status, counters);
    if (!rc && traj_count > 0) rc = torj_bundle_trajectories(b, traj_s, traj_xyz, traj_P, traj_dP_ds, traj_dP_dV_ray);
    if (rc) {  // do not keep a workspace in an unknown state
        torj_bundle_destroy(b);
        c->ws = nullptr;
    }
    return rc;
}

// ---- single-process multi-GPU front end (what a Julia make_beam on an 8-GPU box calls) -------------------------
struct torj_multi {
    std::vector<torj_ctx*> ctx;
};
struct torj_mplasma {
    std::vector<torj_plasma*> p;
};

int torj_multi_create(int32_t n_devices, torj_multi** out) {
    if (!out) FAIL("torj_multi_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_err = std::string("torj_multi_create: no CUDA device (") + cudaGetErrorString(e) + "); libtorj_cuda has no CPU path";
        return 1;
    }
    if (n_devices <= 0 || n_devices > ndev) n_devices = ndev;
    torj_multi* m = new torj_multi();
    for (int d = 0; d < n_devices; ++d) {
        torj_ctx* c = nullptr;
        if (torj_ctx_create(d, nullptr, &c)) { for (auto* q : m->ctx) torj_ctx_destroy(q); delete m; return 1; }
        m->ctx.push_back(c);
    }
    *out = m;
    return 0;
}

void torj_multi_destroy(torj_multi* m) {
    if (!m) return;
    for (auto* c : m->ctx) torj_ctx_destroy(c);
    delete m;
}

int32_t torj_multi_device_count(const torj_multi* m) { return (int32_t)m->ctx.size(); }

int torj_multi_abs_init(torj_multi* m, int32_t n, const double* nodes, const double* weights) {
    for (auto* c : m->ctx)
        if (int rc = torj_abs_init(c, n, nodes, weights)) return rc;
    return 0;
}

int torj_multi_plasma_create_from_data(torj_multi* m, const torj_grid* g, const double* psi_norm, const double* psi_prof,
                                       const double* ne_prof, const double* Te_prof, int32_t n_prof, const double* BR,
                                       const double* BZ, const double* Bphi, const double* psi_1d, const double* vol_1d,
                                       int32_t n_1d, torj_mplasma** out) {
    if (!out) FAIL("torj_multi_plasma_create_from_data: out is NULL");
    torj_mplasma* mp = new torj_mplasma();
    for (auto* c : m->ctx) {  // the tables (3.2 MB for 257x257) are replicated on every device
        torj_plasma* p = nullptr;
        int rc = torj_plasma_create_from_data(c, g, psi_norm, psi_prof, ne_prof, Te_prof, n_prof, BR, BZ, Bphi, psi_1d, vol_1d,
                                              n_1d, &p);
        if (rc) { for (auto* q : mp->p) torj_plasma_destroy(q); delete mp; return rc; }
        mp->p.push_back(p);
    }
    *out = mp;
    return 0;
}

void torj_multi_plasma_destroy(torj_mplasma* mp) {
    if (!mp) return;
    for (auto* p : mp->p) torj_plasma_destroy(p);
    delete mp;
}

// Rays are split into contiguous blocks, one per device, traced concurrently by one host thread per device; profiles
// and deposited power are summed on the host in device order (deterministic), per-ray outputs land in their slices.
int torj_multi_trace(torj_multi* m, const torj_mplasma* mp, const torj_options* opt, int64_t n_rays, const double* pos,
                     const double* dir, const double* weight, const double* freq_hz, const int32_t* mode, int32_t per_ray_fm,
                     double s_max, int32_t n_psi, const double* psi_edges, int32_t n_beams, const int32_t* beam_id,
                     double* dP_dV, double* deposited_power, double* P_final, double* P_dep, int32_t* n_points,
                     int32_t* status, int64_t traj_first, int64_t traj_count, int32_t traj_max_pts, double* traj_s,
                     double* traj_xyz, double* traj_P, double* traj_dP_ds, double* traj_dP_dV_ray, torj_counters* counters) {
    const int nd = (int)m->ctx.size();
    if ((int)mp->p.size() != nd) FAIL("torj_multi_trace: plasma was created on a different device set");
    if (n_rays < 1) FAIL("torj_multi_trace: n_rays < 1");
    if (n_beams < 1 || !beam_id) n_beams = 1;
    const int used = (int)std::min<int64_t>(nd, n_rays);
    std::vector<int> rcs(used, 0);
    std::vector<std::string> errs(used);
    std::vector<std::vector<double>> prof(used), dep(used);
    std::vector<torj_counters> cnts(used);
    std::vector<std::thread> th;
    const int64_t base = n_rays / used, rem = n_rays % used;
    for (int d = 0; d < used; ++d) {
        th.emplace_back([&, d]() {
            const int64_t lo = d * base + std::min<int64_t>(d, rem), n = base + (d < rem ? 1 : 0);
            std::vector<double> p3(3 * n), d3(3 * n);
            for (int c = 0; c < 3; ++c) {
                memcpy(p3.data() + c * n, pos + c * n_rays + lo, n * sizeof(double));
                memcpy(d3.data() + c * n, dir + c * n_rays + lo, n * sizeof(double));
            }
            prof[d].assign((size_t)n_beams * n_psi, 0.0);
            dep[d].assign(n_beams, 0.0);
            // the part of the trajectory window that falls into this block
            int64_t w0 = std::max(traj_first, lo), w1 = std::min(traj_first + traj_count, lo + n);
            int64_t wc = std::max<int64_t>(0, w1 - w0), wo = w0 - traj_first;
            auto off = [&](double* b, size_t row) { return (b && wc > 0) ? b + (size_t)wo * row : nullptr; };
            memset(&cnts[d], 0, sizeof(torj_counters));
            rcs[d] = torj_trace(m->ctx[d], mp->p[d], opt, n, p3.data(), d3.data(), weight + lo,
                                per_ray_fm ? freq_hz + lo : freq_hz, per_ray_fm ? mode + lo : mode, per_ray_fm, s_max, n_psi,
                                psi_edges, n_beams, (n_beams > 1) ? beam_id + lo : nullptr, prof[d].data(), dep[d].data(),
                                P_final ? P_final + lo : nullptr, P_dep ? P_dep + lo : nullptr, n_points ? n_points + lo : nullptr,
                                status ? status + lo : nullptr, wc > 0 ? w0 - lo : 0, wc, traj_max_pts,
                                off(traj_s, traj_max_pts), off(traj_xyz, 3 * (size_t)traj_max_pts), off(traj_P, traj_max_pts),
                                off(traj_dP_ds, traj_max_pts), off(traj_dP_dV_ray, n_psi), &cnts[d]);
            if (rcs[d]) errs[d] = g_err;  // g_err is thread-local
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < used; ++d)
        if (rcs[d]) { g_err = "device " + std::to_string(d) + ": " + errs[d]; return rcs[d]; }
    if (dP_dV) for (size_t k = 0; k < (size_t)n_beams * n_psi; ++k) { double s = 0.0; for (int d = 0; d < used; ++d) s += prof[d][k]; dP_dV[k] = s; }
    if (deposited_power) for (int b = 0; b < n_beams; ++b) { double s = 0.0; for (int d = 0; d < used; ++d) s += dep[d][b]; deposited_power[b] = s; }
    if (counters) {
        memset(counters, 0, sizeof(torj_counters));
        for (int d = 0; d < used; ++d) {
            counters->n_acc += cnts[d].n_acc; counters->n_rej += cnts[d].n_rej; counters->n_rhs += cnts[d].n_rhs;
            counters->n_alpha += cnts[d].n_alpha; counters->n_harm += cnts[d].n_harm; counters->n_rays_ok += cnts[d].n_rays_ok;
            counters->n_harm_pruned += cnts[d].n_harm_pruned; counters->n_alpha_skipped += cnts[d].n_alpha_skipped;
        }
    }
    return 0;
}

int torj_fp64_peak(torj_ctx* c, int32_t iters, double* tflops, double* ms) {
    if (set_device(c)) return 1;
    dev_ptr<double> o_d;
    CK(dev_alloc(o_d, 1));
    double* d = o_d.get();
    struct Ev {
        cudaEvent_t e = nullptr;
        ~Ev() { if (e) cudaEventDestroy(e); }
    } ev0, ev1;
    CK(cudaEventCreate(&ev0.e));
    CK(cudaEventCreate(&ev1.e));
    cudaEvent_t e0 = ev0.e, e1 = ev1.e;
    int blocks = c->num_sms * 8, threads = 256;
    k_dfma<<<blocks, threads, 0, c->stream>>>(iters / 8 + 1, d);  // warm-up
    c->launches++;
    CK(cudaEventRecord(e0, c->stream));
    k_dfma<<<blocks, threads, 0, c->stream>>>(iters, d);
    c->launches++;
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float t = 0;
    CK(cudaEventElapsedTime(&t, e0, e1));
    double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    if (ms) *ms = t;
    if (tflops) *tflops = flops / (t * 1e-3) / 1e12;
    return 0;
}

// cycles per dependent DFMA (single warp, single chain)
int torj_fp64_latency(torj_ctx* c, int32_t iters, double* cycles_per_dfma) {
    if (set_device(c)) return 1;
    dev_ptr<double> o_d; dev_ptr<long long> o_dc;
    CK(dev_alloc(o_d, 1));
    CK(dev_alloc(o_dc, 1));
    double* d = o_d.get(); long long* dc = o_dc.get();
    k_dfma_latency<<<1, 32, 0, c->stream>>>(iters, 1.0, d, dc);
    c->launches++;
    long long cyc = 0;
    CK(cudaMemcpyAsync(&cyc, dc, sizeof cyc, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *cycles_per_dfma = (double)cyc / (16.0 * (double)iters);
    return 0;
}

// rcp_fast / rsqrt_fast / sqrt_fast / exp_fast of the device code on host arrays: out[4][n]
int torj_math_probe(torj_ctx* c, int64_t n, const double* x, double* out) {
    if (!c || n < 1 || !x || !out) FAIL("torj_math_probe: bad argument");
    if (set_device(c)) return 1;
    dev_ptr<double> o_x, o_out;
    CK(dev_alloc(o_x, (size_t)n));
    CK(dev_alloc(o_out, 4 * (size_t)n));
    double *dx = o_x.get(), *dout = o_out.get();
    CK(cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_math_probe<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(n, dx, dout);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"

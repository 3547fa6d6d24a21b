// TEST INFRASTRUCTURE — CPU oracle for the TorJ ray-tracing hot path (C ABI for ctypes). Not part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// PARITY UNPINNED against the Julia reference (no Julia, no golden artifact in this environment); third-party
// semantics are pinned against scipy/numpy (same FITPACK Fortran as Dierckx; CubicSpline; jv; hermgauss).
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "absorption.hpp"
#include "deposition.hpp"
#include "fitpack.hpp"
#include "launch.hpp"
#include "plasma.hpp"
#include "ray.hpp"

using namespace torj_oracle;

extern "C" {

struct oracle_options {  // the reference's hard-coded solver constants (the oracle has no alpha_floor: it is exact)
    int32_t scheme;
    int32_t n_segments;
    double dtmax, abstol, reltol, psi_stop, p_stop, te_min;
    int32_t max_harmonic;
    int32_t max_steps_per_segment;
    int32_t absorption_model;  // 0 Albajar, 1 warm-plasma α (general_absorption.jl)
    int32_t reserved_;
    double alpha_floor;        // > 0: the CUDA path's two work-saving rules, for a like-for-like CPU timing
};

static Options to_opts(const oracle_options* o) {
    Options r;
    if (!o) return r;
    r.scheme = o->scheme; r.n_segments = o->n_segments; r.dtmax = o->dtmax; r.abstol = o->abstol; r.reltol = o->reltol;
    r.psi_stop = o->psi_stop; r.p_stop = o->p_stop; r.te_min = o->te_min; r.max_harmonic = o->max_harmonic;
    r.max_steps_per_segment = o->max_steps_per_segment;
    r.absorption_model = o->absorption_model;
    r.alpha_floor = o->alpha_floor;
    return r;
}

void* oracle_plasma_create(const double* R, int nR, const double* Z, int nZ, const double* psi_norm,
                           const double* psi_prof, const double* ne_prof, const double* Te_prof, int nprof,
                           const double* BR, const double* BZ, const double* Bphi, const double* psi1d,
                           const double* vol1d, int n1d) {
    Plasma* p = new Plasma();
    p->build(R, nR, Z, nZ, psi_norm, psi_prof, ne_prof, Te_prof, nprof, BR, BZ, Bphi, psi1d, vol1d, n1d);
    return p;
}
void oracle_plasma_destroy(void* h) { delete (Plasma*)h; }

static const Spline2D* pick(const Plasma* p, int which) {
    switch (which) {
        case 0: return &p->psi; case 1: return &p->lnne; case 2: return &p->lnTe;
        case 3: return &p->BR; case 4: return &p->BZ; case 5: return &p->Bphi;
    }
    return nullptr;
}
// coefficient table (nR+2)x(nZ+2), R fastest
int oracle_plasma_coefs(void* h, int which, double* out) {
    const Spline2D* s = pick((Plasma*)h, which);
    if (!s) return -1;
    std::memcpy(out, s->c.data(), s->c.size() * sizeof(double));
    return 0;
}
int oracle_volume_coefs(void* h, double* out, double* x0, double* dx) {
    Plasma* p = (Plasma*)h;
    std::memcpy(out, p->volume.c.data(), p->volume.c.size() * sizeof(double));
    *x0 = p->volume.x0; *dx = p->volume.h;
    return p->volume.n;
}
double oracle_psi_prof_max(void* h) { return ((Plasma*)h)->psi_prof_max; }
int oracle_spline_eval(void* h, int which, int n, const double* R, const double* Z, double* v, double* dR, double* dZ) {
    const Spline2D* s = pick((Plasma*)h, which);
    if (!s) return -1;
    for (int i = 0; i < n; ++i) s->eval(R[i], Z[i], &v[i], &dR[i], &dZ[i]);
    return 0;
}
void oracle_volume_eval(void* h, int n, const double* psi, double* out) {
    Plasma* p = (Plasma*)h;
    for (int i = 0; i < n; ++i) out[i] = p->volume(psi[i]);
}
// out: X, Y, N_par, b[3], Te, Lambda
void oracle_eval_plasma(void* h, const double* x, const double* N, double omega, int mode, double* out) {
    Plasma* p = (Plasma*)h;
    PlasmaPoint<double> pp = eval_plasma<double>(*p, x, N, omega);
    out[0] = pp.X; out[1] = pp.Y; out[2] = pp.N_par; out[3] = pp.b[0]; out[4] = pp.b[1]; out[5] = pp.b[2];
    out[6] = std::exp(evaluate<double>(p->lnTe, x));
    out[7] = dispersion_relation<double>(*p, x, N, omega, mode);
}
void oracle_rhs(void* h, const double* gl_t, const double* gl_w, int n_gl, const double* u, double f, int mode,
                double te_min, int max_harmonic, double* du) {
    AbsQuad q; q.t.assign(gl_t, gl_t + n_gl); q.w.assign(gl_w, gl_w + n_gl);
    RayParams rp; rp.omega = 2.0 * M_PI * f; rp.mode = mode; rp.te_min = te_min; rp.max_harmonic = max_harmonic;
    grad_lambda(*(Plasma*)h, q, rp, u, du, nullptr);
}
void oracle_rhs_model(void* h, const double* gl_t, const double* gl_w, int n_gl, const double* u, double f, int mode,
                      double te_min, int max_harmonic, int absorption_model, double* du) {
    AbsQuad q; q.t.assign(gl_t, gl_t + n_gl); q.w.assign(gl_w, gl_w + n_gl);
    RayParams rp; rp.omega = 2.0 * M_PI * f; rp.mode = mode; rp.te_min = te_min; rp.max_harmonic = max_harmonic;
    rp.absorption_model = absorption_model;
    grad_lambda(*(Plasma*)h, q, rp, u, du, nullptr);
}
double oracle_abs_albajar(const double* gl_t, const double* gl_w, int n_gl, double omega, double X, double Y,
                          double N_abs, double N_par, double Te, int mode, double te_min, int max_harmonic) {
    AbsQuad q; q.t.assign(gl_t, gl_t + n_gl); q.w.assign(gl_w, gl_w + n_gl);
    return abs_Albajar_fast(q, omega, X, Y, N_abs, N_par, Te, mode, te_min, max_harmonic, nullptr);
}
double oracle_besselj(int n, double x) { return jn(n, x); }

// ---- warm-plasma absorption (reference src/general_absorption.jl), exposed piece by piece for pinning ----
double oracle_expei(double x) { return warm::expei(x); }
double oracle_gammln(double x) { return warm::gammln(x); }
double oracle_fact(int k) { return warm::fact(k); }
// out[l+2]: orders n..l+2
int oracle_ssbi(double zz, int n, int l, double* out) {
    std::vector<double> v = warm::ssbi(zz, n, l);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}
void oracle_zetac(double xi, double yi, double* re_im) {
    warm::cplx z = warm::zetac(xi, yi);
    re_im[0] = z.real(); re_im[1] = z.imag();
}
int oracle_larmornumber(double yg, double npl, double mu) { return warm::larmornumber(yg, npl, mu); }
// rr[(2 lrm + 1)][3][lrm + 1] (n = -lrm..lrm, k, m), row-major
void oracle_hermitian(double yg, double anpl, double amu, int lrm, int iwarm, double* rr) {
    warm::RR r = warm::hermitian(yg, anpl, amu, lrm, iwarm);
    std::memcpy(rr, r.v.data(), r.v.size() * sizeof(double));
}
// ri[lrm][3][lrm] (n = 1..lrm, k, m = 1..lrm), row-major
void oracle_antihermitian(double yg, double anpl, double amu, int lrm, double* ri) {
    warm::RI r = warm::antihermitian(yg, anpl, amu, lrm);
    std::memcpy(ri, r.v.data(), r.v.size() * sizeof(double));
}
// out[10]: Re/Im of N_perp, ex, ey, ez, then ierr, iterations
void oracle_warmdisp(double xg, double yg, double anpl, double amu, double anprc, int sox, int iwarm, int lrm, double* out) {
    warm::WarmDisp w = warm::warmdisp(xg, yg, anpl, amu, anprc, sox, iwarm, lrm);
    out[0] = w.anpr.real(); out[1] = w.anpr.imag(); out[2] = w.ex.real(); out[3] = w.ex.imag(); out[4] = w.ey.real();
    out[5] = w.ey.imag(); out[6] = w.ez.real(); out[7] = w.ez.imag(); out[8] = w.ierr; out[9] = w.iterations;
}
// α(omega, X, Y, N_r, theta, te, v_g_perp, imod) of reference src/general_absorption.jl:1328-1337.
// out[7]: N_warm, alpha, lrm, ierr, iterations, Re N_perp, Im N_perp
void oracle_warm_alpha(double omega, double X, double Y, double N_r, double theta, double te, double v_g_perp, int imod,
                       double* out) {
    warm::AlphaOut o = warm::alpha(omega, X, Y, N_r, theta, te, v_g_perp, imod);
    out[0] = o.N_warm; out[1] = o.alpha; out[2] = o.lrm; out[3] = o.ierr; out[4] = o.iterations;
    out[5] = o.N_perp.real(); out[6] = o.N_perp.imag();
}

void oracle_gausshermite(int n, double* x, double* w) {
    std::vector<double> xv, wv;
    gausshermite(n, xv, wv);
    std::memcpy(x, xv.data(), n * sizeof(double));
    std::memcpy(w, wv.data(), n * sizeof(double));
}
// returns n rays (or <0); call with pos==NULL to size
int oracle_launch(const double* x0, const double* N0, double w, double inv_Rc, double f, int N_rings, int min_az,
                  int normalize, double* pos, double* dir, double* weight) {
    Bundle b;
    int rc = launch_peripheral_rays(x0, N0, w, inv_Rc, f, N_rings, min_az, normalize != 0, b);
    if (rc != 0) return rc;
    if (pos) {
        std::memcpy(pos, b.pos.data(), b.pos.size() * sizeof(double));
        std::memcpy(dir, b.dir.data(), b.dir.size() * sizeof(double));
        std::memcpy(weight, b.weight.data(), b.weight.size() * sizeof(double));
    }
    return b.n;
}

// ray initialisation only: out = p_plasma[3], N_plasma[3], s0 ; returns status
int oracle_ray_init(void* h, const double* x0, const double* N0, double f, int mode, double* out) {
    Plasma* p = (Plasma*)h;
    double pp[3], Np[3];
    int st = first_point(*p, x0, N0, pp);
    if (st != RAY_OK) return st;
    if (!(p->psi_at(pp) <= p->psi_prof_max)) return RAY_INIT_FAILED;
    st = vacuum_plasma_refraction(*p, pp, N0, 2.0 * M_PI * f, mode, Np);
    if (st != RAY_OK) return st;
    for (int k = 0; k < 3; ++k) { out[k] = pp[k]; out[3 + k] = Np[k]; }
    out[6] = std::sqrt((pp[0] - x0[0]) * (pp[0] - x0[0]) + (pp[1] - x0[1]) * (pp[1] - x0[1]) + (pp[2] - x0[2]) * (pp[2] - x0[2]));
    return RAY_OK;
}

// ---- single ray with full trajectory (make_ray, reference src/solve.jl:135-181) ----
struct RayHandle {
    RayResult r;
    std::vector<double> dP_dV, dP_shell;
    double deposited = 0.0;
};
// depo_kind: 0 = faithful (reference algorithm), 1 = streaming (GPU algorithm)
void* oracle_make_ray(void* h, const double* gl_t, const double* gl_w, int n_gl, const oracle_options* o,
                      const double* x0, const double* N0, double f, int mode, double s_max, const double* psi_grid,
                      int npsi, int depo_kind) {
    Plasma* p = (Plasma*)h;
    AbsQuad q; q.t.assign(gl_t, gl_t + n_gl); q.w.assign(gl_w, gl_w + n_gl);
    RayHandle* rh = new RayHandle();
    make_ray(*p, q, to_opts(o), x0, N0, f, mode, s_max, rh->r, true);
    rh->dP_dV.assign(npsi, 0.0); rh->dP_shell.assign(npsi, 0.0);
    RayResult& r = rh->r;
    if (r.status == RAY_OK && npsi > 1) {
        int n = (int)r.s.size();
        if (depo_kind == 0)
            rh->deposited = deposition_faithful(*p, r.s.data(), r.x.data(), r.y.data(), r.z.data(), r.dP_ds.data(), n,
                                                psi_grid, npsi, rh->dP_dV.data(), rh->dP_shell.data());
        else
            rh->deposited = deposition_streaming(*p, r.s.data(), r.psi.data(), r.dpsi_ds.data(), r.P.data(), r.aP.data(), n,
                                                 psi_grid, npsi, rh->dP_dV.data(), rh->dP_shell.data());
    }
    return rh;
}
int oracle_ray_status(void* rh) { return ((RayHandle*)rh)->r.status; }
int oracle_ray_npoints(void* rh) { return (int)((RayHandle*)rh)->r.s.size(); }
double oracle_ray_deposited(void* rh) { return ((RayHandle*)rh)->deposited; }
// which: 0 s,1 x,2 y,3 z,4 P,5 dP_ds,6 psi,7 dpsi_ds,8 aP ; 9 dP_dV (npsi), 10 dP_shell (npsi), 11 u_final (7), 12 counters (5)
int oracle_ray_get(void* rhv, int which, double* out) {
    RayHandle* rh = (RayHandle*)rhv;
    const std::vector<double>* v = nullptr;
    switch (which) {
        case 0: v = &rh->r.s; break; case 1: v = &rh->r.x; break; case 2: v = &rh->r.y; break; case 3: v = &rh->r.z; break;
        case 4: v = &rh->r.P; break; case 5: v = &rh->r.dP_ds; break; case 6: v = &rh->r.psi; break;
        case 7: v = &rh->r.dpsi_ds; break; case 8: v = &rh->r.aP; break; case 9: v = &rh->dP_dV; break;
        case 10: v = &rh->dP_shell; break;
        case 11: std::memcpy(out, rh->r.u_final, 7 * sizeof(double)); return 7;
        case 12:
            out[0] = (double)rh->r.cnt.n_acc; out[1] = (double)rh->r.cnt.n_rej; out[2] = (double)rh->r.cnt.n_rhs;
            out[3] = (double)rh->r.cnt.abs.n_alpha; out[4] = (double)rh->r.cnt.abs.n_harm;
            return 5;
        default: return -1;
    }
    std::memcpy(out, v->data(), v->size() * sizeof(double));
    return (int)v->size();
}
void oracle_ray_free(void* rh) { delete (RayHandle*)rh; }

// ---- bundle (make_beam's parallel region + reduction, reference src/solve.jl:219-240), OpenMP over rays ----
// pos/dir: [3][n] component-major. freq/mode: per ray if per_ray_fm != 0 else length 1.
// counters[5]: n_acc, n_rej, n_rhs, n_alpha, n_harm (summed over rays)
int oracle_trace_bundle(void* h, const double* gl_t, const double* gl_w, int n_gl, const oracle_options* o, int64_t n_rays,
                        const double* pos, const double* dir, const double* weight, const double* freq,
                        const int32_t* mode, int per_ray_fm, double s_max, const double* psi_grid, int npsi,
                        int depo_kind, int n_threads, double* dP_dV, double* deposited_power, double* P_final,
                        int32_t* n_points, int32_t* status, double* counters, double* dP_dV_streaming /* optional */) {
    Plasma* p = (Plasma*)h;
    AbsQuad q; q.t.assign(gl_t, gl_t + n_gl); q.w.assign(gl_w, gl_w + n_gl);
    Options opt = to_opts(o);
    std::vector<double> prof(npsi, 0.0), prof2(npsi, 0.0);
    double dep = 0.0;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        std::vector<double> lprof(npsi, 0.0), lprof2(npsi, 0.0), rayprof(npsi);
        double ldep = 0.0, l0 = 0, l1 = 0, l2 = 0, l3 = 0, l4 = 0;
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < n_rays; ++i) {
            double x0[3] = {pos[i], pos[n_rays + i], pos[2 * n_rays + i]};
            double N0[3] = {dir[i], dir[n_rays + i], dir[2 * n_rays + i]};
            double f = per_ray_fm ? freq[i] : freq[0];
            int md = per_ray_fm ? mode[i] : mode[0];
            RayResult r;
            make_ray(*p, q, opt, x0, N0, f, md, s_max, r, true);
            status[i] = r.status;
            n_points[i] = (int32_t)r.s.size();
            P_final[i] = r.status == RAY_OK ? r.u_final[6] : 0.0;
            if (r.status == RAY_OK && npsi > 1) {
                int n = (int)r.s.size();
                double P;
                if (depo_kind == 0)
                    P = deposition_faithful(*p, r.s.data(), r.x.data(), r.y.data(), r.z.data(), r.dP_ds.data(), n, psi_grid,
                                            npsi, rayprof.data(), nullptr);
                else
                    P = deposition_streaming(*p, r.s.data(), r.psi.data(), r.dpsi_ds.data(), r.P.data(), r.aP.data(), n,
                                             psi_grid, npsi, rayprof.data(), nullptr);
                for (int j = 0; j < npsi; ++j) lprof[j] += rayprof[j] * weight[i];  // src/solve.jl:238
                ldep += P * weight[i];                                              // src/solve.jl:239
                if (dP_dV_streaming) {
                    deposition_streaming(*p, r.s.data(), r.psi.data(), r.dpsi_ds.data(), r.P.data(), r.aP.data(), n,
                                         psi_grid, npsi, rayprof.data(), nullptr);
                    for (int j = 0; j < npsi; ++j) lprof2[j] += rayprof[j] * weight[i];
                }
            }
            l0 += r.cnt.n_acc; l1 += r.cnt.n_rej; l2 += r.cnt.n_rhs; l3 += r.cnt.abs.n_alpha; l4 += r.cnt.abs.n_harm;
        }
#pragma omp critical
        {
            for (int j = 0; j < npsi; ++j) { prof[j] += lprof[j]; prof2[j] += lprof2[j]; }
            dep += ldep; c0 += l0; c1 += l1; c2 += l2; c3 += l3; c4 += l4;
        }
    }
    for (int j = 0; j < npsi; ++j) dP_dV[j] = prof[j];
    if (dP_dV_streaming) for (int j = 0; j < npsi; ++j) dP_dV_streaming[j] = prof2[j];
    *deposited_power = dep;
    if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3; counters[4] = c4; }
    return 0;
}

// ---- FITPACK restatement, exposed for pinning against scipy ----
void* oracle_fp_fit(const double* x, const double* y, int m) {
    InterpSpline* s = new InterpSpline();
    if (!s->fit(x, y, m)) { delete s; return nullptr; }
    return s;
}
void oracle_fp_free(void* s) { delete (InterpSpline*)s; }
int oracle_fp_tc(void* sv, double* t, double* c) {
    InterpSpline* s = (InterpSpline*)sv;
    std::memcpy(t, s->t.data(), s->t.size() * sizeof(double));
    std::memcpy(c, s->c.data(), s->c.size() * sizeof(double));
    return s->m;
}
double oracle_fp_value(void* s, double x) { return ((InterpSpline*)s)->value(x); }
double oracle_fp_integral(void* s, double a, double b) { return ((InterpSpline*)s)->integral(a, b); }
int oracle_fp_roots(void* s, double level, int maxn, double* out) {
    std::vector<double> r;
    ((InterpSpline*)s)->roots(level, maxn, r);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
}

// tableau of scheme (0 Tsit5, 1 OwrenZen3): a[7][7] row-major, btilde[7]; returns the number of stages
int oracle_tableau(int scheme, double* a, double* bt) {
    const Tableau& t = scheme == 1 ? tableau_owrenzen3() : tableau_tsit5();
    for (int i = 0; i < 7; ++i) { for (int j = 0; j < 7; ++j) a[i * 7 + j] = t.a[i][j]; bt[i] = t.btilde[i]; }
    return t.stages;
}

int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"

"""TEST INFRASTRUCTURE — ctypes wrapper of the CPU oracle (oracle/libtorj_oracle.so). Not part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY UNPINNED against the Julia reference (see oracle/api.cpp header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class OracleOptions(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("n_segments", C.c_int32), ("dtmax", C.c_double), ("abstol", C.c_double),
                ("reltol", C.c_double), ("psi_stop", C.c_double), ("p_stop", C.c_double), ("te_min", C.c_double),
                ("max_harmonic", C.c_int32), ("max_steps_per_segment", C.c_int32), ("absorption_model", C.c_int32),
                ("reserved_", C.c_int32), ("alpha_floor", C.c_double)]

    @classmethod
    def default(cls, **kw):
        o = cls(0, 100, 1e-4, 1e-6, 1e-6, 1.0, 1e-6, 20.0, 3, 100000, 0, 0, 0.0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libtorj_oracle.so")
    if force or not os.path.exists(path):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_plasma_create.restype = C.c_void_p
        L.oracle_plasma_create.argtypes = [c_dp, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp,
                                           c_dp, c_dp, C.c_int]
        L.oracle_plasma_destroy.argtypes = [C.c_void_p]
        L.oracle_plasma_coefs.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.oracle_volume_coefs.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]
        L.oracle_psi_prof_max.restype = C.c_double
        L.oracle_psi_prof_max.argtypes = [C.c_void_p]
        L.oracle_spline_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.oracle_volume_eval.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        L.oracle_eval_plasma.argtypes = [C.c_void_p, c_dp, c_dp, C.c_double, C.c_int, c_dp]
        L.oracle_rhs.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, c_dp, C.c_double, C.c_int, C.c_double, C.c_int, c_dp]
        L.oracle_rhs_model.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, c_dp, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, c_dp]
        for nm in ("oracle_expei", "oracle_gammln"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [C.c_double]
        L.oracle_fact.restype = C.c_double
        L.oracle_fact.argtypes = [C.c_int]
        L.oracle_ssbi.argtypes = [C.c_double, C.c_int, C.c_int, c_dp]
        L.oracle_zetac.argtypes = [C.c_double, C.c_double, c_dp]
        L.oracle_larmornumber.argtypes = [C.c_double, C.c_double, C.c_double]
        L.oracle_hermitian.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, c_dp]
        L.oracle_antihermitian.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, c_dp]
        L.oracle_warmdisp.argtypes = [C.c_double] * 5 + [C.c_int, C.c_int, C.c_int, c_dp]
        L.oracle_warm_alpha.argtypes = [C.c_double] * 7 + [C.c_int, c_dp]
        L.oracle_abs_albajar.restype = C.c_double
        L.oracle_abs_albajar.argtypes = [c_dp, c_dp, C.c_int] + [C.c_double] * 6 + [C.c_int, C.c_double, C.c_int]
        L.oracle_besselj.restype = C.c_double
        L.oracle_besselj.argtypes = [C.c_int, C.c_double]
        L.oracle_gausshermite.argtypes = [C.c_int, c_dp, c_dp]
        L.oracle_launch.argtypes = [c_dp, c_dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.oracle_ray_init.argtypes = [C.c_void_p, c_dp, c_dp, C.c_double, C.c_int, c_dp]
        L.oracle_make_ray.restype = C.c_void_p
        L.oracle_make_ray.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, C.POINTER(OracleOptions), c_dp, c_dp, C.c_double,
                                      C.c_int, C.c_double, c_dp, C.c_int, C.c_int]
        L.oracle_ray_status.argtypes = [C.c_void_p]
        L.oracle_ray_npoints.argtypes = [C.c_void_p]
        L.oracle_ray_deposited.restype = C.c_double
        L.oracle_ray_deposited.argtypes = [C.c_void_p]
        L.oracle_ray_get.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.oracle_ray_free.argtypes = [C.c_void_p]
        L.oracle_trace_bundle.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, C.POINTER(OracleOptions), C.c_int64, c_dp, c_dp,
                                          c_dp, c_dp, c_ip, C.c_int, C.c_double, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp,
                                          c_dp, c_ip, c_ip, c_dp, c_dp]
        L.oracle_fp_fit.restype = C.c_void_p
        L.oracle_fp_fit.argtypes = [c_dp, c_dp, C.c_int]
        L.oracle_fp_free.argtypes = [C.c_void_p]
        L.oracle_fp_tc.argtypes = [C.c_void_p, c_dp, c_dp]
        L.oracle_fp_value.restype = C.c_double
        L.oracle_fp_value.argtypes = [C.c_void_p, C.c_double]
        L.oracle_fp_integral.restype = C.c_double
        L.oracle_fp_integral.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oracle_fp_roots.argtypes = [C.c_void_p, C.c_double, C.c_int, c_dp]
        L.oracle_tableau.argtypes = [C.c_int, c_dp, c_dp]
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _f2(a):
    """[nR, nZ] array -> R-fastest buffer (Julia column-major)."""
    return _d(np.asarray(a, dtype=np.float64).T)


FIELDS = {"psi": 0, "lnne": 1, "lnTe": 2, "BR": 3, "BZ": 4, "Bphi": 5}


class OraclePlasma:
    """Oracle-side Plasma(...) constructor, same positional arguments as reference src/plasma.jl:30-32."""

    def __init__(self, R_coords, Z_coords, psi_norm_data, psi_prof, ne_prof, Te_prof, Br_data, Bz_data, Bphi_data,
                 eqt1d_psi_norm, eqt1d_volume):
        L = lib()
        self.nR, self.nZ = len(R_coords), len(Z_coords)
        keep = [_d(R_coords), _d(Z_coords), _f2(psi_norm_data), _d(psi_prof), _d(ne_prof), _d(Te_prof), _f2(Br_data),
                _f2(Bz_data), _f2(Bphi_data), _d(eqt1d_psi_norm), _d(eqt1d_volume)]
        p = [k[1] for k in keep]
        self.h = L.oracle_plasma_create(p[0], self.nR, p[1], self.nZ, p[2], p[3], p[4], p[5], len(psi_prof), p[6], p[7],
                                        p[8], p[9], p[10], len(eqt1d_psi_norm))
        self.psi_prof_max = L.oracle_psi_prof_max(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().oracle_plasma_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def coefs(self, field: str) -> np.ndarray:
        """(nZ+2, nR+2) C-order array == (nR+2)x(nZ+2) with R fastest."""
        out = np.empty((self.nZ + 2, self.nR + 2))
        lib().oracle_plasma_coefs(self.h, FIELDS[field], out.ctypes.data_as(c_dp))
        return out

    def volume_coefs(self):
        out = np.empty(4096)
        x0, dx = C.c_double(), C.c_double()
        n = lib().oracle_volume_coefs(self.h, out.ctypes.data_as(c_dp), C.byref(x0), C.byref(dx))
        return out[: n + 2].copy(), x0.value, dx.value

    def spline(self, field: str, R, Z):
        R, pR = _d(np.atleast_1d(R)); Z, pZ = _d(np.atleast_1d(Z))
        v = np.empty_like(R); dR = np.empty_like(R); dZ = np.empty_like(R)
        lib().oracle_spline_eval(self.h, FIELDS[field], len(R), pR, pZ, v.ctypes.data_as(c_dp), dR.ctypes.data_as(c_dp),
                                 dZ.ctypes.data_as(c_dp))
        return v, dR, dZ

    def volume(self, psi):
        psi, pp = _d(np.atleast_1d(psi))
        out = np.empty_like(psi)
        lib().oracle_volume_eval(self.h, len(psi), pp, out.ctypes.data_as(c_dp))
        return out

    def eval_plasma(self, x, N, omega, mode=1):
        x, px = _d(x); N, pN = _d(N)
        out = np.empty(8)
        lib().oracle_eval_plasma(self.h, px, pN, omega, mode, out.ctypes.data_as(c_dp))
        return dict(X=out[0], Y=out[1], N_par=out[2], b=out[3:6].copy(), Te=out[6], Lambda=out[7])

    def rhs(self, u, f, mode, gl, te_min=20.0, max_harmonic=3, absorption_model=0):
        u, pu = _d(u); t, pt = _d(gl[0]); w, pw = _d(gl[1])
        du = np.empty(7)
        lib().oracle_rhs_model(self.h, pt, pw, len(t), pu, f, mode, te_min, max_harmonic, absorption_model,
                               du.ctypes.data_as(c_dp))
        return du

    def ray_init(self, x0, N0, f, mode):
        x0, p0 = _d(x0); N0, pN = _d(N0)
        out = np.zeros(7)
        st = lib().oracle_ray_init(self.h, p0, pN, f, mode, out.ctypes.data_as(c_dp))
        return st, out

    def make_ray(self, x0, N0, f, mode, s_max, psi_grid, gl, opts=None, deposition="faithful"):
        """Oracle make_ray (reference src/solve.jl:135-181). Returns a dict of arrays."""
        L = lib()
        opts = opts or OracleOptions.default()
        x0, p0 = _d(x0); N0, pN = _d(N0); t, pt = _d(gl[0]); w, pw = _d(gl[1]); g, pg = _d(psi_grid)
        rh = L.oracle_make_ray(self.h, pt, pw, len(t), C.byref(opts), p0, pN, f, mode, s_max, pg, len(g),
                               0 if deposition == "faithful" else 1)
        try:
            n = L.oracle_ray_npoints(rh)
            res = dict(status=L.oracle_ray_status(rh), deposited_power=L.oracle_ray_deposited(rh))
            names = ["s", "x", "y", "z", "P", "dP_ds", "psi", "dpsi_ds", "aP"]
            for i, nm in enumerate(names):
                a = np.empty(max(n, 1))
                k = L.oracle_ray_get(rh, i, a.ctypes.data_as(c_dp))
                res[nm] = a[:k].copy()
            for i, nm, ln in ((9, "dP_dV", len(g)), (10, "dP_shell", len(g)), (11, "u_final", 7), (12, "counters", 5)):
                a = np.empty(ln)
                L.oracle_ray_get(rh, i, a.ctypes.data_as(c_dp))
                res[nm] = a
            return res
        finally:
            L.oracle_ray_free(rh)

    def trace_bundle(self, pos, dir, weight, freq, mode, s_max, psi_grid, gl, opts=None, deposition="faithful",
                     n_threads=0, also_streaming=False):
        """pos/dir: [N,3] arrays (as launch_peripheral_rays returns). freq/mode scalar or per-ray.
        also_streaming: additionally return the streaming-algorithm profile (key dP_dV_streaming) from the same rays."""
        L = lib()
        opts = opts or OracleOptions.default()
        pos = np.asarray(pos, dtype=np.float64); dir = np.asarray(dir, dtype=np.float64)
        n = pos.shape[0]
        posT, pp = _d(pos.T); dirT, pd = _d(dir.T); wt, pw = _d(weight)
        per_ray = int(np.ndim(freq) > 0)
        fr, pf = _d(np.atleast_1d(freq))
        md = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(mode), fr.shape), dtype=np.int32)  # per-ray f with one mode
        t, pt = _d(gl[0]); w, pgw = _d(gl[1]); g, pg = _d(psi_grid)
        prof = np.zeros(len(g)); dep = C.c_double()
        Pf = np.zeros(n); npts = np.zeros(n, dtype=np.int32); st = np.zeros(n, dtype=np.int32); cnt = np.zeros(5)
        prof2 = np.zeros(len(g)) if also_streaming else None
        L.oracle_trace_bundle(self.h, pt, pgw, len(t), C.byref(opts), n, pp, pd, pw, pf, md.ctypes.data_as(c_ip), per_ray,
                              s_max, pg, len(g), 0 if deposition == "faithful" else 1, n_threads, prof.ctypes.data_as(c_dp),
                              C.byref(dep), Pf.ctypes.data_as(c_dp), npts.ctypes.data_as(c_ip), st.ctypes.data_as(c_ip),
                              cnt.ctypes.data_as(c_dp), prof2.ctypes.data_as(c_dp) if also_streaming else None)
        return dict(dP_dV_streaming=prof2, dP_dV=prof, deposited_power=dep.value, P_final=Pf, n_points=npts, status=st,
                    counters=dict(n_acc=cnt[0], n_rej=cnt[1], n_rhs=cnt[2], n_alpha=cnt[3], n_harm=cnt[4]))


def abs_albajar(omega, X, Y, N_abs, N_par, Te, mode, gl, te_min=20.0, max_harmonic=3):
    t, pt = _d(gl[0]); w, pw = _d(gl[1])
    return lib().oracle_abs_albajar(pt, pw, len(t), omega, X, Y, N_abs, N_par, Te, mode, te_min, max_harmonic)


# ---- warm-plasma absorption: reference src/general_absorption.jl ----
def expei(x):
    return lib().oracle_expei(float(x))


def gammln(x):
    return lib().oracle_gammln(float(x))


def fact(k):
    return lib().oracle_fact(int(k))


def ssbi(zz, n, l):
    out = np.zeros(l + 2)
    k = lib().oracle_ssbi(float(zz), int(n), int(l), out.ctypes.data_as(c_dp))
    return out[:k]


def zetac(xi, yi):
    out = np.zeros(2)
    lib().oracle_zetac(float(xi), float(yi), out.ctypes.data_as(c_dp))
    return complex(out[0], out[1])


def larmornumber(yg, npl, mu):
    return lib().oracle_larmornumber(float(yg), float(npl), float(mu))


def hermitian(yg, anpl, amu, lrm, iwarm=3):
    """rr[n + lrm, k, m] for n in -lrm..lrm, k in 0..2, m in 0..lrm."""
    rr = np.zeros((2 * lrm + 1, 3, lrm + 1))
    lib().oracle_hermitian(float(yg), float(anpl), float(amu), int(lrm), int(iwarm), rr.ctypes.data_as(c_dp))
    return rr


def antihermitian(yg, anpl, amu, lrm):
    """ri[n - 1, k, m - 1] for n in 1..lrm, k in 0..2, m in 1..lrm."""
    ri = np.zeros((lrm, 3, lrm))
    lib().oracle_antihermitian(float(yg), float(anpl), float(amu), int(lrm), ri.ctypes.data_as(c_dp))
    return ri


def warmdisp(xg, yg, anpl, amu, anprc, sox, iwarm, lrm):
    o = np.zeros(10)
    lib().oracle_warmdisp(float(xg), float(yg), float(anpl), float(amu), float(anprc), int(sox), int(iwarm), int(lrm),
                          o.ctypes.data_as(c_dp))
    return dict(anpr=complex(o[0], o[1]), ex=complex(o[2], o[3]), ey=complex(o[4], o[5]), ez=complex(o[6], o[7]),
                ierr=int(o[8]), iterations=int(o[9]))


def warm_alpha(omega, X, Y, N_r, theta, te, v_g_perp, imod):
    """α(omega, X, Y, N_r, theta, te, v_g_perp, imod) -> dict(N_warm, alpha, lrm, ierr, iterations, N_perp)."""
    o = np.zeros(7)
    lib().oracle_warm_alpha(float(omega), float(X), float(Y), float(N_r), float(theta), float(te), float(v_g_perp),
                            int(imod), o.ctypes.data_as(c_dp))
    return dict(N_warm=o[0], alpha=o[1], lrm=int(o[2]), ierr=int(o[3]), iterations=int(o[4]), N_perp=complex(o[5], o[6]))


def tableau(scheme):
    """(a[S,S], btilde[S]) of the oracle's RK scheme (0 Tsit5, 1 OwrenZen3); last row of a is b (FSAL)."""
    a = np.zeros(49); bt = np.zeros(7)
    S = lib().oracle_tableau(scheme, a.ctypes.data_as(c_dp), bt.ctypes.data_as(c_dp))
    return a.reshape(7, 7)[:S, :S].copy(), bt[:S].copy()


def max_threads():
    return lib().oracle_max_threads()


def besselj(n, x):
    return lib().oracle_besselj(int(n), float(x))


def gausshermite(n):
    x = np.empty(n); w = np.empty(n)
    lib().oracle_gausshermite(n, x.ctypes.data_as(c_dp), w.ctypes.data_as(c_dp))
    return x, w


def launch_peripheral_rays(x0, N0, w, inverse_curvature_radius, f, N_rings=3, min_azimuthal_points=5,
                           normalize_weight_sum=True):
    """Oracle restatement of reference src/launch.jl:24-132. Returns ([N,3], [N,3], [N])."""
    L = lib()
    x0, p0 = _d(x0); N0, pN = _d(N0)
    n = L.oracle_launch(p0, pN, w, inverse_curvature_radius, f, N_rings, min_azimuthal_points, int(normalize_weight_sum),
                        None, None, None)
    if n < 0:
        raise ValueError(f"N_rings = {N_rings} < 2 which is the minimum")
    pos = np.empty((3, n)); dr = np.empty((3, n)); wt = np.empty(n)
    L.oracle_launch(p0, pN, w, inverse_curvature_radius, f, N_rings, min_azimuthal_points, int(normalize_weight_sum),
                    pos.ctypes.data_as(c_dp), dr.ctypes.data_as(c_dp), wt.ctypes.data_as(c_dp))
    return pos.T.copy(), dr.T.copy(), wt


class FitpackSpline:
    """Oracle restatement of Dierckx.Spline1D(x, y, k=3) / roots / integrate (reference src/plasma.jl:101-134)."""

    def __init__(self, x, y):
        x, px = _d(x); y, py = _d(y)
        self.m = len(x)
        self.h = lib().oracle_fp_fit(px, py, self.m)
        if not self.h:
            raise ValueError("fit failed (x not strictly increasing or m<4)")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().oracle_fp_free(self.h); self.h = None
        except Exception:
            pass

    def tc(self):
        t = np.empty(self.m + 4); c = np.empty(self.m)
        lib().oracle_fp_tc(self.h, t.ctypes.data_as(c_dp), c.ctypes.data_as(c_dp))
        return t, c

    def __call__(self, x):
        return np.array([lib().oracle_fp_value(self.h, float(v)) for v in np.atleast_1d(x)])

    def integral(self, a, b):
        return lib().oracle_fp_integral(self.h, float(a), float(b))

    def roots(self, level=0.0, maxn=8):
        out = np.empty(maxn + 1)
        n = lib().oracle_fp_roots(self.h, float(level), maxn, out.ctypes.data_as(c_dp))
        return out[:n].copy()

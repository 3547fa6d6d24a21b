"""abs_Al_init and the absorption coefficients — mirror of reference src/absorption.jl:1-7,191-235 and of the
warm-plasma α of src/general_absorption.jl:1328-1337."""
from __future__ import annotations

import numpy as np

from . import _lib

# module-level state like the reference's globals _int_absz / _int_weights (src/constants.jl:7-8)
_int_absz = np.zeros(0)
_int_weights = np.zeros(0)


def abs_Al_init(N_absz: int, ctx=None):
    """Gauss-Legendre nodes/weights on [-1,1] (FastGaussQuadrature.gausslegendre) -> device constant memory."""
    global _int_absz, _int_weights
    _int_absz, _int_weights = np.polynomial.legendre.leggauss(int(N_absz))
    ctx = ctx or _lib.context()
    _lib.check(_lib.lib().torj_abs_init(ctx, int(N_absz), _int_absz.ctypes.data_as(_lib.c_dp),
                                        _int_weights.ctypes.data_as(_lib.c_dp)))


def alpha_approx(x, N, plasma, omega, mode, ctx=None):
    """α_approx(x, N, plasma, omega, mode) — reference src/absorption.jl:228-235 (x, N: [n,3] or [3])."""
    single = np.ndim(x) == 1
    out = plasma.probe(x, N, omega / (2.0 * np.pi), mode, ctx)["alpha"]
    return float(out[0]) if single else out


def warm_alpha(omega, X, Y, N_r, theta, te, v_g_perp, imod, ctx=None):
    """α(omega, X, Y, N_r, theta, te, v_g_perp, imod) — reference src/general_absorption.jl:1328-1337, evaluated on the
    device (scalars or arrays, broadcast). Returns (N_warm, alpha) like the reference; `warm_alpha.last` holds the Larmor
    order lrm and the error flag (99: negative solution, :1229-1232; 98: lrm < 1, where the reference throws)."""
    ctx = ctx or _lib.context()
    single = all(np.ndim(a) == 0 for a in (omega, X, Y, N_r, theta, te, v_g_perp))
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in np.broadcast_arrays(*np.atleast_1d(omega, X, Y, N_r, theta, te, v_g_perp))]
    n = len(arrs[0])
    Nw = np.zeros(n); al = np.zeros(n); lrm = np.zeros(n, dtype=np.int32); ierr = np.zeros(n, dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(_lib.c_dp)
    _lib.check(_lib.lib().torj_warm_alpha(ctx, n, *[dp(a) for a in arrs], int(imod), dp(Nw), dp(al),
                                          lrm.ctypes.data_as(_lib.c_ip), ierr.ctypes.data_as(_lib.c_ip)))
    warm_alpha.last = dict(lrm=lrm, ierr=ierr)
    return (float(Nw[0]), float(al[0])) if single else (Nw, al)

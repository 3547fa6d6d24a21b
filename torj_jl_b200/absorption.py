"""abs_Al_init and the absorption coefficient — mirror of reference src/absorption.jl:1-7,191-235."""
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

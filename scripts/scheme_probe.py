"""k_trace time of the 65 543-ray bench bundle for both integrator schemes (0 = Tsit5, 1 = OwrenZen3)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
ctx = _lib.context(0)
tj.abs_Al_init(24, ctx)
pl = tj.Plasma(*tj.solovev_arrays().values())
x0 = np.array([2.5, 0, 0.4]); N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30), 0.0)
pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, 95e9, N_rings=66, min_azimuthal_points=14)
psi = np.linspace(0, 1, 1000)
ms = C.c_double()
for scheme in (0, 1):
    for i in range(3):
        r = tj.trace_bundle(pl, pos, dirs, w, 95e9, 1, 1.0, psi, ctx=ctx, options=tj.default_options(scheme=scheme))
        _lib.check(tj.lib().torj_ctx_last_trace_ms(ctx, C.byref(ms)))
    n = r["counters"]["n_acc"]
    print(f"scheme {scheme}: k_trace {ms.value:7.1f} ms  steps {n}  {n / ms.value * 1e3:.4g} ray-steps/s  absorbed {r['deposited_power']:.12f}")

"""Would grouping rays that reach the absorbing layer at the same arc length into the same warps pay?
Experiment: trace the bench bundle, sort the rays by their number of accepted steps (a perfect proxy for the arc length
at which they are absorbed), trace the permuted bundle, compare the k_trace times."""
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
def run(p, d, ww, label):
    for i in range(3):
        r = tj.trace_bundle(pl, p, d, ww, 95e9, 1, 1.0, psi, ctx=ctx)
        _lib.check(tj.lib().torj_ctx_last_trace_ms(ctx, C.byref(ms)))
    print(f"{label}: k_trace {ms.value:7.2f} ms  steps {r['counters']['n_acc']}  absorbed {r['deposited_power']:.12f}")
    return r
r = run(pos, dirs, w, "launcher order")
npts = r["n_points"]
print("n_points min/median/max", npts.min(), int(np.median(npts)), npts.max())
o = np.argsort(npts, kind="stable")
run(pos[o], dirs[o], w[o], "sorted by steps ")
rng = np.random.default_rng(0); o = rng.permutation(len(w))
run(pos[o], dirs[o], w[o], "random order    ")

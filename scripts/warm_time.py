"""Warm-model throughput on one 1 025-ray beam (T_e0 = 10 keV, 110 GHz, a warp per ray) and on a 37 888-ray bundle (a thread
per ray): device time of the trace kernel, ray-steps/s and the algorithmic FP64 rate."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
import bench

L = tj.lib(); ctx = _lib.context(); tj.abs_Al_init(24)
PSI = np.linspace(0.0, 1.0, 1000)
plw = tj.Plasma(*bench.hot_arrays(10e3).values())
x0 = np.array([2.5, 0.0, 0.4]); N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.0)
out = {}
for name, nr, ma, lanes, smax in (("warm1025_lanes32", 7, 20, 32, 1.0), ("warm9k_lanes32", 24, 14, 32, 1.0), ("warm38k_lanes1", 50, 14, 1, 0.45)):
    p, d, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, 110e9, N_rings=nr, min_azimuthal_points=ma)
    opt = tj.default_options(absorption_model=1, lanes_per_ray=lanes)
    r = tj.trace_bundle(plw, p, d, w, 110e9, 1, smax, PSI, options=opt)
    t = C.c_double(); _lib.check(L.torj_ctx_last_trace_ms(ctx, C.byref(t)))
    c = r["counters"]
    out[name] = dict(rays=len(w), ms=t.value, steps=c["n_acc"], n_harm=c["n_harm"], gated=c["n_harm_pruned"], dep=r["deposited_power"],
                     ray_steps_per_s=c["n_acc"] / (t.value * 1e-3), tflops=bench.algorithmic_flops(c, 1) / (t.value * 1e-3) / 1e12)
    print(name, json.dumps(out[name]), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

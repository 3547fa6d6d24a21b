"""Where does the end-to-end time of torj_trace go? (diagnostic)"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
use_torch = len(sys.argv) > 1 and sys.argv[1] == "torch"
if use_torch:
    import torch
    torch.cuda.set_device(0)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    ctx = _lib.context(0, s.cuda_stream)
else:
    ctx = _lib.context(0)
tj.abs_Al_init(24, ctx)
pl = tj.Plasma(*tj.solovev_arrays().values())
x0 = np.array([2.5, 0, 0.4]); N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30), 0.0)
pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, 95e9, N_rings=66, min_azimuthal_points=14)
psi = np.linspace(0, 1, 1000)
ms = C.c_double()
for i in range(6):
    t0 = time.perf_counter()
    r = tj.trace_bundle(pl, pos, dirs, w, 95e9, 1, 1.0, psi, ctx=ctx)
    t1 = time.perf_counter()
    _lib.check(tj.lib().torj_ctx_last_trace_ms(ctx, C.byref(ms)))
    print(f"call {i}: e2e {1e3*(t1-t0):8.1f} ms   k_trace {ms.value:8.1f} ms  steps {r['counters']['n_acc']}")

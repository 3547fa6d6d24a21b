"""Lanes per ray against bundle size (device time of the trace, best of 2 after a warm-up):
cold 4 keV beam of 1 025 rays, and the hot (10 keV, 110/170 GHz, 2 launchers) scan at 1 025 .. 16 400 rays.
  python scripts/lpr_probe.py out.json"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
import bench

L = tj.lib()
ctx = _lib.context()
tj.abs_Al_init(24)
PSI = np.linspace(0.0, 1.0, 1000)
out = {}


def run(pl, pos, dirs, w, f, reps=2, beam_id=None, n_beams=1, **kw):
    opt = tj.default_options(**kw)
    ms = []
    for _ in range(reps + 1):
        r = tj.trace_bundle(pl, pos, dirs, w, f, 1, 1.0, PSI, options=opt, beam_id=beam_id, n_beams=n_beams)
        t = C.c_double()
        _lib.check(L.torj_ctx_last_trace_ms(ctx, C.byref(t)))
        ms.append(t.value)
    return min(ms[1:])


pl = tj.Plasma(*tj.solovev_arrays(257, 257).values())
plh = tj.Plasma(*tj.solovev_arrays(257, 257, Te0=10e3).values())
ps, ds, ws = bench.beam_bundle("small")
P, D, W, F, B = [], [], [], [], []
b = 0
for tor in (0.1, 0.0, 0.2, 0.3):
    for f in (110e9, 170e9):
        for z0, pol in ((0.4, 30.0), (-0.4, -30.0)):
            p, d, ww = tj.launch_peripheral_rays(np.array([2.5, 0.0, z0]), tj.pol_tor_angles_2_vector(np.deg2rad(pol), tor), 0.0174,
                                                 1 / 3.99, f, N_rings=7, min_azimuthal_points=20)
            P.append(p); D.append(d); W.append(ww); F.append(np.full(len(ww), f)); B.append(np.full(len(ww), b, dtype=np.int32)); b += 1
P, D, W, F, B = map(np.concatenate, (P, D, W, F, B))
lprs = [int(x) for x in os.environ.get("LPRS", "1,2,4,8").split(",")]
for lanes in lprs:
    row = {"cold1025": run(pl, ps, ds, ws, 95e9, lanes_per_ray=lanes)}
    for nb in (1, 2, 4, 8, 16):
        n = nb * 1025
        row[f"hot{n}"] = run(plh, P[:n], D[:n], W[:n], F[:n], reps=1, lanes_per_ray=lanes, beam_id=B[:n], n_beams=nb)
    out[f"lanes{lanes}"] = row
    print(lanes, json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

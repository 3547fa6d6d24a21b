"""Small end-to-end case for compute-sanitizer / quick checks: 46-ray beam, short path, trajectory window, probes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
tj.abs_Al_init(24)
pl = tj.Plasma(*tj.solovev_arrays(65, 65).values())
x0 = np.array([2.5, 0.0, 0.4]); N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.1)
pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, 95e9)
psi = np.linspace(0, 1, 100)
r = tj.trace_bundle(pl, pos, dirs, w, 95e9, 1, 0.45, psi, trajectories=(3, 5))
print("status", np.unique(r["status"]), "dep", r["deposited_power"], "steps", r["counters"]["n_acc"])
s, u, P, prof, dep = tj.make_ray(pl, x0, N0, 95e9, -1, 0.1, psi)
print("make_ray pts", len(s), "P_end", P[-1])
print("probe", pl.probe(pos[:4], dirs[:4], 95e9)["psi"])

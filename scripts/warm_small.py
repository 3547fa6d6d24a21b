"""A short warm-model trace for profiling: 1 184 rays (one per resident warp), T_e0 = 10 keV, 110 GHz, s_max = 6 cm past the
entry into the absorbing layer (launched from inside the plasma edge so that every step evaluates the quadrature)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
import bench

tj.abs_Al_init(24)
pl = tj.Plasma(*bench.hot_arrays(10e3).values())
p, d, w = tj.launch_peripheral_rays(np.array([2.5, 0.0, 0.4]), tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.0), 0.0174, 1 / 3.99,
                                    95e9, N_rings=8, min_azimuthal_points=17)
n = min(len(w), 1184)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for _ in range(2):
    r = tj.trace_bundle(pl, p[:n], d[:n], w[:n], 95e9, 1, 0.06, np.linspace(0, 1, 1000),
                        options=tj.default_options(absorption_model=1, lanes_per_ray=lanes, n_segments=6))
print(n, "rays", r["counters"], "absorbed", r["deposited_power"])

"""Exploratory GPU check (not a test): parity of probes / RHS / rays against the CPU oracle + first timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
from oracle import torj_oracle as O
import ctypes as C

arr = tj.solovev_arrays()
pl = tj.Plasma(*arr.values())
opl = O.OraclePlasma(*arr.values())
tj.abs_Al_init(24)
gl = np.polynomial.legendre.leggauss(24)
ctx = _lib.context()
tf, ms = C.c_double(), C.c_double()
_lib.check(_lib.lib().torj_fp64_peak(ctx, 20000, C.byref(tf), C.byref(ms)))
print(f"FP64 DFMA peak: {tf.value:.2f} TFLOP/s ({ms.value:.2f} ms)")

f = 95e9
N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30), 0.0)
x0 = np.array([2.5, 0.0, 0.4])
psi_grid = np.linspace(0, 1, 1000)
# oracle ray
t = time.time(); ro = opl.make_ray(x0, N0, f, 1, 0.4, psi_grid, gl); t_or = time.time() - t
# probes along the oracle ray
idx = np.linspace(2, len(ro['s']) - 1, 64).astype(int)
X = np.stack([ro['x'][idx], ro['y'][idx], ro['z'][idx]], 1)
Nn = np.tile(N0 * 0.9, (len(idx), 1))
pr = pl.probe(X, Nn, f, 1)
for i in (0, 30, 63):
    e = opl.eval_plasma(X[i], Nn[i], 2 * np.pi * f)
    print("probe", i, "dX", pr['X'][i] - e['X'], "dY", pr['Y'][i] - e['Y'], "dNpar", pr['N_par'][i] - e['N_par'], "dTe", pr['Te'][i] - e['Te'], "dL", pr['Lambda'][i] - e['Lambda'])
# rhs parity at points on the ray with consistent N (integrate oracle RHS states)
st, init = opl.ray_init(x0, N0, f, 1)
u = np.concatenate([init[:6], [1.0]])
U = []
for k in range(40):
    U.append(u.copy())
    du = opl.rhs(u, f, 1, gl)
    u = u + 0.009 * du
U = np.array(U)
dg = pl.rhs(U, f, 1)
do = np.array([opl.rhs(uu, f, 1, gl) for uu in U])
print("rhs max abs diff per comp", np.abs(dg - do).max(0))
print("rhs alpha*P oracle", do[:, 6][[0, 10, 20, 30, 39]], "gpu", dg[:, 6][[0, 10, 20, 30, 39]])
# single ray
t = time.time(); s, uu, P, prof, dep = tj.make_ray(pl, x0, N0, f, 1, 0.4, psi_grid); t_g = time.time() - t
print("make_ray: gpu pts", len(s), "oracle pts", len(ro['s']), "t_gpu", t_g, "t_oracle", t_or)
n = min(len(s), len(ro['s']))
xyz = np.array(uu)
print("max |ds|", np.abs(s[:n] - ro['s'][:n]).max(), "max |dx|", np.abs(xyz[:n, 0] - ro['x'][:n]).max(), np.abs(xyz[:n, 2] - ro['z'][:n]).max(),
      "max |dP|", np.abs(P[:n] - ro['P'][:n]).max(), "P_end", P[-1], ro['P'][-1])
print("dep", dep, ro['deposited_power'], "prof L2 rel vs faithful", np.linalg.norm(prof - ro['dP_dV']) / np.linalg.norm(ro['dP_dV']))
# beam 46 rays
t = time.time(); res = tj.make_beam(pl, 2.5, 0.0, 0.4, 0.0, np.deg2rad(30), 0.0174, 1 / 3.99, f, 1, 1.0, psi_grid); t_g = time.time() - t
pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, f)
t = time.time(); bo = opl.trace_bundle(pos, dirs, w, f, 1, 1.0, psi_grid, gl); t_or = time.time() - t
print("beam46: dep gpu", res[4], "oracle", bo['deposited_power'], "rel", abs(res[4] - bo['deposited_power']) / bo['deposited_power'], "t_gpu", t_g, "t_or", t_or)
print("beam46 prof L2 rel", np.linalg.norm(res[3] - bo['dP_dV']) / np.linalg.norm(bo['dP_dV']))
# timings
for nr, ma in ((7, 20), (66, 14)):
    pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, f, N_rings=nr, min_azimuthal_points=ma)
    for rep in range(2):
        t = time.time(); r = tj.trace_bundle(pl, pos, dirs, w, f, 1, 1.0, psi_grid); dt = time.time() - t
        c = r['counters']
        print(f"n={len(w)} t={dt:.3f}s steps={c['n_acc']} rej={c['n_rej']} rhs={c['n_rhs']} harm={c['n_harm']} ok={c['n_rays_ok']} "
              f"ray-steps/s={c['n_acc']/dt:.3e} rays/s={len(w)/dt:.1f} dep={r['deposited_power']:.9f} status!=0: {(r['status']!=0).sum()}")

"""BASELINE.json configs[4] as a scan: 2 launchers (z = +-0.4 m) x {110, 170} GHz x T_e0 in {10, 15, 25} keV, 1 025-ray
beams, one deposition profile per beam, one device call per equilibrium.  Absorption model: Albajar (the reference's
warm-plasma file is never included by its module).  Usage: python scripts/config5_scan.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj

tj.abs_Al_init(24)
psi = np.linspace(0, 1, 1000)
print("Te0[keV]  f[GHz]  z0[m]  absorbed  psi_N of peak dP/dV   rays  ray-steps")
for Te0 in (10e3, 15e3, 25e3):
    pl = tj.Plasma(*tj.solovev_arrays(257, 257, Te0=Te0).values())
    Ls = [dict(r=2.5, phi=0.0, z=z0, steering_angle_pol=np.deg2rad(pol), steering_angle_tor=0.1, spot_size=0.0174,
               inverse_curvature_radius=1 / 3.99, f=f, mode=1)
          for f in (110e9, 170e9) for z0, pol in ((0.4, 30.0), (-0.4, -30.0))]
    t0 = time.perf_counter()
    dP, dep, W, Pf, res = tj.make_beams(pl, Ls, 1.0, psi, N_rings=7, min_azimuthal_points=20)
    dt = time.perf_counter() - t0
    for b, L in enumerate(Ls):
        print(f"{Te0/1e3:7.0f} {L['f']/1e9:7.0f} {L['z']:6.1f}  {dep[b]:.6f}  {psi[np.argmax(dP[b])]:.3f}  {len(W[b]):5d}")
    print(f"   -> {len(res['status'])} rays, {res['counters']['n_acc']} steps in {dt*1e3:.0f} ms "
          f"({res['counters']['n_acc']/dt:.3e} ray-steps/s end to end)")

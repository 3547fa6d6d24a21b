"""Generates the committed fixtures in tests/golden/.  Run from the repo root: python tests/golden/make_golden.py

Sources (none of them is the repo's own CUDA path):
  fitpack_case.npz  scipy.interpolate.splrep/sproot/splint = the FITPACK Fortran behind Dierckx.jl
                    (reference src/plasma.jl:101-134)
  launch_known.npz  numpy hermgauss restatement of reference src/launch.jl:72-83,120,128: ray counts and the
                    un-normalised weight sum asserted by reference test/tests/test_launch_weights.jl:50
  bessel_gl.npz     scipy.special.jv (= SpecialFunctions.besselj), numpy leggauss (= FastGaussQuadrature.gausslegendre)
  oracle_ray.npz    REGRESSION fixture produced by the oracle itself (config-1 stand-in ray, every 50th sample);
                    the Julia reference cannot run here, so this pins the oracle against drift, not against Julia.
  warm_alpha.npz    REGRESSION fixture of the oracle's restatement of α (reference src/general_absorption.jl:1328-1337)
                    on a deterministic grid of (X, Y, N, theta, Te, mode); same caveat: drift only, not Julia. The special
                    functions underneath are pinned against scipy in tests/test_oracle_warm.py.
"""
import os
import sys

import numpy as np
from scipy import interpolate, special

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def fitpack_case():
    rng = np.random.default_rng(20261018)
    x = np.cumsum(rng.uniform(0.2, 1.0, 120))
    y = np.cos(x / 4.0) * np.exp(-x / 60.0) + 0.05 * np.sin(1.7 * x)
    tck = interpolate.splrep(x, y, s=0, k=3)
    levels = np.array([-0.3, 0.0, 0.25, 0.6])
    roots = [interpolate.sproot((tck[0], tck[1] - lv * (np.arange(len(tck[1])) < len(x)), 3), mest=64) for lv in levels]
    pairs = np.array([[x[0], x[-1]], [x[5] + 0.1, x[77] - 0.05], [x[0] - 3.0, x[-1] + 2.0], [x[40], x[41]]])
    ints = np.array([interpolate.splint(a, b, tck) for a, b in pairs])
    np.savez(os.path.join(OUT, "fitpack_case.npz"), x=x, y=y, t=tck[0], c=tck[1][: len(x)], levels=levels,
             roots=np.array([np.pad(r, (0, 64 - len(r)), constant_values=np.nan) for r in roots]), pairs=pairs, ints=ints)


def launch_known():
    def count_and_sum(N_rings, min_az, w=0.0174):
        gx, gw = np.polynomial.hermite.hermgauss(2 * N_rings + 2)
        r = gx[N_rings + 1:] * w / np.sqrt(2.0)
        rw = gw[N_rings + 1:] * w / np.sqrt(2.0)
        nth = [max(1, int(np.round(min_az * r[i] / r[0]))) for i in range(N_rings)]
        tot = sum(r[i] * rw[i] * 2 * np.pi for i in range(N_rings)) * 2.0 / (w**2 * np.pi)
        return nth, tot
    cases = [(3, 5), (21, 11), (7, 20), (66, 14), (88, 120)]
    out = {}
    for nr, ma in cases:
        nth, tot = count_and_sum(nr, ma)
        out[f"n_{nr}_{ma}"] = np.array(sum(nth)); out[f"nth_{nr}_{ma}"] = np.array(nth); out[f"sumw_{nr}_{ma}"] = np.array(tot)
    np.savez(os.path.join(OUT, "launch_known.npz"), **out)


def bessel_gl():
    z = np.linspace(0.0, 8.0, 81)
    J = np.array([special.jv(n, z) for n in range(1, 5)])
    t, w = np.polynomial.legendre.leggauss(24)
    np.savez(os.path.join(OUT, "bessel_gl.npz"), z=z, J=J, gl_t=t, gl_w=w)


def oracle_ray():
    import torj_jl_b200 as tj
    from oracle import torj_oracle as O
    arr = tj.solovev_arrays(257, 257)
    pl = O.OraclePlasma(*arr.values())
    gl = np.polynomial.legendre.leggauss(24)
    x0 = np.array([2.5, 0.0, 0.4]); N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.0)
    psi = np.linspace(0, 1, 1000)
    r = pl.make_ray(x0, N0, 95e9, 1, 0.4, psi, gl)
    sl = slice(None, None, 50)
    np.savez(os.path.join(OUT, "oracle_ray.npz"), n=len(r["s"]), s=r["s"][sl], x=r["x"][sl], y=r["y"][sl], z=r["z"][sl],
             P=r["P"][sl], deposited=r["deposited_power"], P_end=r["P"][-1], dP_dV=r["dP_dV"][::10], counters=r["counters"])


def warm_alpha():
    from oracle import torj_oracle as O
    rows, outs = [], []
    for f in (95e9, 110e9, 170e9):
        for X in (0.05, 0.3):
            for Y in (0.31, 0.34, 0.49, 0.505, 0.52, 0.98, 1.03):
                for th in (np.pi / 2 - 1e-3, 1.35, 1.1, 2.0):
                    for te in (800.0, 5e3, 25e3):
                        for imod in (1, -1):
                            row = [2 * np.pi * f, X, Y, 0.9 - 0.3 * X, th, te, 0.55, imod]
                            o = O.warm_alpha(*row[:7], imod)
                            rows.append(row)
                            outs.append([o["N_warm"], o["alpha"], o["lrm"], o["ierr"], o["iterations"]])
    np.savez(os.path.join(OUT, "warm_alpha.npz"), inputs=np.array(rows), outputs=np.array(outs))


if __name__ == "__main__":
    fitpack_case(); launch_known(); bessel_gl(); oracle_ray(); warm_alpha()
    print("golden fixtures written to", OUT)

"""torj_warm_alpha on one point per resident lane (37 888 points) in the regime of the config-5 scan: the alpha call alone
(warp-cooperative quadrature + per-lane solve), for ncu and for a calls/s figure."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj

n = 37888
k = np.arange(n)
om = 2 * np.pi * np.where(k % 2 == 0, 110e9, 170e9)
Y = np.where(k % 2 == 0, 0.46, 0.30) + 0.10 * ((k // 2) % 97) / 97.0          # around the 2nd / 3rd harmonic
X = 0.05 + 0.25 * ((k // 194) % 11) / 11.0
th = 1.15 + 0.7 * ((k // 2134) % 7) / 7.0
te = np.array([10e3, 15e3, 25e3])[(k // 14938) % 3]
Nr = 0.97 - 0.4 * X
tj.warm_alpha(om[:64], X[:64], Y[:64], Nr[:64], th[:64], te[:64], 0.55, 1)      # warm-up
t0 = time.perf_counter()
Nw, al = tj.warm_alpha(om, X, Y, Nr, th, te, 0.55, 1)
dt = time.perf_counter() - t0
print(f"{n} alpha calls in {dt*1e3:.1f} ms end to end ({n/dt:.3e} calls/s incl. copies); alpha range {np.nanmin(al):.3e} .. {np.nanmax(al):.3e}; "
      f"lrm histogram {np.bincount(tj.warm_alpha.last['lrm'])}")

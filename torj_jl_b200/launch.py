"""launch_peripheral_rays — host-side mirror of reference src/launch.jl:24-132 (Gaussian beam -> concentric-ring
ray bundle with quadrature weights).  Same arguments, keywords, return shapes and error as the reference,
quirks included."""
from __future__ import annotations

import numpy as np

C_LIGHT = 2.99792458e8  # reference src/constants.jl:15


def launch_peripheral_rays(x0, N0, w, inverse_curvature_radius, f, *, N_rings=3, min_azimuthal_points=5,
                           normalize_weight_sum=True, **kwargs):
    """Returns (ray_positions[N,3], ray_directions[N,3], ray_weights[N])."""
    if N_rings < 2:
        raise ValueError(f"N_rings = {N_rings} < 2 which is the minimum")  # ArgumentError, src/launch.jl:27-29
    x0 = np.asarray(x0, dtype=np.float64)
    N0 = np.asarray(N0, dtype=np.float64)
    n0 = N0 / np.linalg.norm(N0)
    curved = np.isfinite(inverse_curvature_radius)
    if curved:  # src/launch.jl:34-47
        R_curv = 1.0 / inverse_curvature_radius
        lam = C_LIGHT / f
        w0 = (lam * abs(R_curv) * w) / np.sqrt(lam**2 * R_curv**2 + np.pi**2 * w**4)
        z_waist = np.pi**2 * R_curv * w**4 / (lam**2 * R_curv**2 + np.pi**2 * w**4)
        x_waist = x0 - n0 * z_waist
    else:
        w0 = w
    e_chi = np.array([1.0, 0.0, -n0[0] / n0[2]])                      # src/launch.jl:54-57
    e_ups = np.array([-n0[0] * n0[1] / n0[2], n0[2] - n0[0], -n0[1]])  # src/launch.jl:61-64
    e_chi /= np.linalg.norm(e_chi)
    e_ups /= np.linalg.norm(e_ups)
    gx, gw = np.polynomial.hermite.hermgauss(2 * N_rings + 2)         # FastGaussQuadrature.gausshermite
    r_pts = gx[N_rings + 1:] * (w / np.sqrt(2.0))                     # v[N_rings+2:end], src/launch.jl:72-76
    r_weights = gw[N_rings + 1:] * (w / np.sqrt(2.0))
    N_theta = np.array([max(1, int(np.round(min_azimuthal_points * r_pts[i] / r_pts[0]))) for i in range(N_rings)])
    total = int(N_theta.sum())
    pos = np.zeros((total, 3)); dirs = np.zeros((total, 3)); wts = np.zeros(total)
    k = 0
    for i in range(N_rings):
        nt = int(N_theta[i])
        th = 2.0 * np.pi * np.arange(nt) / nt
        chi = r_pts[i] * np.cos(th); ups = r_pts[i] * np.sin(th)
        off = chi[:, None] * e_chi[None, :] + ups[:, None] * e_ups[None, :]
        p = off + x0
        pos[k:k + nt] = p
        if curved:  # src/launch.jl:102-113
            d = w0 / w * off * np.sign(inverse_curvature_radius) + x_waist
            d = d - p if inverse_curvature_radius < 0.0 else -d + p
            d /= np.linalg.norm(d, axis=1)[:, None]
        else:
            d = np.tile(n0, (nt, 1))
        dirs[k:k + nt] = d
        wts[k:k + nt] = r_pts[i] * r_weights[i] * (2.0 * np.pi / nt)    # src/launch.jl:120
        k += nt
    if normalize_weight_sum:
        wts /= wts.sum()
    else:
        wts *= 2.0 / (w**2 * np.pi)
    return pos, dirs, wts

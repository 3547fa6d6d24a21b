"""Synthetic Solov'ev equilibrium: the arrays ``Plasma(...)`` takes (reference src/plasma.jl:30-32).

Stands in for the reference's test artifact (test/Artifacts.toml:4-7, not available offline) in configs 1-5
of BASELINE.json.  Specification: SURVEY.md §8(d) —
    psi(R,Z) = A[(R^2-R0^2)^2/4 + R^2 Z^2/E^2],  psi_N = psi/psi(R0+a, 0)
    B_R = -(1/R) dpsi/dZ,  B_Z = (1/R) dpsi/dR,  B_phi = B0 R0 / R
    n_e = ne0 (1-psi_N) + 1e17 m^-3,  T_e = Te0 (1-psi_N)^2 + 50 eV   on 101 uniform psi_N in [0,1]
    V(psi_N) = 2 pi^2 R0 a^2 E psi_N
No random numbers anywhere, hence no seeds.
"""
from __future__ import annotations

import numpy as np


def solovev_arrays(nR: int = 257, nZ: int = 257, *, R0: float = 1.7, a: float = 0.6, E: float = 1.7,
                   A: float = 0.08, B0: float = 2.0, ne0: float = 3.0e19, Te0: float = 4.0e3,
                   R_range=(0.5, 2.7), Z_range=(-1.4, 1.4), nprof: int = 101) -> dict:
    """Returns a dict of the eleven positional arguments of ``Plasma`` (2-D arrays shaped [nR, nZ])."""
    R = np.linspace(R_range[0], R_range[1], nR)
    Z = np.linspace(Z_range[0], Z_range[1], nZ)
    RR, ZZ = np.meshgrid(R, Z, indexing="ij")
    psi = A * ((RR**2 - R0**2) ** 2 / 4.0 + RR**2 * ZZ**2 / E**2)
    psi_b = A * (((R0 + a) ** 2 - R0**2) ** 2 / 4.0)
    psi_norm = psi / psi_b
    BR = -2.0 * A * RR * ZZ / E**2
    BZ = A * ((RR**2 - R0**2) + 2.0 * ZZ**2 / E**2)
    Bphi = B0 * R0 / RR
    psi_prof = np.linspace(0.0, 1.0, nprof)
    ne_prof = ne0 * (1.0 - psi_prof) + 1.0e17
    Te_prof = Te0 * (1.0 - psi_prof) ** 2 + 50.0
    vol = 2.0 * np.pi**2 * R0 * a**2 * E * psi_prof
    return dict(R_coords=R, Z_coords=Z, psi_norm_data=psi_norm, psi_prof=psi_prof, ne_prof=ne_prof,
                Te_prof=Te_prof, Br_data=BR, Bz_data=BZ, Bphi_data=Bphi, eqt1d_psi_norm=psi_prof.copy(),
                eqt1d_volume=vol)


# launcher of reference test/tests/setup.jl:64-74, re-used on the Solov'ev box
LAUNCHER = dict(R0=2.5, phi0=0.0, z0=0.4, spot_size=0.0174, inverse_curvature_radius=1.0 / 3.99,
                steering_angle_pol=np.deg2rad(30.0), steering_angle_tor=0.0, f=95.0e9, mode=1)


def pol_tor_angles_2_vector(pol: float, tor: float) -> np.ndarray:
    """IMAS.pol_tor_angles_2_vector (un-vendored; reference src/solve.jl:211): IMAS DD convention
    (k_R, k_phi, k_Z) = (-cos pol cos tor, sin tor, -sin pol cos tor), used by the reference directly as a
    Cartesian direction."""
    return np.array([-np.cos(pol) * np.cos(tor), np.sin(tor), -np.sin(pol) * np.cos(tor)])

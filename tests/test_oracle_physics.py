"""Oracle invariants that hold without the reference's artifact (SURVEY.md §8(c) 'what still pins results') and the
regression fixture.  CPU only; sizes chosen so the file runs in well under a minute."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import torj_oracle as O

PSI = np.linspace(0.0, 1.0, 1000)


def test_oracle_is_multithreaded():
    assert O.max_threads() >= 1


def test_constants_and_derived(oracle_full):
    # 2 f_ce(R0) = 112 GHz on the Solov'ev equilibrium
    e = oracle_full.eval_plasma([1.7, 0.0, 0.0], [1.0, 0.0, 0.0], 2 * np.pi * 112e9)
    assert abs(2 * e["Y"] - 1.0) < 2e-2
    assert abs(np.linalg.norm(e["b"]) - 1.0) < 1e-14


def test_rhs_matches_finite_differences(oracle_full, gl24, launcher):
    st, init = oracle_full.ray_init(launcher["x0"], launcher["N0"], launcher["f"], 1)
    assert st == 0
    u = np.concatenate([init[:6], [1.0]])
    for _ in range(12):
        u = u + 0.02 * oracle_full.rhs(u, launcher["f"], 1, gl24)
    du = oracle_full.rhs(u, launcher["f"], 1, gl24)
    om = 2 * np.pi * launcher["f"]
    lam = lambda v: oracle_full.eval_plasma(v[:3], v[3:6], om, 1)["Lambda"]
    g = np.zeros(6)
    for k in range(6):
        h = 1e-6
        up, um = u.copy(), u.copy()
        up[k] += h; um[k] -= h
        g[k] = (lam(up) - lam(um)) / (2 * h)
    n = np.linalg.norm(g[3:6])
    assert np.allclose(du[:3], g[3:6] / n, atol=2e-8)
    assert np.allclose(du[3:6], -g[:3] / n, atol=2e-7)
    assert abs(np.linalg.norm(du[:3]) - 1.0) < 1e-14      # arc-length parametrisation


def test_config1_single_ray_invariants_and_golden(oracle_full, gl24, launcher):
    r = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], 1, 0.4, PSI, gl24)
    assert r["status"] == 0
    g = np.load(os.path.join(GOLDEN, "oracle_ray.npz"))
    assert len(r["s"]) == int(g["n"])
    for k in ("s", "x", "y", "z", "P"):
        assert np.abs(r[k][::50] - g[k]).max() < 1e-12
    assert abs(r["deposited_power"] - float(g["deposited"])) < 1e-12
    # first two samples: launch point and plasma entry (reference src/solve.jl:149-153)
    assert r["s"][0] == 0.0 and np.allclose([r["x"][0], r["y"][0], r["z"][0]], launcher["x0"])
    assert abs(r["psi"][1] - oracle_full.psi_prof_max) < 1e-6 and r["psi"][1] <= oracle_full.psi_prof_max
    assert r["P"][0] == r["P"][1] == 1.0 and r["dP_ds"][0] == r["dP_ds"][1] == 0.0
    # steps are pinned at dtmax outside the absorption layer; segments end exactly on the 100-segment grid
    ds = np.diff(r["s"][1:])
    assert ds.max() <= 1e-4 + 1e-14 and np.median(ds) > 0.99e-4
    # P decreasing, all absorbed at the second harmonic
    assert np.all(np.diff(r["P"]) <= 1e-16) and r["P"][-1] < 1e-6
    # the profile integrates to the absorbed power (reference test/tests/test_make_beam.jl:14-21)
    assert abs(r["deposited_power"] - (1.0 - r["P"][-1])) < 1e-3


def test_dispersion_conserved_along_ray(oracle_full, gl24, launcher):
    r = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], 1, 0.2, PSI, gl24)
    assert r["status"] == 0
    e = oracle_full.eval_plasma(r["u_final"][:3], r["u_final"][3:6], 2 * np.pi * launcher["f"], 1)
    assert abs(e["Lambda"]) < 1e-9


def test_streaming_deposition_matches_faithful(oracle_full, gl24, launcher):
    a = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], 1, 0.4, PSI, gl24, deposition="faithful")
    b = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], 1, 0.4, PSI, gl24, deposition="streaming")
    l2 = np.linalg.norm(a["dP_dV"] - b["dP_dV"]) / np.linalg.norm(a["dP_dV"])
    assert l2 < 1e-6, l2
    assert abs(a["deposited_power"] - b["deposited_power"]) < 1e-7
    assert a["dP_dV"][-1] == 0.0 and b["dP_dV"][-1] == 0.0


def test_volume_integral_identity(oracle_full, gl24, launcher):
    """reference test/tests/test_make_beam.jl:23-31: sum dV/dpsi * dP/dV * dpsi == absorbed fraction (1e-3)."""
    r = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], 1, 0.4, PSI, gl24)
    dpsi = PSI[1] - PSI[0]
    dVdpsi = (oracle_full.volume(PSI + 1e-6) - oracle_full.volume(PSI - 1e-6)) / 2e-6
    P_test = float(np.sum(dVdpsi * r["dP_dV"] * dpsi))
    assert abs(P_test - r["deposited_power"]) < 1e-3


def test_low_temperature_gate_and_vacuum_limit(arrays_small, gl24, launcher):
    arr = dict(arrays_small)
    arr["Te_prof"] = np.full_like(arr["Te_prof"], 10.0)          # below the 20 eV gate (src/absorption.jl:194)
    arr["ne_prof"] = np.full_like(arr["ne_prof"], 1e10)          # vacuum-like: straight ray
    pl = O.OraclePlasma(*arr.values())
    r = pl.make_ray(launcher["x0"], launcher["N0"], launcher["f"], 1, 0.3, np.linspace(0, 1, 50), gl24)
    assert r["status"] == 0 and r["P"][-1] == 1.0 and r["counters"][3] == 0
    d = np.array([r["x"][-1] - r["x"][1], r["y"][-1] - r["y"][1], r["z"][-1] - r["z"][1]])
    assert np.allclose(d / np.linalg.norm(d), launcher["N0"], atol=1e-8)


def test_init_failures(oracle_small, arrays_small, launcher):
    st, _ = oracle_small.ray_init(launcher["x0"], -launcher["N0"], launcher["f"], 1)   # pointing away: no bracket
    assert st == 2
    dense = dict(arrays_small)
    dense["ne_prof"] = np.full_like(dense["ne_prof"], 1e20)                            # X = 2.2 at 60 GHz at the edge
    st, _ = O.OraclePlasma(*dense.values()).ray_init(launcher["x0"], launcher["N0"], 60e9, 1)
    assert st == 1                                                                     # cut-off, reference src/solve.jl:55-59


def test_off_grid_launcher_enters_box(oracle_small, launcher):
    x_far = launcher["x0"] - 0.9 * launcher["N0"]          # R ~ 3.28 > R_last = 2.7
    st0, a = oracle_small.ray_init(launcher["x0"], launcher["N0"], launcher["f"], 1)
    st1, b = oracle_small.ray_init(x_far, launcher["N0"], launcher["f"], 1)
    assert st0 == 0 and st1 == 0
    assert np.abs(a[:6] - b[:6]).max() < 1e-9 and abs(b[6] - a[6] - 0.9) < 1e-9


def test_beam46_self_consistency(oracle_small, gl24, launcher):
    """reference test/tests/test_make_beam.jl:14-21 on the default 46-ray bundle (coarse grid, short path)."""
    pos, dirs, w = O.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"])
    psi = np.linspace(0, 1, 200)
    r = oracle_small.trace_bundle(pos, dirs, w, launcher["f"], 1, 0.5, psi, gl24)
    assert (r["status"] == 0).all()
    absorbed = 1.0 - float(np.sum(w * r["P_final"]))
    assert abs(r["deposited_power"] - absorbed) < 1e-3
    s = oracle_small.trace_bundle(pos, dirs, w, launcher["f"], 1, 0.5, psi, gl24, deposition="streaming")
    assert np.linalg.norm(s["dP_dV"] - r["dP_dV"]) / np.linalg.norm(r["dP_dV"]) < 1e-6
    assert s["counters"]["n_acc"] == r["counters"]["n_acc"] == int(np.sum(r["n_points"] - 2))


def test_per_ray_frequency_with_one_mode(oracle_small, gl24, launcher):
    """trace_bundle(freq per ray, mode scalar): the mode is broadcast (the C side indexes mode[i] whenever the frequency is per
    ray; a length-1 array there read garbage modes for i > 0)."""
    import torj_jl_b200 as tj
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"],
                                             N_rings=2, min_azimuthal_points=3)
    psi = np.linspace(0, 1, 50)
    a = oracle_small.trace_bundle(pos, dirs, w, launcher["f"], 1, 0.5, psi, gl24, deposition="streaming")
    b = oracle_small.trace_bundle(pos, dirs, w, np.full(len(w), launcher["f"]), 1, 0.5, psi, gl24, deposition="streaming")
    assert np.array_equal(a["n_points"], b["n_points"]) and np.array_equal(a["P_final"], b["P_final"])
    c = oracle_small.trace_bundle(pos, dirs, w, np.full(len(w), launcher["f"]), -1, 0.5, psi, gl24, deposition="streaming")
    assert not np.array_equal(a["P_final"], c["P_final"])       # O-mode is a different ray

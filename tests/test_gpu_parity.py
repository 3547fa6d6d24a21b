"""Parity of the CUDA path (through the C ABI of libtorj_cuda.so) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star, SURVEY.md §8(d) "Parity gates"), all FP64:
  trajectories            <= 1e-9 m at equal step sequences (same number of saved points required)
  final absorbed fraction <= 1e-6 relative
  deposition profile      ||d||_2/||ref||_2 <= 1e-4 against the reference's spline-root algorithm ("faithful"),
                          <= 1e-9 against the oracle's restatement of the streaming algorithm (like for like)
"""
import numpy as np
import pytest

import torj_jl_b200 as tj
from oracle import torj_oracle as O

pytestmark = pytest.mark.gpu

PSI = np.linspace(0.0, 1.0, 1000)  # reference test/tests/setup.jl:77
TRAJ_TOL = 1e-9
FRAC_TOL = 1e-6
L2_FAITHFUL = 1e-4
L2_LIKE = 1e-9


def EXACT():
    """alpha_floor = 0: every harmonic integral evaluated, as the reference does."""
    return tj.default_options(alpha_floor=0.0)


def l2rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module", autouse=True)
def _abs_al_init():
    tj.abs_Al_init(24)  # reference test/tests/setup.jl:80


@pytest.fixture(scope="module")
def gpu_full(arrays_full):
    tj.abs_Al_init(24)  # reference test/tests/setup.jl:80
    return tj.Plasma(*arrays_full.values())


@pytest.fixture(scope="module")
def gpu_small(arrays_small):
    tj.abs_Al_init(24)
    return tj.Plasma(*arrays_small.values())


def _ray_points(res, i=0):
    n = int(res["n_points"][i])
    return n, res["traj_s"][i, :n], res["traj_xyz"][i, :, :n], res["traj_P"][i, :n], res["traj_dP_ds"][i, :n]


# ------------------------------------------------------------------------------------------------------------------
# fields, dispersion, absorption, RHS  (reference test/tests/test_trajectory.jl, test_absorption.jl)
# ------------------------------------------------------------------------------------------------------------------
def test_fields_and_dispersion_along_ray(gpu_full, oracle_full, gl24, launcher):
    ro = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], 1, 0.4, PSI, gl24)
    idx = np.linspace(1, len(ro["s"]) - 1, 200).astype(int)
    X = np.stack([ro["x"][idx], ro["y"][idx], ro["z"][idx]], 1)
    rng = np.random.default_rng(1)
    N = rng.normal(size=X.shape) * 0.5
    pr = gpu_full.probe(X, N, launcher["f"], 1)
    om = 2 * np.pi * launcher["f"]
    for i in range(0, len(idx), 7):
        e = oracle_full.eval_plasma(X[i], N[i], om, 1)
        assert abs(pr["X"][i] - e["X"]) <= 1e-13 * e["X"]
        assert abs(pr["Y"][i] - e["Y"]) < 1e-14
        assert abs(pr["N_par"][i] - e["N_par"]) < 1e-14
        assert abs(pr["Te"][i] - e["Te"]) <= 1e-12 * e["Te"]
        assert abs(pr["Lambda"][i] - e["Lambda"]) < 1e-13
        assert abs(pr["psi"][i] - ro["psi"][idx[i]]) < 1e-14
        Babs = e["Y"] * 9.1093837015e-31 * om / 1.602176634e-19
        assert np.abs(pr["B"][i] - e["b"] * Babs).max() < 1e-6          # test_trajectory.jl:10 asks 1e-6 T


def test_fields_outside_the_grid_use_line_extrapolation(gpu_small, oracle_small, arrays_small):
    R = arrays_small["R_coords"]; Z = arrays_small["Z_coords"]
    pts = np.array([[R[-1] + 0.3, 0.0, 0.2], [0.0, R[0] - 0.1, Z[0] - 0.2], [1.2, 0.9, Z[-1] + 0.05], [R[-1], 0.0, Z[-1]]])
    N = np.tile([0.3, 0.2, -0.1], (len(pts), 1))
    pr = gpu_small.probe(pts, N, 95e9, 1)
    for i, p in enumerate(pts):
        v, _, _ = oracle_small.spline("psi", [np.hypot(p[0], p[1])], [p[2]])
        assert abs(pr["psi"][i] - v[0]) < 1e-12 * max(1.0, abs(v[0]))
        e = oracle_small.eval_plasma(p, N[i], 2 * np.pi * 95e9, 1)
        assert abs(pr["Y"][i] - e["Y"]) < 1e-13 and abs(pr["N_par"][i] - e["N_par"]) < 1e-13


def test_rhs_matches_dual_number_oracle(gpu_full, oracle_full, gl24, launcher):
    """Hand-derived gradients (kernel) vs ForwardDiff-style duals (oracle), X and O mode, with toroidal angle."""
    for mode, tor in ((1, 0.0), (-1, 0.0), (1, 0.25)):
        N0 = tj.pol_tor_angles_2_vector(np.deg2rad(25.0), tor)
        st, init = oracle_full.ray_init(launcher["x0"], N0, launcher["f"], mode)
        assert st == 0
        u = np.concatenate([init[:6], [0.7]])
        U = []
        for _ in range(45):
            U.append(u.copy())
            u = u + 0.008 * oracle_full.rhs(u, launcher["f"], mode, gl24)
        U = np.array(U)
        dg = gpu_full.rhs(U, launcher["f"], mode, options=EXACT())
        do = np.array([oracle_full.rhs(v, launcher["f"], mode, gl24) for v in U])
        assert np.abs(dg[:, :3] - do[:, :3]).max() < 1e-13
        assert np.abs(dg[:, 3:6] - do[:, 3:6]).max() < 1e-11
        scale = np.maximum(np.abs(do[:, 6]), 1e-300)
        assert (np.abs(dg[:, 6] - do[:, 6]) / scale)[np.abs(do[:, 6]) > 1e-250].max() < 1e-11


def test_absorption_coefficient_scan(gpu_full, oracle_full, gl24):
    """alpha over the second/third-harmonic layer incl. Bessel-series range switches; Te gate (src/absorption.jl:194)."""
    R = np.linspace(1.45, 2.28, 120)
    X = np.stack([R, np.zeros_like(R), 0.05 * np.ones_like(R)], 1)
    for ang in (0.0, 0.3, 0.6):
        N = np.tile([-0.9 * np.cos(ang), 0.9 * np.sin(ang), -0.1], (len(R), 1))
        for mode in (1, -1):
            a = gpu_full.probe(X, N, 95e9, mode, options=EXACT())["alpha"]
            pruned = gpu_full.probe(X, N, 95e9, mode)["alpha"]          # default alpha_floor = 1e-14 1/m
            assert np.abs(pruned - a).max() <= 2e-14 and np.any(pruned != a) and np.any(pruned == 0.0)
            ref = -np.array([oracle_full.rhs(np.concatenate([X[i], N[i], [1.0]]), 95e9, mode, gl24)[6] for i in range(len(R))])
            big = np.abs(ref) > 1e-200
            assert np.abs(a[big] - ref[big]).max() <= 1e-10 * np.abs(ref[big]).max()
            assert (np.abs(a[big] - ref[big]) / np.abs(ref[big])).max() < 1e-9
    opt = tj.default_options(te_min=1e9)
    assert np.all(gpu_full.probe(X, N, 95e9, 1, options=opt)["alpha"] == 0.0)


# ------------------------------------------------------------------------------------------------------------------
# ray initialisation
# ------------------------------------------------------------------------------------------------------------------
def test_ray_init_matches_oracle(gpu_full, oracle_full, launcher):
    N0 = tj.pol_tor_angles_2_vector(np.deg2rad(28.0), 0.15)
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], N0, launcher["spot"], launcher["inv_Rc"], launcher["f"])
    res = tj.trace_bundle(gpu_full, pos, dirs, w, launcher["f"], 1, 1e-3, PSI, options=tj.default_options(n_segments=1),
                          trajectories=(0, len(w)), traj_max_pts=16)
    assert (res["status"] == 0).all()
    for i in range(0, len(w), 5):
        st, init = oracle_full.ray_init(pos[i], dirs[i], launcher["f"], 1)
        assert st == 0
        assert np.abs(res["traj_xyz"][i, :, 1] - init[:3]).max() < 1e-12
        assert abs(res["traj_s"][i, 1] - init[6]) < 1e-12


def test_init_failure_statuses_do_not_abort_the_batch(gpu_small, arrays_small, launcher):
    pos = np.array([launcher["x0"], launcher["x0"], launcher["x0"] - 0.9 * launcher["N0"]])
    dirs = np.array([launcher["N0"], -launcher["N0"], launcher["N0"]])
    res = tj.trace_bundle(gpu_small, pos, dirs, [0.5, 0.25, 0.25], launcher["f"], 1, 0.3, np.linspace(0, 1, 64))
    assert list(res["status"]) == [0, 2, 0]                 # second ray points away: no bracket (src/solve.jl:29)
    assert res["counters"]["n_rays_ok"] == 2 and res["P_final"][1] == 0.0 and res["P_deposited_ray"][1] == 0.0
    assert res["n_points"][1] == -2                         # diagnostic of the failing check, never a power (include/torj_cuda.h)
    assert abs(res["P_final"][0] - res["P_final"][2]) < 1e-9   # off-grid launcher enters through the box (src/solve.jl:22-25)
    dense = dict(arrays_small); dense["ne_prof"] = np.full_like(dense["ne_prof"], 1e20)
    r2 = tj.trace_bundle(tj.Plasma(*dense.values()), pos[:1], dirs[:1], [1.0], 60e9, 1, 0.3, np.linspace(0, 1, 64))
    assert list(r2["status"]) == [1]                        # cut-off at entry (src/solve.jl:55-59)
    with pytest.raises(AssertionError, match="cut-off"):
        tj.make_ray(tj.Plasma(*dense.values()), pos[0], dirs[0], 60e9, 1, 0.3, np.linspace(0, 1, 64))


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json config 1: single EC ray (stand-in for reference test/tests/test_make_ray.jl)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,scheme", [(1, 0), (-1, 0), (1, 1)])
def test_config1_single_ray_trajectory(gpu_full, oracle_full, gl24, launcher, mode, scheme):
    opt = tj.default_options(scheme=scheme)
    s, u, P, prof, dep = tj.make_ray(gpu_full, launcher["x0"], launcher["N0"], launcher["f"], mode, 0.4, PSI, options=opt)
    ro = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], mode, 0.4, PSI, gl24,
                              opts=O.OracleOptions.default(scheme=scheme))
    assert ro["status"] == 0 and len(s) == len(ro["s"])
    xyz = np.array(u)
    assert np.abs(s - ro["s"]).max() < 1e-12
    assert max(np.abs(xyz[:, 0] - ro["x"]).max(), np.abs(xyz[:, 1] - ro["y"]).max(), np.abs(xyz[:, 2] - ro["z"]).max()) < TRAJ_TOL
    assert np.abs(P - ro["P"]).max() < 1e-9
    if ro["deposited_power"] > 1e-8:
        assert abs(dep - ro["deposited_power"]) <= FRAC_TOL * ro["deposited_power"]
        assert l2rel(prof, ro["dP_dV"]) < L2_FAITHFUL
        rs = oracle_full.make_ray(launcher["x0"], launcher["N0"], launcher["f"], mode, 0.4, PSI, gl24,
                                  opts=O.OracleOptions.default(scheme=scheme), deposition="streaming")
        assert l2rel(prof, rs["dP_dV"]) < L2_LIKE


def test_alpha_floor_pruning_is_invisible_at_tolerance(gpu_full, launcher):
    """Default alpha_floor (1e-14 1/m) against exact evaluation on the default 46-ray beam."""
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"])
    a = tj.trace_bundle(gpu_full, pos, dirs, w, launcher["f"], 1, 1.0, PSI)
    b = tj.trace_bundle(gpu_full, pos, dirs, w, launcher["f"], 1, 1.0, PSI, options=EXACT())
    ca, cb = a["counters"], b["counters"]
    assert ca["n_harm_pruned"] > 0 and ca["n_alpha_skipped"] > 0 and cb["n_harm_pruned"] == 0 and cb["n_alpha_skipped"] == 0
    assert ca["n_alpha"] + ca["n_alpha_skipped"] == cb["n_alpha"] == cb["n_rhs"] == ca["n_rhs"]
    assert ca["n_harm"] < cb["n_harm"]
    assert np.array_equal(a["n_points"], b["n_points"])
    assert np.abs(a["P_final"] - b["P_final"]).max() < 1e-13
    assert abs(a["deposited_power"] - b["deposited_power"]) < 1e-12
    assert l2rel(a["dP_dV"], b["dP_dV"]) < 1e-12


def test_make_ray_return_shapes(gpu_small, launcher):
    """reference src/solve.jl:180: (s, u, P_beam, dP_dV_ray, deposited_power)."""
    psi = np.linspace(0, 1, 50)
    s, u, P, prof, dep = tj.make_ray(gpu_small, launcher["x0"], launcher["N0"], launcher["f"], 1, 0.1, psi)
    assert s.ndim == 1 and len(u) == len(s) == len(P) and u[0].shape == (3,) and prof.shape == (50,) and isinstance(dep, float)
    assert s[0] == 0.0 and np.allclose(u[0], launcher["x0"]) and P[0] == P[1] == 1.0 and prof[-1] == 0.0
    assert len(s) == 2 + 100 * 10                                   # 100 segments of 1 mm at dtmax = 0.1 mm
    assert abs(s[-1] - s[1] - 0.1) < 1e-12


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json config 2: 1 025-ray beam, absorbed fraction + deposition profile
# ------------------------------------------------------------------------------------------------------------------
def test_config2_beam_1025_rays(gpu_full, oracle_full, gl24, launcher):
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"],
                                             N_rings=7, min_azimuthal_points=20)
    assert len(w) == 1025
    res = tj.trace_bundle(gpu_full, pos, dirs, w, launcher["f"], 1, 1.0, PSI)
    ref = oracle_full.trace_bundle(pos, dirs, w, launcher["f"], 1, 1.0, PSI, gl24, deposition="faithful", also_streaming=True)
    assert (res["status"] == 0).all() and (ref["status"] == 0).all()
    assert np.array_equal(res["n_points"], ref["n_points"])
    assert res["counters"]["n_acc"] == int(ref["counters"]["n_acc"]) and res["counters"]["n_rej"] == int(ref["counters"]["n_rej"])
    assert np.abs(res["P_final"] - ref["P_final"]).max() < 1e-12
    assert abs(res["deposited_power"] - ref["deposited_power"]) <= FRAC_TOL * ref["deposited_power"]
    assert l2rel(res["dP_dV"], ref["dP_dV"]) < L2_FAITHFUL
    assert l2rel(res["dP_dV"], ref["dP_dV_streaming"]) < L2_LIKE
    # identities of reference test/tests/test_make_beam.jl:14-31
    absorbed = 1.0 - float(np.sum(w * res["P_final"]))
    assert abs(res["deposited_power"] - absorbed) < 1e-3
    dpsi = PSI[1] - PSI[0]
    dVdpsi = (gpu_full.volume(PSI + 1e-6) - gpu_full.volume(PSI - 1e-6)) / 2e-6
    assert abs(float(np.sum(dVdpsi * res["dP_dV"] * dpsi)) - res["deposited_power"]) < 1e-3


def test_make_beam_default_bundle(gpu_full, oracle_full, gl24, launcher):
    """make_beam(...) of reference src/solve.jl:209-242 with the default 46-ray bundle and kwargs pass-through."""
    arc, traj, powers, prof, dep, w = tj.make_beam(gpu_full, 2.5, 0.0, 0.4, 0.0, np.deg2rad(30.0), launcher["spot"],
                                                    launcher["inv_Rc"], launcher["f"], 1, 1.0, PSI)
    assert len(arc) == len(traj) == len(powers) == len(w) == 46 and prof.shape == PSI.shape
    pos, dirs, w2 = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"])
    ref = oracle_full.trace_bundle(pos, dirs, w2, launcher["f"], 1, 1.0, PSI, gl24)
    assert abs(dep - ref["deposited_power"]) <= FRAC_TOL * ref["deposited_power"]
    assert l2rel(prof, ref["dP_dV"]) < L2_FAITHFUL
    assert all(abs(powers[i][-1] - ref["P_final"][i]) < 1e-12 for i in range(46))
    out = tj.make_beam(gpu_full, 2.5, 0.0, 0.4, 0.0, np.deg2rad(30.0), launcher["spot"], launcher["inv_Rc"], launcher["f"], 1,
                       0.2, PSI, N_rings=2, min_azimuthal_points=3)
    assert len(out[5]) == 3 + round(3 * np.polynomial.hermite.hermgauss(6)[0][4] / np.polynomial.hermite.hermgauss(6)[0][3])


# ------------------------------------------------------------------------------------------------------------------
# weak absorption / exits / odd grids / per-ray frequency
# ------------------------------------------------------------------------------------------------------------------
def test_low_density_o_mode_rays_leave_the_plasma(arrays_small, gl24, launcher):
    arr = dict(arrays_small); arr["ne_prof"] = arr["ne_prof"] * 0.3; arr["Te_prof"] = arr["Te_prof"] * 0.25
    pl = tj.Plasma(*arr.values()); opl = O.OraclePlasma(*arr.values())
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], 140e9,
                                             N_rings=2, min_azimuthal_points=3)
    psi = np.concatenate([np.linspace(0, 0.5, 40) ** 2 * 2, np.linspace(0.51, 1.0, 70)])   # non-uniform levels
    res = tj.trace_bundle(pl, pos, dirs, w, 140e9, -1, 3.0, psi, trajectories=(0, 2))
    ref = opl.trace_bundle(pos, dirs, w, 140e9, -1, 3.0, psi, gl24)
    assert (res["status"] == 0).all() and np.array_equal(res["n_points"], ref["n_points"])
    assert res["n_points"].max() < 2 + 100 * 300                   # stopped by psi > 1 (src/solve.jl:174), not by s_max
    assert np.abs(res["P_final"] - ref["P_final"]).max() < 1e-10
    assert 0.0 < ref["deposited_power"] < 0.999
    assert abs(res["deposited_power"] - ref["deposited_power"]) <= FRAC_TOL * ref["deposited_power"]
    assert l2rel(res["dP_dV"], ref["dP_dV"]) < L2_FAITHFUL
    ro = opl.make_ray(pos[1], dirs[1], 140e9, -1, 3.0, psi, gl24)
    n, s, xyz, P, dP = _ray_points(res, 1)
    assert n == len(ro["s"]) and np.abs(xyz[0] - ro["x"]).max() < TRAJ_TOL and np.abs(xyz[2] - ro["z"]).max() < TRAJ_TOL
    assert np.abs(dP - ro["dP_ds"]).max() <= 1e-9 * max(1.0, np.abs(ro["dP_ds"]).max())


def test_per_ray_frequency_and_mode(gpu_small, oracle_small, gl24, launcher):
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], 95e9,
                                             N_rings=2, min_azimuthal_points=3)
    n = len(w)
    f = np.where(np.arange(n) % 2 == 0, 95e9, 110e9)
    mode = np.where(np.arange(n) % 3 == 0, -1, 1)
    psi = np.linspace(0, 1, 120)
    res = tj.trace_bundle(gpu_small, pos, dirs, w, f, mode, 0.5, psi)
    ref = oracle_small.trace_bundle(pos, dirs, w, f, mode, 0.5, psi, gl24)
    assert np.array_equal(res["n_points"], ref["n_points"]) and np.abs(res["P_final"] - ref["P_final"]).max() < 1e-11
    assert l2rel(res["dP_dV"], ref["dP_dV"]) < L2_FAITHFUL


def test_batched_beams_equal_separate_calls(gpu_small, oracle_small, gl24, launcher):
    """make_beams: three launchers (angles, frequency, mode differ) in one device call == three make_beam-style calls."""
    psi = np.linspace(0, 1, 150)
    base = dict(r=2.5, phi=0.0, z=0.4, spot_size=launcher["spot"], inverse_curvature_radius=launcher["inv_Rc"])
    Ls = [dict(base, steering_angle_pol=np.deg2rad(30.0), steering_angle_tor=0.0, f=95e9, mode=1),
          dict(base, steering_angle_pol=np.deg2rad(24.0), steering_angle_tor=0.12, f=110e9, mode=1),
          dict(base, steering_angle_pol=np.deg2rad(33.0), steering_angle_tor=-0.08, f=95e9, mode=-1)]
    dP, dep, W, Pf, res = tj.make_beams(gpu_small, Ls, 0.6, psi)
    assert dP.shape == (3, 150) and dep.shape == (3,) and (res["status"] == 0).all()
    for b, L in enumerate(Ls):
        N0 = tj.pol_tor_angles_2_vector(L["steering_angle_pol"], L["steering_angle_tor"])
        pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], N0, L["spot_size"], L["inverse_curvature_radius"], L["f"])
        one = tj.trace_bundle(gpu_small, pos, dirs, w, L["f"], L["mode"], 0.6, psi)
        assert np.array_equal(one["P_final"], Pf[b])
        assert np.abs(one["dP_dV"] - dP[b]).max() <= 1e-12 * max(1e-300, np.abs(one["dP_dV"]).max())
        assert abs(one["deposited_power"] - dep[b]) < 1e-13
        ref = oracle_small.trace_bundle(pos, dirs, w, L["f"], L["mode"], 0.6, psi, gl24)
        assert abs(dep[b] - ref["deposited_power"]) <= FRAC_TOL * max(ref["deposited_power"], 1e-12)
        if ref["deposited_power"] > 1e-6:
            assert l2rel(dP[b], ref["dP_dV"]) < L2_FAITHFUL
    with pytest.raises(tj.TorjError, match="beam_id out of range"):
        tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.1, psi, beam_id=[3], n_beams=2)


def test_device_side_plasma_construction(arrays_small, oracle_small, gl24, launcher):
    """torj_plasma_create_from_data (prefilter + packing on the GPU, reference src/plasma.jl:16-58) against the
    host-prefiltered tables, on a non-uniform profile grid so the natural-cubic resampling is exercised too."""
    arr = dict(arrays_small)
    psi = np.linspace(0, 1, 101) ** 1.3
    arr["psi_prof"] = psi; arr["ne_prof"] = 3e19 * (1 - psi) + 1e17; arr["Te_prof"] = 4e3 * (1 - psi) ** 2 + 50
    host = tj.Plasma(*arr.values())
    dev = tj.Plasma(*arr.values(), build="device")
    rng = np.random.default_rng(5)
    R = rng.uniform(0.4, 2.8, 400); ph = rng.uniform(-3, 3, 400); Z = rng.uniform(-1.5, 1.5, 400)   # incl. off-grid points
    X = np.stack([R * np.cos(ph), R * np.sin(ph), Z], 1)
    N = rng.normal(size=X.shape) * 0.4
    a = host.probe(X, N, 95e9, 1); b = dev.probe(X, N, 95e9, 1)
    for k in ("psi", "Y", "N_par", "Lambda"):
        assert np.abs(a[k] - b[k]).max() <= 1e-12 * max(1.0, np.abs(a[k]).max()), k
    for k in ("ne", "Te", "X"):
        assert (np.abs(a[k] - b[k]) / np.maximum(np.abs(a[k]), 1e-300)).max() < 1e-10, k
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"])
    grid = np.linspace(0, 1, 100)
    ra = tj.trace_bundle(host, pos, dirs, w, launcher["f"], 1, 0.5, grid)
    rb = tj.trace_bundle(dev, pos, dirs, w, launcher["f"], 1, 0.5, grid)
    assert np.array_equal(ra["n_points"], rb["n_points"]) and np.abs(ra["P_final"] - rb["P_final"]).max() < 1e-10
    assert l2rel(rb["dP_dV"], ra["dP_dV"]) < 1e-9
    ref = O.OraclePlasma(*arr.values()).trace_bundle(pos, dirs, w, launcher["f"], 1, 0.5, grid, gl24)
    assert abs(rb["deposited_power"] - ref["deposited_power"]) <= FRAC_TOL * ref["deposited_power"]


def test_config5_high_te_third_harmonic_scan(gl24):
    """BASELINE.json configs[4] geometry with the Albajar model (the warm-plasma file of the reference is dead code,
    DESIGN.md §7): launchers at z = +-0.4, 110 and 170 GHz, T_e0 = 25 keV, one profile per beam in one call.
    170 GHz is absorbed at the THIRD harmonic (larger Bessel arguments, m = 3 dominant)."""
    tj.abs_Al_init(24)
    arr = tj.solovev_arrays(129, 129, Te0=25e3)
    pl = tj.Plasma(*arr.values()); opl = O.OraclePlasma(*arr.values())
    psi = np.linspace(0, 1, 200)
    Ls = [dict(r=2.5, phi=0.0, z=z0, steering_angle_pol=np.deg2rad(pol), steering_angle_tor=0.1, spot_size=0.0174,
               inverse_curvature_radius=1 / 3.99, f=f, mode=1)
          for f in (110e9, 170e9) for z0, pol in ((0.4, 30.0), (-0.4, -30.0))]
    dP, dep, W, Pf, res = tj.make_beams(pl, Ls, 1.0, psi, N_rings=2, min_azimuthal_points=3)
    assert (res["status"] == 0).all() and dP.shape == (4, 200)
    assert res["counters"]["n_harm"] > 0
    for b, L in enumerate(Ls):
        x0 = np.array([L["r"], 0.0, L["z"]])
        N0 = tj.pol_tor_angles_2_vector(L["steering_angle_pol"], L["steering_angle_tor"])
        pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, L["f"], N_rings=2, min_azimuthal_points=3)
        ref = opl.trace_bundle(pos, dirs, w, L["f"], 1, 1.0, psi, gl24)
        assert np.abs(Pf[b] - ref["P_final"]).max() < 1e-10
        assert abs(dep[b] - ref["deposited_power"]) <= FRAC_TOL * ref["deposited_power"]
        assert l2rel(dP[b], ref["dP_dV"]) < L2_FAITHFUL
    assert dep[2] > 0.99 and Pf[2].max() > 1e-5           # 170 GHz: strong but incomplete third-harmonic absorption


def test_randomised_rays_against_oracle(arrays_small, gl24):
    """Seeded random launchers (position, angles, frequency 60-180 GHz, X/O mode) on equilibria with scaled density and
    temperature: statuses, step counts, final power and end points must agree ray by ray, including the rays that fail
    to initialise (no bracket, cut-off)."""
    rng = np.random.default_rng(20261018)
    psi = np.linspace(0, 1, 60)
    n_checked = n_bad = 0
    for trial in range(3):
        arr = dict(arrays_small)
        arr["ne_prof"] = arr["ne_prof"] * rng.uniform(0.3, 2.5)
        arr["Te_prof"] = arr["Te_prof"] * rng.uniform(0.2, 4.0)
        pl = tj.Plasma(*arr.values()); opl = O.OraclePlasma(*arr.values())
        n = 48
        z0 = rng.uniform(-0.6, 0.6, n); R0 = rng.uniform(2.35, 2.9, n); ph = rng.uniform(-0.3, 0.3, n)
        pol = np.deg2rad(rng.uniform(-40, 40, n)) + np.arctan2(z0, R0 - 1.7) * 0.8
        tor = np.deg2rad(rng.uniform(-25, 25, n))
        pos = np.stack([R0 * np.cos(ph), R0 * np.sin(ph), z0], 1)
        dirs = np.stack([tj.pol_tor_angles_2_vector(p, t) for p, t in zip(pol, tor)])
        f = rng.choice([60e9, 82.7e9, 95e9, 110e9, 140e9, 170e9], n)
        mode = rng.choice([1, -1], n)
        w = np.full(n, 1.0 / n)
        res = tj.trace_bundle(pl, pos, dirs, w, f, mode, 0.8, psi, options=tj.default_options(n_segments=40),
                              trajectories=(0, n), traj_max_pts=2 + 40 * 210)
        ref = opl.trace_bundle(pos, dirs, w, f, mode, 0.8, psi, gl24, opts=O.OracleOptions.default(n_segments=40),
                               deposition="streaming")
        assert np.array_equal(res["status"], ref["status"]), (res["status"], ref["status"])
        ok = ref["status"] == 0
        n_checked += int(ok.sum()); n_bad += int((~ok).sum())
        assert np.array_equal(res["n_points"][ok], ref["n_points"][ok])
        assert np.abs(res["P_final"][ok] - ref["P_final"][ok]).max() < 1e-9
        if ref["deposited_power"] > 1e-9:
            assert abs(res["deposited_power"] - ref["deposited_power"]) <= FRAC_TOL * ref["deposited_power"] + 1e-12
            assert l2rel(res["dP_dV"], ref["dP_dV"]) < 1e-6
        for i in np.nonzero(ok)[0][:6]:
            ro = opl.make_ray(pos[i], dirs[i], f[i], int(mode[i]), 0.8, psi, gl24, opts=O.OracleOptions.default(n_segments=40))
            m = int(res["n_points"][i])
            assert m == len(ro["s"])
            end = res["traj_xyz"][i, :, m - 1]
            assert max(abs(end[0] - ro["x"][-1]), abs(end[1] - ro["y"][-1]), abs(end[2] - ro["z"][-1])) < TRAJ_TOL
    assert n_checked > 60 and n_bad > 0      # the sample must contain both traced and rejected rays


def test_harmonics_above_the_third(gl24):
    """torj_options.max_harmonic > 3: the harmonics the reference leaves out
    (src/absorption.jl:199,213), same integral, against the oracle's loop over m = 2..max_harmonic. 225 GHz in the
    25 keV Solov'ev plasma is absorbed at the fourth and fifth harmonic only."""
    arr = tj.solovev_arrays(65, 65, Te0=25e3, ne0=6e19)
    pl = tj.Plasma(*arr.values()); opl = O.OraclePlasma(*arr.values())
    psi = np.linspace(0, 1, 100)
    x0 = np.array([2.5, 0.0, 0.4]); N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.1)
    pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, 225e9, N_rings=2, min_azimuthal_points=3)
    absorbed = {}
    for mh in (1, 3, 4, 5):
        res = tj.trace_bundle(pl, pos, dirs, w, 225e9, 1, 1.0, psi, options=tj.default_options(max_harmonic=mh, n_segments=50))
        ref = opl.trace_bundle(pos, dirs, w, 225e9, 1, 1.0, psi, gl24, opts=O.OracleOptions.default(max_harmonic=mh, n_segments=50),
                               deposition="streaming")
        assert (res["status"] == 0).all() and np.array_equal(res["n_points"], ref["n_points"])
        assert np.abs(res["P_final"] - ref["P_final"]).max() < 1e-9
        absorbed[mh] = res["deposited_power"]
        if mh <= 3:
            assert res["deposited_power"] == 0.0 and res["counters"]["n_harm"] == 0
        else:
            assert abs(res["deposited_power"] - ref["deposited_power"]) <= FRAC_TOL * ref["deposited_power"]
            assert l2rel(res["dP_dV"], ref["dP_dV"]) < 1e-6
    assert 0.5 < absorbed[4] < absorbed[5] < 1.0 + 1e-12


def test_segment_handoff_schedule_gives_the_same_rays(gpu_small, arrays_small, launcher):
    """torj_options.schedule: 2 = a ray changes lanes (and SMs) between segments, its state travelling through global
    memory; 1 = a lane keeps its ray. Per-ray results must be IDENTICAL (same arithmetic in the same order); the
    profile differs only by the order of the atomic adds. Covers failing rays, a trajectory window and two beams."""
    L = launcher
    pos, dirs, w = tj.launch_peripheral_rays(L["x0"], L["N0"], L["spot"], L["inv_Rc"], L["f"],
                                             N_rings=4, min_azimuthal_points=7)
    n = len(w)
    pos = pos.copy(); pos[3] = [9.0, 0.0, 0.0]; pos[7] = [2.5, 0.0, 3.0]       # two rays that never reach the plasma
    f = np.where(np.arange(n) % 2 == 0, 95e9, 110e9); mode = np.where(np.arange(n) % 3 == 0, -1, 1)
    beam = (np.arange(n) % 2).astype(np.int32)
    psi = np.linspace(0, 1, 80)
    out = {}
    for sched in (1, 2):
        out[sched] = tj.trace_bundle(gpu_small, pos, dirs, w, f, mode, 0.6, psi, options=tj.default_options(schedule=sched, n_segments=30),
                                     trajectories=(0, n), traj_max_pts=2 + 30 * 210, beam_id=beam, n_beams=2)
    a, b = out[1], out[2]
    assert np.array_equal(a["status"], b["status"]) and (a["status"] != 0).sum() == 2
    ok = a["status"] == 0
    assert np.array_equal(a["n_points"], b["n_points"])
    assert np.array_equal(a["P_final"][ok], b["P_final"][ok])
    assert np.array_equal(a["P_deposited_ray"], b["P_deposited_ray"])
    for k in ("traj_s", "traj_xyz", "traj_P", "traj_dP_ds", "traj_dP_dV_ray"):
        assert np.array_equal(a[k], b[k]), k
    assert a["dP_dV"].shape == (2, 80) and l2rel(b["dP_dV"].ravel(), a["dP_dV"].ravel()) < 1e-13
    assert np.allclose(a["deposited_power"], b["deposited_power"], rtol=1e-13, atol=0)
    assert a["counters"]["n_acc"] == b["counters"]["n_acc"] and a["counters"]["n_rhs"] == b["counters"]["n_rhs"]


def test_option_validation(gpu_small, launcher):
    psi = np.linspace(0, 1, 10)
    for bad in (dict(max_harmonic=0), dict(max_harmonic=17), dict(scheme=2), dict(n_segments=0), dict(dtmax=0.0),
                dict(abstol=-1.0), dict(alpha_floor=-1.0), dict(max_steps_per_segment=0), dict(schedule=4)):
        with pytest.raises(tj.TorjError):
            tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.2, psi,
                            options=tj.default_options(**bad))
    with pytest.raises(tj.TorjError):
        tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.0, psi)
    with pytest.raises(tj.TorjError):
        tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.2, psi[::-1].copy())


def test_smallest_inputs(gpu_small, launcher):
    r = tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.2, np.array([0.0, 1.0]))
    assert r["status"][0] == 0 and r["dP_dV"].shape == (2,) and r["dP_dV"][1] == 0.0
    with pytest.raises(tj.TorjError, match="strictly increasing"):
        tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.2, np.array([0.0, 0.5, 0.5]))
    with pytest.raises(tj.TorjError):
        tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.2, np.array([0.5]))
    with pytest.raises(tj.TorjError, match="scheme"):
        tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.2, PSI,
                        options=tj.default_options(scheme=7))


def test_trajectory_window_truncation_keeps_physics(gpu_small, launcher):
    psi = np.linspace(0, 1, 100)
    a = tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.3, psi, trajectories=(0, 1))
    b = tj.trace_bundle(gpu_small, launcher["x0"][None], launcher["N0"][None], [1.0], 95e9, 1, 0.3, psi, trajectories=(0, 1),
                        traj_max_pts=100)
    assert a["status"][0] == 0 and b["status"][0] == 6
    assert a["n_points"][0] == b["n_points"][0] and a["P_final"][0] == b["P_final"][0]
    assert np.array_equal(a["traj_s"][0, :100], b["traj_s"][0, :100])


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json config 3 size (65 543 rays): size-independent properties
# ------------------------------------------------------------------------------------------------------------------
def test_config3_full_size_properties(gpu_full, oracle_full, gl24, launcher):
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"],
                                             N_rings=66, min_azimuthal_points=14)
    n = len(w)
    assert n == 65543
    res = tj.trace_bundle(gpu_full, pos, dirs, w, launcher["f"], 1, 1.0, PSI)
    assert (res["status"] == 0).all() and res["counters"]["n_rays_ok"] == n
    assert res["counters"]["n_acc"] == int(np.sum(res["n_points"] - 2))
    # reference test/tests/test_make_beam.jl:14-31 identities
    absorbed = 1.0 - float(np.sum(w * res["P_final"]))
    assert abs(res["deposited_power"] - absorbed) < 1e-3
    assert abs(float(np.sum(w * res["P_deposited_ray"])) - res["deposited_power"]) < 1e-12
    dpsi = PSI[1] - PSI[0]
    dVdpsi = (gpu_full.volume(PSI + 1e-6) - gpu_full.volume(PSI - 1e-6)) / 2e-6
    assert abs(float(np.sum(dVdpsi * res["dP_dV"] * dpsi)) - res["deposited_power"]) < 1e-3
    # linearity over shards: two halves (different refill order, different atomics order) add up to the whole
    h = n // 2
    a = tj.trace_bundle(gpu_full, pos[:h], dirs[:h], w[:h], launcher["f"], 1, 1.0, PSI)
    b = tj.trace_bundle(gpu_full, pos[h:], dirs[h:], w[h:], launcher["f"], 1, 1.0, PSI)
    assert np.array_equal(np.concatenate([a["P_final"], b["P_final"]]), res["P_final"])
    assert np.abs(a["dP_dV"] + b["dP_dV"] - res["dP_dV"]).max() <= 1e-12 * res["dP_dV"].max()
    # spot check of 24 rays spread over the bundle against the oracle
    pick = np.linspace(0, n - 1, 24).astype(int)
    ref = oracle_full.trace_bundle(pos[pick], dirs[pick], w[pick], launcher["f"], 1, 1.0, PSI, gl24, deposition="streaming")
    assert np.array_equal(res["n_points"][pick], ref["n_points"])
    assert np.abs(res["P_final"][pick] - ref["P_final"]).max() < 1e-12
    sub = tj.trace_bundle(gpu_full, pos[pick], dirs[pick], w[pick], launcher["f"], 1, 1.0, PSI)
    assert l2rel(sub["dP_dV"], ref["dP_dV"]) < L2_LIKE


def test_config4_angle_sweep_properties(gpu_full, oracle_full, gl24):
    """BASELINE.json configs[3] geometry at reduced count: 8 x 8 launcher angles x 1 025-ray beams = 65 600 rays in one call,
    one profile per beam; per-beam identities of reference test/tests/test_make_beam.jl:14-21 and an oracle spot check."""
    psi = np.linspace(0, 1, 400)
    Ls = []
    for pol in np.deg2rad(np.linspace(10.0, 40.0, 8)):
        for tor in np.deg2rad(np.linspace(-15.0, 15.0, 8)):
            Ls.append(dict(r=2.5, phi=0.0, z=0.4, steering_angle_pol=pol, steering_angle_tor=tor, spot_size=0.0174,
                           inverse_curvature_radius=1 / 3.99, f=95e9, mode=1))
    dP, dep, W, Pf, res = tj.make_beams(gpu_full, Ls, 1.0, psi, N_rings=7, min_azimuthal_points=20)
    assert dP.shape == (64, 400) and len(res["status"]) == 64 * 1025
    ok = res["status"] == 0
    assert ok.mean() > 0.999
    off = np.cumsum([0] + [len(w) for w in W])
    for b in range(64):
        sl = slice(off[b], off[b + 1])
        okb = ok[sl]
        absorbed = float(np.sum(W[b][okb] * (1.0 - Pf[b][okb])))
        assert abs(dep[b] - absorbed) < 1e-3                       # profile-integrated power == 1 - sum w P_end
        assert abs(float(np.sum(W[b][okb] * res["P_deposited_ray"][sl][okb])) - dep[b]) < 1e-12
    # the whole-bundle counters are the sums of the per-ray ones
    assert res["counters"]["n_acc"] == int(np.sum(np.maximum(res["n_points"] - 2, 0)))
    # oracle spot check: three beams, 12 rays each
    for b in (0, 27, 63):
        L = Ls[b]
        N0 = tj.pol_tor_angles_2_vector(L["steering_angle_pol"], L["steering_angle_tor"])
        pos, dirs, w = tj.launch_peripheral_rays(np.array([2.5, 0.0, 0.4]), N0, 0.0174, 1 / 3.99, 95e9, N_rings=7, min_azimuthal_points=20)
        pick = np.linspace(0, len(w) - 1, 12).astype(int)
        ref = oracle_full.trace_bundle(pos[pick], dirs[pick], w[pick], 95e9, 1, 1.0, psi, gl24, deposition="streaming")
        good = ref["status"] == 0
        assert np.array_equal(res["status"][off[b] + pick] == 0, good)
        assert np.array_equal(res["n_points"][off[b] + pick][good], ref["n_points"][good])
        assert np.abs(Pf[b][pick][good] - ref["P_final"][good]).max() < 1e-11


def test_device_side_ray_generation(gpu_small, launcher):
    """torj_bundle_create_from_launchers == launch_peripheral_rays (reference src/launch.jl:24-132) per launcher,
    divergent / convergent / paraxial beams, with and without weight normalisation."""
    import ctypes as C
    from torj_jl_b200 import _lib
    base = dict(r=2.5, phi=0.0, z=0.4, spot_size=launcher["spot"], f=95e9, mode=1)
    Ls = [dict(base, steering_angle_pol=np.deg2rad(30.0), steering_angle_tor=0.0, inverse_curvature_radius=launcher["inv_Rc"]),
          dict(base, steering_angle_pol=np.deg2rad(22.0), steering_angle_tor=0.2, inverse_curvature_radius=-1 / 2.5, f=110e9, mode=-1),
          dict(base, steering_angle_pol=np.deg2rad(35.0), steering_angle_tor=-0.1, inverse_curvature_radius=np.inf, spot_size=0.03, phi=0.3)]
    psi = np.linspace(0, 1, 90)
    for norm in (True, False):
        kw = dict(N_rings=4, min_azimuthal_points=7, normalize_weight_sum=norm)
        a = tj.make_beams(gpu_small, Ls, 0.4, psi, **kw)
        b = tj.make_beams(gpu_small, Ls, 0.4, psi, device_launch=True, **kw)
        for q in range(3):
            assert np.abs(a[2][q] - b[2][q]).max() <= 1e-14 * a[2][q].max()            # ray weights
            assert np.array_equal(a[4]["n_points"], b[4]["n_points"])
            assert np.abs(a[3][q] - b[3][q]).max() < 1e-10                             # P_final per ray
            assert np.abs(a[0][q] - b[0][q]).max() <= 1e-9 * max(np.abs(a[0][q]).max(), 1e-300)
        assert np.abs(a[1] - b[1]).max() < 1e-10
    with pytest.raises((ValueError, tj.TorjError)):
        tj.make_beams(gpu_small, Ls, 0.4, psi, device_launch=True, N_rings=1)


def test_single_process_multi_gpu_front_end(gpu_small, arrays_small, launcher):
    """torj_multi_trace: shards over all visible GPUs in one process (1 GPU here at round end, N under gpurun --gpus N);
    the host-side ordered sum reproduces the single-device result."""
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"],
                                             N_rings=4, min_azimuthal_points=6)
    psi = np.linspace(0, 1, 120)
    bid = (np.arange(len(w)) % 3).astype(np.int32)
    one = tj.trace_bundle(gpu_small, pos, dirs, w, launcher["f"], 1, 0.5, psi, beam_id=bid, n_beams=3)
    mg = tj.MultiGPU()
    try:
        mg.abs_Al_init(24)
        r = mg.trace_bundle(gpu_small, pos, dirs, w, launcher["f"], 1, 0.5, psi, beam_id=bid, n_beams=3)
        assert mg.n_devices >= 1
        assert np.array_equal(r["n_points"], one["n_points"]) and np.array_equal(r["status"], one["status"])
        assert np.abs(r["P_final"] - one["P_final"]).max() < 1e-11          # tables built by the device-side constructor
        assert np.abs(r["dP_dV"] - one["dP_dV"]).max() <= 1e-9 * np.abs(one["dP_dV"]).max()
        assert np.abs(r["deposited_power"] - one["deposited_power"]).max() < 1e-11
        assert r["counters"]["n_acc"] == one["counters"]["n_acc"]
        again = mg.trace_bundle(gpu_small, pos, dirs, w, launcher["f"], 1, 0.5, psi, beam_id=bid, n_beams=3)
        if mg.n_devices == 1:
            pass  # atomics order inside one device is not fixed
        assert np.abs(again["dP_dV"] - r["dP_dV"]).max() <= 1e-12 * np.abs(r["dP_dV"]).max()
    finally:
        mg.close()
        tj.abs_Al_init(24)


def test_measurement_helpers():
    import ctypes as C
    from torj_jl_b200 import _lib
    ctx = _lib.context()
    lat, ms = C.c_double(), C.c_double()
    _lib.check(tj.lib().torj_fp64_latency(ctx, 2000, C.byref(lat)))
    assert 4.0 < lat.value < 16.0                                   # measured 8.2 cycles on B200
    _lib.check(tj.lib().torj_ctx_last_trace_ms(ctx, C.byref(ms)))    # earlier tests in this module have traced
    assert ms.value > 0.0


def test_fp64_peak_probe_and_launch_counter():
    import ctypes as C
    from torj_jl_b200 import _lib
    ctx = _lib.context()
    n0 = tj.lib().torj_ctx_launch_count(ctx)
    tf, ms = C.c_double(), C.c_double()
    _lib.check(tj.lib().torj_fp64_peak(ctx, 4000, C.byref(tf), C.byref(ms)))
    assert 20.0 < tf.value < 45.0                                  # nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2
    assert tj.lib().torj_ctx_launch_count(ctx) == n0 + 2


def test_device_elementary_functions_against_libm():
    """rcp_fast / rsqrt_fast / sqrt_fast / exp_fast (MUFU seed + one third-order FMA correction; branch-free exp) against
    numpy in units of the last place, over the magnitudes the path produces (1e-30 .. 1e30; exponents in +-700)."""
    import ctypes as C
    from torj_jl_b200 import _lib
    rng = np.random.default_rng(7)
    x = np.concatenate([10.0 ** rng.uniform(-30, 30, 200000), rng.uniform(0.5, 2.0, 100000),
                        np.nextafter(2.0 ** np.arange(-40, 40), np.inf), 2.0 ** np.arange(-40, 40)])
    out = np.empty((4, x.size))
    _lib.check(tj.lib().torj_math_probe(_lib.context(), x.size, x.ctypes.data_as(_lib.c_dp), out.ctypes.data_as(_lib.c_dp)))
    ld = np.longdouble   # x87 extended precision: 11 more bits than double
    assert np.finfo(ld).nmant >= 63
    ulp = lambda got, ref: float(np.max(np.abs(got.astype(ld) - ref) / np.spacing(np.abs(ref.astype(np.float64))).astype(ld)))
    xl = x.astype(ld)
    assert ulp(out[0], 1 / xl) <= 1.0
    assert ulp(out[1], 1 / np.sqrt(xl)) <= 1.5
    assert ulp(out[2], np.sqrt(xl)) <= 1.0
    e = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-1, 1, 100000), [0.0, -708.0, 708.0, -800.0]])
    out = np.empty((4, e.size))
    _lib.check(tj.lib().torj_math_probe(_lib.context(), e.size, e.ctypes.data_as(_lib.c_dp), out.ctypes.data_as(_lib.c_dp)))
    assert ulp(out[3], np.exp(np.clip(e, -708.0, 708.0).astype(ld))) <= 1.5

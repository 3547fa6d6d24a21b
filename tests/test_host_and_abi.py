"""Host logic of the reference-facing mirror and the C-ABI surface.  CPU only (no compute calls on the library)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
from torj_jl_b200.distributed import shard_range
from oracle import torj_oracle as O


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "torj_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(torj_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/torj_cuda.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert tj.lib().torj_abi_version() == 3 == int(re.search(r"#define TORJ_ABI_VERSION (\d+)", hdr).group(1))


def test_default_options_are_the_reference_constants():
    o = tj.default_options()
    assert (o.scheme, o.n_segments, o.max_harmonic) == (0, 100, 3)           # src/solve.jl:145, src/absorption.jl:199
    assert (o.dtmax, o.abstol, o.reltol) == (1e-4, 1e-6, 1e-6)               # src/solve.jl:157
    assert (o.psi_stop, o.p_stop, o.te_min) == (1.0, 1e-6, 20.0)             # src/solve.jl:174,176; src/absorption.jl:194
    assert o.alpha_floor == 1e-14
    assert (o.schedule, o.absorption_model, o.lanes_per_ray, o.reserved_) == (0, 0, 0, 0) and C.sizeof(o) == 88
    with pytest.raises(AttributeError):
        tj.default_options(nonexistent=1)


@pytest.mark.skipif(has_gpu(), reason="checks the loud failure on a box without CUDA")
def test_no_cpu_fallback():
    with pytest.raises(tj.TorjError, match="no CUDA device"):
        tj.abs_Al_init(24)


def test_plasma_tables_match_oracle(arrays_small, oracle_small):
    pl = tj.Plasma(*arrays_small.values())
    for k in ("psi", "lnne", "lnTe", "BR", "BZ", "Bphi"):
        ref = oracle_small.coefs(k)
        assert pl.coefs[k].shape == ref.shape == (67, 67)
        assert np.abs(pl.coefs[k] - ref).max() <= 1e-13 * np.abs(ref).max()
    vc, x0, dx = oracle_small.volume_coefs()
    assert np.abs(vc - pl.vol_coef).max() < 1e-13 and (x0, dx) == (pl.vol_psi0, pl.vol_dpsi)
    assert pl.psi_prof_max == oracle_small.psi_prof_max == 1.0
    psi = np.array([-0.1, 0.0, 0.37, 1.0, 1.2])
    assert np.abs(pl.volume(psi) - oracle_small.volume(psi)).max() < 1e-13
    with pytest.raises(ValueError):
        bad = dict(arrays_small); bad["psi_norm_data"] = bad["psi_norm_data"][:, :-1]
        tj.Plasma(*bad.values())


def test_nonuniform_profile_grid_is_resampled(arrays_small):
    """make_2d_prof_spline resamples the 1-D profile on a uniform psi range (reference src/plasma.jl:17-18)."""
    arr = dict(arrays_small)
    psi = np.linspace(0, 1, 101) ** 1.5
    arr["psi_prof"] = psi
    arr["ne_prof"] = 3e19 * (1 - psi) + 1e17
    arr["Te_prof"] = 4e3 * (1 - psi) ** 2 + 50
    pl = tj.Plasma(*arr.values())
    opl = O.OraclePlasma(*arr.values())
    for k in ("lnne", "lnTe"):
        ref = opl.coefs(k)
        assert np.abs(pl.coefs[k] - ref).max() <= 1e-12 * np.abs(ref).max()


def test_launch_matches_oracle_and_reference_counts(launcher):
    for kw in (dict(), dict(N_rings=7, min_azimuthal_points=20), dict(N_rings=21, min_azimuthal_points=11, normalize_weight_sum=False)):
        for tor in (0.0, 0.2):
            N0 = tj.pol_tor_angles_2_vector(np.deg2rad(25.0), tor)
            a = tj.launch_peripheral_rays(launcher["x0"], N0, launcher["spot"], launcher["inv_Rc"], launcher["f"], **kw)
            b = O.launch_peripheral_rays(launcher["x0"], N0, launcher["spot"], launcher["inv_Rc"], launcher["f"], **kw)
            assert a[0].shape == b[0].shape
            for x, y in zip(a, b):
                assert np.abs(x - y).max() < 1e-13
    pos, dirs, w = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], launcher["spot"], launcher["inv_Rc"], launcher["f"])
    assert pos.shape == (46, 3) and abs(w.sum() - 1) < 1e-14           # default bundle: N_theta = 5, 15, 26
    # convergent beam and paraxial beam branches (src/launch.jl:102-117)
    pc = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], 0.03, -1 / 2.0, 95e9)
    oc = O.launch_peripheral_rays(launcher["x0"], launcher["N0"], 0.03, -1 / 2.0, 95e9)
    assert np.abs(pc[1] - oc[1]).max() < 1e-13
    pp = tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], 0.03, np.inf, 95e9)
    assert np.allclose(pp[1], launcher["N0"] / np.linalg.norm(launcher["N0"]))
    with pytest.raises(ValueError, match="N_rings = 1 < 2"):
        tj.launch_peripheral_rays(launcher["x0"], launcher["N0"], 0.03, 0.25, 95e9, N_rings=1)


def test_weights_integrate_a_gaussian():
    """reference test/tests/test_launch_weights.jl:27-50."""
    p, d, w = tj.launch_peripheral_rays([0, 0, 0.0], [0, 0, 1.0], 0.0174, 1 / 3.99, 92.5e9, N_rings=21, min_azimuthal_points=11,
                                        normalize_weight_sum=False)
    assert len(w) == 5165 and abs(w.sum() - 1.0) / 1.0 < 0.01


def test_pol_tor_vector():
    v = tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.0)
    assert np.allclose(v, [-np.sqrt(3) / 2, 0.0, -0.5]) and abs(np.linalg.norm(v) - 1) < 1e-15


def test_shard_range_partitions():
    for n in (1, 46, 1025, 65543, 1049600):
        for ws in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_prefilter_helpers_match_scipy():
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(3)
    y = rng.normal(size=37)
    c = np.empty(39)
    _lib.check(tj.lib().torj_bspline_prefilter_1d(37, y.ctypes.data_as(_lib.c_dp), c.ctypes.data_as(_lib.c_dp)))
    x = np.arange(37.0)
    cs = CubicSpline(x, y, bc_type="natural")
    xq = np.linspace(0, 36, 400)
    from torj_jl_b200.plasma import _eval_1d_line
    assert np.abs(_eval_1d_line(c, 0.0, 1.0, 37, xq) - cs(xq)).max() < 1e-13
    for n in (2, 3, 4):
        yy = rng.normal(size=n); cc = np.empty(n + 2)
        _lib.check(tj.lib().torj_bspline_prefilter_1d(n, yy.ctypes.data_as(_lib.c_dp), cc.ctypes.data_as(_lib.c_dp)))
        assert np.abs(_eval_1d_line(cc, 0.0, 1.0, n, np.arange(n * 1.0)) - yy).max() < 1e-14


def test_block_cyclic_sharding_partitions_the_rays():
    from torj_jl_b200.distributed import shard_block_cyclic
    for n, block, world in ((1049600, 1025, 8), (10, 3, 4), (7, 7, 2), (100, 1, 3)):
        parts = [shard_block_cyclic(n, block, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(n))
        assert all(np.all(np.diff(p) > 0) for p in parts if len(p) > 1)
    p = shard_block_cyclic(1049600, 1025, 3, 8)
    assert p[0] == 3 * 1025 and p[1025] == 11 * 1025 and len(p) == 128 * 1025


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the package, its CUDA sources or the built library may import, include
    or link it, and importing the package must not load it."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "torj_jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                pat = (r"^\s*(from|import)\s+[^#\n]*oracle|ctypes[^\n]*oracle|CDLL[^\n]*oracle" if f.endswith(".py")
                       else r"^\s*#\s*include[^\n]*oracle" if not f.endswith("Makefile") else r"^[^#\n]*oracle")
                assert not re.search(pat, txt, flags=re.M | re.I), f"{os.path.join(dirpath, f)} uses the oracle"
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    code = "import sys; sys.path.insert(0, %r); import torj_jl_b200; print(any('oracle' in m for m in sys.modules))" % ROOT
    assert subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout.strip() == "False"


def test_config5_scan_sharding_gives_every_rank_a_share_of_every_beam():
    """bench.py deals the rays of a multi-beam scan to the ranks in C5_BLOCK-ray blocks: a partition of the bundle, the same
    number of rays per rank to within one block, and every rank holds rays of every 1 025-ray beam (whole beams per rank left
    ranks with cheap beams idle: profiles/README.md, config 5)."""
    import bench
    from torj_jl_b200.distributed import shard_indices

    n, world = 16 * 1025, 8
    beam = np.arange(n) // 1025
    parts = [shard_indices(n, r, world, "block_cyclic", bench.C5_BLOCK) for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= bench.C5_BLOCK
    for p in parts:
        per_beam = np.bincount(beam[p], minlength=16)
        assert per_beam.min() >= 1025 // world - bench.C5_BLOCK and per_beam.max() <= 1025 // world + 2 * bench.C5_BLOCK


def test_exp_table_and_constants_of_the_kernel_give_1p5_ulp():
    """exp_fast (csrc/torj_device.cuh) emulated with the table and constants parsed from the source: n = rint(64 x / ln 2) by the
    1.5 * 2^52 addition, r = x - n ln2/64 in two FMAs, 2^(j/64) from the table, degree-5 expm1, t + t q. Worst error over
    2 M arguments in [-708, 708] below the 1.5 ulp the GPU test allows (long double as the reference, FMA emulated in long double)."""
    import os
    import re

    src = open(os.path.join(os.path.dirname(__file__), "..", "torj_jl_b200", "csrc", "torj_device.cuh")).read()
    tab = np.array([float(v) for v in re.search(r"d_exp_tab\[64\] = \{(.*?)\};", src, re.S).group(1).replace("\n", " ").split(",") if v.strip()])
    cst = [eval(v, {"__builtins__": {}}) for v in re.search(r"c_expt\[6\] = \{(.*?)\};", src, re.S).group(1).split(",")]
    assert len(tab) == 64 and len(cst) == 6
    ld = np.longdouble
    assert np.abs(tab - np.array([float(ld(2) ** (ld(j) / 64)) for j in range(64)])).max() == 0.0   # correctly rounded 2^(j/64)
    fma = lambda a, b, c: (ld(a) * ld(b) + ld(c)).astype(np.float64)
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-708, 708, 1_000_000), rng.uniform(-40, 5, 700_000), rng.uniform(-1, 1, 300_000), [0.0, -708.0, 708.0]])
    magic = 6755399441055744.0
    t = (x * cst[0]) + magic
    nd = t - magic
    n = nd.astype(np.int64)
    r = fma(nd, cst[2], fma(nd, cst[1], x))
    assert np.abs(r).max() <= np.log(2) / 128 * (1 + 1e-9)
    p = (r * cst[3]) + cst[4]
    p = fma(p, r, cst[5])
    p = fma(p, r, 0.5)
    q = fma(p * r, r, r)
    v = fma(tab[n & 63], q, tab[n & 63])
    got = np.ldexp(v, n >> 6)
    ref = np.exp(x.astype(ld))
    ulp = np.abs(got.astype(ld) - ref) / np.spacing(np.abs(ref.astype(np.float64))).astype(ld)
    assert float(ulp.max()) < 1.5
    assert v.min() > 0.99 and v.max() < 2.0 and (n >> 6).min() >= -1022 and (n >> 6).max() <= 1022   # exponent add stays normal

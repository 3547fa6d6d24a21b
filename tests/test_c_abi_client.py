"""The C ABI used from plain C (tests/c_abi/abi_smoke.c): compiles against include/torj_cuda.h with gcc, links
libtorj_cuda.so, no Python or torch in the process.  CPU: builds + fails loudly without a device.  GPU: its numbers
equal the Python mirror's on the same inputs."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu

SRC = os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c")
LIBDIR = os.path.join(ROOT, "torj_jl_b200")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", exe, SRC, "-I" + os.path.join(ROOT, "include"), "-L" + LIBDIR,
                           "-ltorj_cuda", "-lm", "-Wl,-rpath," + LIBDIR])
    return exe


@pytest.mark.skipif(has_gpu(), reason="checks the loud failure on a box without CUDA")
def test_c_client_builds_and_fails_loudly_without_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_c_client_matches_python_mirror(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.strip().splitlines()}
    assert out["status"] == ["0", "2", "0"]                      # second ray points away from the plasma
    import torj_jl_b200 as tj
    t, w = np.polynomial.legendre.leggauss(8)
    tj.abs_Al_init(8)
    arr = tj.solovev_arrays(65, 65, nprof=51)
    pl = tj.Plasma(*arr.values(), build="device")
    pos = np.array([[2.5, 0.0, 0.4], [2.5, 0.0, 0.4], [2.5, 0.01, 0.38]])
    c = np.sqrt(3) / 2
    dirs = np.array([[-0.8660254037844387, 0, -0.5], [0.8660254037844387, 0, 0.5], [-0.8660254037844387, 0, -0.5]])
    res = tj.trace_bundle(pl, pos, dirs, [0.5, 0.25, 0.25], 95e9, 1, 0.5, np.linspace(0, 1, 40), trajectories=(0, 1),
                          traj_max_pts=2 + 100 * 58)
    tj.abs_Al_init(24)
    assert [int(v) for v in out["npts"]] == list(res["n_points"])
    assert np.allclose([float(v) for v in out["P_final"]], res["P_final"], rtol=0, atol=1e-10)
    assert abs(float(out["deposited"][0]) - res["deposited_power"]) < 1e-10
    assert int(out["n_acc"][0]) == res["counters"]["n_acc"]
    n0 = int(res["n_points"][0])
    assert abs(float(out["traj0"][1]) - res["traj_s"][0, n0 - 1]) < 1e-10
    assert abs(float(out["traj0"][3]) - res["traj_xyz"][0, 0, n0 - 1]) < 1e-10
    assert abs(float(out["profile_sum"][0]) - res["dP_dV"].sum()) <= 1e-9 * abs(res["dP_dV"].sum())

"""IMAS dd ingestion (reference test/tests/setup.jl:31-62): host logic only, runs on CPU."""
import json

import numpy as np
import pytest

from torj_jl_b200 import imas
from torj_jl_b200.synthetic import solovev_arrays

KEYS2D = ("psi_norm_data", "Br_data", "Bz_data", "Bphi_data")


def _same(a, b):
    for k in b:
        tol = 4e-16 if k in ("psi_norm_data", "psi_prof", "eqt1d_psi_norm") else 0.0  # one subtraction + one division
        assert a[k].shape == b[k].shape, k
        assert np.max(np.abs(a[k] - b[k])) <= tol * max(1.0, np.max(np.abs(b[k]))), k


def test_round_trip_through_json(tmp_path):
    arr = solovev_arrays(33, 41)  # non-square: the 2-D orientation is checked by shape
    dd = imas.solovev_dd(arr, times=(2.0,))
    f = tmp_path / "dd.json"
    f.write_text(json.dumps(dd))
    got = imas.plasma_arrays_from_dd(imas.load_imas_json(str(f)), 2.0)
    _same(got, arr)
    # transposed 2-D layout on a non-square grid is recognised by its shape
    p2d = dd["equilibrium"]["time_slice"][0]["profiles_2d"][0]
    for k in ("psi", "b_field_r", "b_field_z", "b_field_tor"):
        p2d[k] = np.asarray(p2d[k]).T.tolist()
    _same(imas.plasma_arrays_from_dd(dd, 2.0), arr)


def test_square_grid_orientation_flag():
    arr = solovev_arrays(17, 17)
    dd = imas.solovev_dd(arr)
    _same(imas.plasma_arrays_from_dd(dd, 2.0), arr)
    p2d = dd["equilibrium"]["time_slice"][0]["profiles_2d"][0]
    for k in ("psi", "b_field_r", "b_field_z", "b_field_tor"):
        p2d[k] = np.asarray(p2d[k]).T.tolist()
    _same(imas.plasma_arrays_from_dd(dd, 2.0, matrix_order="dim2_outer"), arr)
    with pytest.raises(ValueError):
        imas.plasma_arrays_from_dd(dd, 2.0, matrix_order="columns")


def test_causal_time_slice_and_density_scale():
    arr = solovev_arrays(17, 19)
    dd = imas.solovev_dd(arr, times=(1.0, 2.0, 3.0), ne_scale_per_slice=(1.0, 2.0, 3.0))
    for t, sc in ((1.0, 1.0), (1.99, 1.0), (2.0, 2.0), (2.5, 2.0), (9.0, 3.0)):
        got = imas.plasma_arrays_from_dd(dd, t)
        assert np.array_equal(got["ne_prof"], arr["ne_prof"] * sc)
    with pytest.raises(ValueError):
        imas.plasma_arrays_from_dd(dd, 0.5)
    with pytest.raises(ValueError):
        imas.plasma_arrays_from_dd(dd)  # three slices and no global_time
    low = imas.plasma_arrays_from_dd(dd, 2.0, density_scale=0.3)  # plasma_low_density, setup.jl:57-58
    assert np.array_equal(low["ne_prof"], arr["ne_prof"] * 2.0 * 0.3)
    assert imas.causal_index([0.0, 1.0, 2.0], 2.0) == 2


def test_plasma_from_dd_matches_direct_construction():
    import torj_jl_b200 as tj
    arr = solovev_arrays(33, 33)
    pl_dd = tj.plasma_from_dd(tj.solovev_dd(arr), 2.0)
    pl = tj.Plasma(*[arr[k] for k in ("R_coords", "Z_coords", "psi_norm_data", "psi_prof", "ne_prof", "Te_prof", "Br_data",
                                      "Bz_data", "Bphi_data", "eqt1d_psi_norm", "eqt1d_volume")])
    for k in pl.coefs:
        assert np.max(np.abs(pl_dd.coefs[k] - pl.coefs[k])) <= 1e-13 * np.max(np.abs(pl.coefs[k])), k
    assert abs(pl_dd.psi_prof_max - pl.psi_prof_max) < 1e-15

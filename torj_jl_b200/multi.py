"""Single-process multi-GPU front end (torj_multi_*): one call shards a bundle by contiguous ray blocks over all
visible GPUs and sums the profiles on the host in device order.  This is what a Julia `make_beam` on a multi-GPU box
binds; bench.py uses the one-process-per-GPU / NCCL route instead."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_dp, c_ip, c_vp
from . import absorption


def _p(a):
    return a.ctypes.data_as(c_dp)


class MultiGPU:
    def __init__(self, n_devices: int = 0):
        self.h = c_vp()
        _lib.check(_lib.lib().torj_multi_create(int(n_devices), C.byref(self.h)))
        self.n_devices = int(_lib.lib().torj_multi_device_count(self.h))
        self._plasmas = {}

    def abs_Al_init(self, N_absz: int):
        t, w = np.polynomial.legendre.leggauss(int(N_absz))
        absorption._int_absz, absorption._int_weights = t, w
        _lib.check(_lib.lib().torj_multi_abs_init(self.h, int(N_absz), _p(t), _p(w)))

    def _plasma(self, plasma):
        key = id(plasma)
        if key not in self._plasmas:
            r = plasma._raw
            g = _lib.TorjGrid(len(plasma.R_coords), len(plasma.Z_coords), plasma.R_coords[0], plasma.R_coords[-1],
                              plasma.Z_coords[0], plasma.Z_coords[-1])
            t = lambda a: np.ascontiguousarray(a.T)
            keep = [t(r["psi"]), t(r["BR"]), t(r["BZ"]), t(r["Bphi"])]
            h = c_vp()
            _lib.check(_lib.lib().torj_multi_plasma_create_from_data(
                self.h, C.byref(g), _p(keep[0]), _p(np.ascontiguousarray(r["psi_prof"])), _p(np.ascontiguousarray(r["ne"])),
                _p(np.ascontiguousarray(r["Te"])), len(r["psi_prof"]), _p(keep[1]), _p(keep[2]), _p(keep[3]),
                _p(np.ascontiguousarray(r["psi1d"])), _p(np.ascontiguousarray(r["vol1d"])), len(r["psi1d"]), C.byref(h)))
            self._plasmas[key] = (h, plasma)
        return self._plasmas[key][0]

    def trace_bundle(self, plasma, ray_positions, ray_directions, ray_weights, f, mode, s_max, psi_dP_dV, *, options=None,
                     beam_id=None, n_beams=1):
        """Same arguments and result keys as torj_jl_b200.trace_bundle (without trajectories)."""
        L = _lib.lib()
        pos = np.ascontiguousarray(np.asarray(ray_positions, dtype=np.float64).T)
        dr = np.ascontiguousarray(np.asarray(ray_directions, dtype=np.float64).T)
        n = pos.shape[1]
        wt = np.ascontiguousarray(ray_weights, dtype=np.float64)
        per_ray = int(np.ndim(f) > 0)
        fr = np.ascontiguousarray(np.atleast_1d(f), dtype=np.float64)
        md = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(mode), fr.shape), dtype=np.int32)
        psi = np.ascontiguousarray(psi_dP_dV, dtype=np.float64)
        opt = options or _lib.default_options()
        n_beams = int(n_beams) if beam_id is not None else 1
        bid = np.ascontiguousarray(beam_id, dtype=np.int32) if beam_id is not None else None
        prof = np.zeros((n_beams, len(psi))); dep = np.zeros(n_beams)
        Pf = np.zeros(n); Pd = np.zeros(n); npts = np.zeros(n, dtype=np.int32); st = np.zeros(n, dtype=np.int32)
        cnt = _lib.TorjCounters()
        _lib.check(L.torj_multi_trace(self.h, self._plasma(plasma), C.byref(opt), n, _p(pos), _p(dr), _p(wt), _p(fr),
                                      md.ctypes.data_as(c_ip), per_ray, float(s_max), len(psi), _p(psi), n_beams,
                                      bid.ctypes.data_as(c_ip) if bid is not None else None, _p(prof), _p(dep), _p(Pf), _p(Pd),
                                      npts.ctypes.data_as(c_ip), st.ctypes.data_as(c_ip), 0, 0, 0, None, None, None, None, None,
                                      C.byref(cnt)))
        return dict(dP_dV=prof if n_beams > 1 else prof[0], deposited_power=dep if n_beams > 1 else float(dep[0]), P_final=Pf,
                    P_deposited_ray=Pd, n_points=npts, status=st, counters=cnt.as_dict())

    def close(self):
        if self.h:
            for h, _ in self._plasmas.values():
                _lib.lib().torj_multi_plasma_destroy(h)
            self._plasmas = {}
            _lib.lib().torj_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""N>1 host path on CPU: two gloo ranks shard a bundle by ray, each traces its block with the CPU oracle (standing in
for the device call), all-reduce the [dP_dV | deposited | sum w] vector; the result equals the unsharded run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torj_jl_b200 as tj
    from torj_jl_b200.distributed import allreduce_profile, shard_range
    from oracle import torj_oracle as O

    arr = tj.solovev_arrays(65, 65)
    opl = O.OraclePlasma(*arr.values())
    gl = np.polynomial.legendre.leggauss(24)
    x0 = np.array([2.5, 0.0, 0.4]); N0 = tj.pol_tor_angles_2_vector(np.deg2rad(30.0), 0.0)
    pos, dirs, w = tj.launch_peripheral_rays(x0, N0, 0.0174, 1 / 3.99, 95e9)
    psi = np.linspace(0, 1, 100)
    lo, hi = shard_range(len(w), rank, world)
    r = opl.trace_bundle(pos[lo:hi], dirs[lo:hi], w[lo:hi], 95e9, 1, 0.45, psi, gl, deposition="streaming", n_threads=1)
    t = torch.from_numpy(np.concatenate([r["dP_dV"], [r["deposited_power"], w[lo:hi].sum()]]))
    allreduce_profile(t)
    if rank == 0:
        full = opl.trace_bundle(pos, dirs, w, 95e9, 1, 0.45, psi, gl, deposition="streaming", n_threads=1)
        np.savez(os.path.join(out_dir, "res.npz"), sharded=t.numpy(), full=np.concatenate([full["dP_dV"], [full["deposited_power"], w.sum()]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_profile_equals_unsharded(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = np.load(tmp_path / "res.npz")
    assert abs(g["sharded"][-1] - 1.0) < 1e-13                       # sum of weights
    assert np.abs(g["sharded"] - g["full"]).max() <= 1e-12 * np.abs(g["full"]).max()

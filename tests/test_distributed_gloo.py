"""N>1 host path on CPU: two and three gloo ranks run the PRODUCT's sharding / all-reduce / per-ray assembly
(torj_jl_b200.distributed.trace_sharded, the function bench.py uses at N>1) around a fake device backend — a closed-form
stand-in for the device trace, so neither a GPU nor the oracle is involved. The assembled result must equal the same
backend applied to the unsharded bundle, for contiguous and block-cyclic sharding, ragged sizes included."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_PSI = 37


def fake_trace(pos, dirs, w):
    """Deterministic per-ray 'physics': what matters is that every output is a function of the ray alone and that the
    profile is the weighted sum over rays (reference src/solve.jl:233-240)."""
    key = pos[:, 0] * 3.0 + pos[:, 2] - dirs[:, 1]
    P_final = np.exp(-np.abs(key))
    shell = (np.floor(np.abs(key) * 7.0).astype(int)) % N_PSI
    prof = np.zeros(N_PSI)
    np.add.at(prof, shell, w * (1.0 - P_final))
    return dict(dP_dV=prof, deposited_power=float(np.sum(w * (1.0 - P_final))), P_final=P_final,
                n_points=(2 + np.floor(100 * np.abs(key))).astype(np.int32), status=(np.abs(key) > 5.0).astype(np.int32) * 2)


def bundle(n):
    k = np.arange(n, dtype=np.float64)
    pos = np.stack([np.sin(k), np.cos(0.3 * k), 0.01 * k], axis=1)
    dirs = np.stack([np.cos(k), np.sin(0.7 * k), np.ones(n)], axis=1)
    w = (1.0 + (k % 5)) / np.sum(1.0 + (k % 5))
    return pos, dirs, w


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torj_jl_b200.distributed import trace_sharded

    res = {}
    for name, n, kw in (("contig", 101, dict(sharding="contiguous")), ("cyclic", 101, dict(sharding="block_cyclic", block=8)),
                        ("tiny", 2, dict(sharding="block_cyclic", block=5))):
        pos, dirs, w = bundle(n)
        r = trace_sharded(fake_trace, pos, dirs, w, **kw)
        assert len(r["idx"]) == len(r["local"]["P_final"])
        for k in ("dP_dV", "deposited_power", "sum_weights", "P_final", "n_points", "status"):
            res[f"{name}_{k}"] = np.asarray(r[k])
        res[f"{name}_idx_{rank}"] = r["idx"]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


def _check(tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for name, n in (("contig", 101), ("cyclic", 101), ("tiny", 2)):
        pos, dirs, w = bundle(n)
        full = fake_trace(pos, dirs, w)
        owned = np.sort(np.concatenate([ranks[r][f"{name}_idx_{r}"] for r in range(world)]))
        assert np.array_equal(owned, np.arange(n))                      # the shards partition the bundle
        for g in ranks:                                                 # every rank holds the same assembled result
            assert abs(g[f"{name}_sum_weights"] - 1.0) < 1e-13
            assert np.abs(g[f"{name}_dP_dV"] - full["dP_dV"]).max() <= 1e-13
            assert abs(g[f"{name}_deposited_power"] - full["deposited_power"]) < 1e-13
            assert np.array_equal(g[f"{name}_P_final"], full["P_final"])
            assert np.array_equal(g[f"{name}_n_points"], full["n_points"]) and np.array_equal(g[f"{name}_status"], full["status"])
    if world > 1:
        a, b = ranks[0]["cyclic_idx_0"], ranks[1]["cyclic_idx_1"]
        assert a[:8].tolist() == list(range(8)) and b[:8].tolist() == list(range(8, 16))   # beams dealt round-robin


def test_two_rank_sharded_trace_equals_unsharded(tmp_path):
    _check(tmp_path, 2)


def test_three_rank_sharded_trace_equals_unsharded(tmp_path):
    _check(tmp_path, 3)

"""Multi-GPU plumbing: rays shard by contiguous index blocks, one process per GPU; the only exchange of the path
is the weighted sum of reference src/solve.jl:233-240 -> one small all-reduce of [dP_dV | deposited_power | sum w]."""
from __future__ import annotations


def shard_range(n_rays: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank; sizes differ by at most one."""
    base, rem = divmod(int(n_rays), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_block_cyclic(n_rays: int, block: int, rank: int, world_size: int):
    """Ray indices of rank when blocks of `block` consecutive rays (one beam of a scan) are dealt round-robin:
    block b goes to rank b % world_size. Contiguous sharding of an angle scan gives every rank a different range of
    launcher angles, i.e. rays of different lengths (measured: 6.2x on 8 GPUs instead of the ideal 8); dealing the
    beams out evens the work. Returns a sorted int64 array; the ranks' arrays partition range(n_rays)."""
    import numpy as np

    n_blocks = -(-int(n_rays) // int(block))
    mine = np.arange(rank, n_blocks, world_size, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < n_rays]


def allreduce_profile(t):
    """In-place sum over ranks of a tensor [n_psi+2] (NCCL on GPUs, gloo in CPU tests); no-op when not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t

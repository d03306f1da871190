"""Data parallelism over environments: one process per GPU, each owning an env slice and a replay shard.

Env stepping needs no communication; Philox keys use the GLOBAL env id so a run is invariant to how envs are sharded.  The only
collective of the path is ONE all-reduce of the flat SAC gradient bucket per update.  The helpers are backend-agnostic (NCCL on
the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard(rank: int, world: int, envs_per_rank: int):
    """-> (env_id_base, n_envs) of this rank's slice of the global env range [0, world * envs_per_rank)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * envs_per_rank, envs_per_rank


def allreduce_mean_(bucket, world: int):
    """In-place mean of the flat gradient bucket over all ranks (sum all-reduce, then scale)."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(bucket)
        bucket.mul_(1.0 / world)
    return bucket


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule of bench.py: the slowest rank defines the step time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank: int, world: int, seconds_max: float) -> float:
    """value = units all ranks processed / max-over-ranks time."""
    return world * units_per_rank / seconds_max


def reduce_path_stats(acc, counts, device=None):
    """Per-epoch collective of the logged statistics: `acc` holds {sum, sum of squares, max, min} x 4 quantities
    (algorithm.batched_path_information), `counts` the sample counts; sums and counts add over the ranks, maxima / minima combine."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return acc, counts
    acc = np.asarray(acc, np.float64)
    sums = torch.tensor(np.concatenate([acc[0::4], acc[1::4], np.asarray(counts, np.float64)]), dtype=torch.float64, device=device)
    mx = torch.tensor(acc[2::4], dtype=torch.float64, device=device)
    mn = torch.tensor(acc[3::4], dtype=torch.float64, device=device)
    dist.all_reduce(sums); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    sums, out = sums.cpu().numpy(), acc.copy()
    out[0::4], out[1::4], out[2::4], out[3::4] = sums[0:4], sums[4:8], mx.cpu().numpy(), mn.cpu().numpy()
    return out, sums[8:]

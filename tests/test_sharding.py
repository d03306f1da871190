"""Multi-rank host logic on CPU (gloo, world_size 2): env sharding, the single gradient-bucket all-reduce, max-over-ranks timing."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robosuite_benchmark_b200.parallel import allreduce_mean_, max_over_ranks, shard, whole_job_rate
    base, n = shard(rank, world, 8)
    bucket = torch.full((1000,), float(rank + 1))
    allreduce_mean_(bucket, world)
    t = max_over_ranks(0.5 + rank)
    out[rank] = (base, n, float(bucket[0]), float(bucket.std()), t, whole_job_rate(8 * 10, world, t))
    dist.destroy_process_group()


def test_two_rank_sharding_allreduce_and_timing():
    world, port = 2, 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert (r0[0], r0[1]) == (0, 8) and (r1[0], r1[1]) == (8, 8)          # disjoint, contiguous global env ids
    assert r0[2] == r1[2] == 1.5 and r0[3] == 0.0                          # mean of the two buckets, identical on both ranks
    assert r0[4] == r1[4] == 1.5                                           # slowest rank defines the time
    assert r0[5] == 2 * 80 / 1.5


def test_results_do_not_depend_on_the_sharding():
    """Env g of the global range behaves identically whether it is env g of one rank or env g - base of another (global Philox ids)."""
    from robosuite_benchmark_b200.controllers import load_controller_config
    from robosuite_benchmark_b200.model.tasks import build_task
    from tests.emu.emu import EmuEnv
    m, t = build_task("Lift", "Panda", load_controller_config(default_controller="OSC_POSE"), ignore_done=True)
    emu = EmuEnv(m, t, 16, 64)
    a = emu.reset(seed=17, env_id=11)
    b = emu.reset(seed=17, env_id=8 + 3)        # rank 1 of a 2 x 8 sharding, local env 3
    c = emu.reset(seed=17, env_id=3)
    assert (a == b).all() and not (a == c).all()
    assert (emu.random_action(17, 11, 5) == emu.random_action(17, 8 + 3, 5)).all()

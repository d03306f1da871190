"""Data-parallel training on 2 GPUs, one process per GPU over NCCL + NVLink symmetric memory:

  * 100 SAC updates with the gradient all-reduce FUSED into the optimizer kernel (csrc/rsb_dp.cu): the replicated parameters stay BIT-IDENTICAL across
    the ranks, and equal the NCCL arm (graph -> dist.all_reduce -> graph) to fp32 round-off of the two-term sum;
  * the end-to-end training loop (`build_experiment`, what `torchrun -m robosuite_benchmark_b200.train` runs): env shards with disjoint global ids,
    one replay shard per rank, statistics reduced over the ranks, rank 0 alone writes progress.csv / params.pkl.
Skipped on a single-GPU box (the round-end driver run); run with `gpurun --gpus 2`.
"""
import csv
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _worker(rank, world, port, tmp, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from robosuite_benchmark_b200.sac import EnvReplayBuffer, ParamStore, SACTrainer
    O, A, B = 42, 7, 128
    res = {}
    finals, early = {}, {}
    for arm in ("fused", "nccl"):
        rb = EnvReplayBuffer(20000, obs_dim=O, action_dim=A, device=dev, seed=100 + rank)          # a different shard per rank
        g = torch.Generator(device=dev); g.manual_seed(5 + rank)
        obs = torch.randn(20000, O, device=dev, generator=g) * 0.5
        rb.add_batch(obs, torch.tanh(torch.randn(20000, A, device=dev, generator=g)), torch.rand(20000, device=dev, generator=g) * 0.1,
                     torch.zeros(20000, dtype=torch.uint8, device=dev), obs + 0.05)
        store = ParamStore(O, A, dev, seed=3, symmetric=(arm == "fused"))
        tr = SACTrainer(store=store, replay_buffer=rb, batch_size=B, discount=0.99, policy_lr=1e-3, qf_lr=5e-4, soft_target_tau=0.005, target_update_period=5,
                        seed=9, world_size=world, rank=rank, allreduce=arm)
        for _ in range(2):
            tr.train_step()
        torch.cuda.synchronize()
        early[arm] = store.flat.clone()
        for _ in range(98):
            tr.train_step()
        torch.cuda.synchronize()
        flat = torch.cat([store.flat, store.target, store.m, store.v])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        res[arm + "_identical"] = bool(all(torch.equal(gathered[0], x) for x in gathered[1:]))
        res[arm + "_finite"] = bool(torch.isfinite(flat).all())
        res[arm + "_alpha"] = float(tr.alpha[0].item())
        finals[arm] = store.flat.clone()
        res[arm + "_dp_timeouts"] = tr.dp_timeouts()
    res["fused_vs_nccl"] = float((early["fused"] - early["nccl"]).abs().max().item())              # after 2 updates (Adam amplifies last-bit differences later on)
    res["fused_vs_nccl_100"] = float((finals["fused"] - finals["nccl"]).abs().max().item())
    # end-to-end loop
    from robosuite_benchmark_b200.algorithm import build_experiment
    v = json.load(open(os.path.join(GOLDEN, "variant_Lift-Panda-OSC-POSE-SEED17.json")))
    v["algorithm_kwargs"].update(num_epochs=2, num_eval_steps_per_epoch=2 * 16 * 10, num_expl_steps_per_train_loop=2 * 16 * 10, num_trains_per_train_loop=20,
                                 min_num_steps_before_training=2 * 16 * 10, expl_max_path_length=10, eval_max_path_length=10)
    v["replay_buffer_size"] = 8192
    algo = build_experiment(v, num_envs=16, log_dir=os.path.join(tmp, "log"), seed=17)
    algo.train()
    torch.cuda.synchronize()
    flat = algo.trainer.store.flat
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    res["loop_identical"] = bool(all(torch.equal(gathered[0], x) for x in gathered[1:]))
    res["loop_replay_size"] = algo.replay_buffer._size
    res["loop_env_base"] = algo.expl_env.sim.info("nenvs")
    res["loop_eval_paths"] = algo.last_eval_info.get("Num Paths")
    res["loop_expl_paths"] = algo.last_expl_info.get("Num Paths")
    res["loop_logger"] = algo.logger is not None
    out[rank] = res
    dist.destroy_process_group()


def test_two_gpu_fused_allreduce_and_training_loop(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, str(tmp_path), out), nprocs=world, join=True)
        r0, r1 = dict(out[0]), dict(out[1])
    print(r0)
    for r in (r0, r1):
        assert r["fused_identical"] and r["fused_finite"] and r["fused_dp_timeouts"] == 0          # replicated parameters: bit-identical after 100 updates
        assert r["nccl_identical"] and r["nccl_finite"]
        assert r["fused_vs_nccl"] < 2e-5, r["fused_vs_nccl"]                                        # same mean as NCCL's up to fp32 round-off of the sum, through 2 Adam steps
        assert r["fused_vs_nccl_100"] < 5e-2                                                        # ... and the same trajectory within Adam's amplification of it
        assert r["loop_identical"] and r["loop_replay_size"] == 3 * 16 * 10                          # per rank: warm-up + 2 epochs, each half of the job's 320 steps
        assert r["loop_eval_paths"] == 32 and r["loop_expl_paths"] == 32                              # statistics cover BOTH ranks' paths
    assert r0["loop_logger"] and not r1["loop_logger"]
    runs = glob.glob(str(tmp_path / "log"))
    rows = list(csv.DictReader(open(os.path.join(runs[0], "progress.csv"))))
    assert len(rows) == 2 and float(rows[1]["evaluation/Num Paths"]) == 32.0

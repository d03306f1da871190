"""`python -m robosuite_benchmark_b200.train --variant V.json [--seed S] [--log_dir D] [--num_envs N]`
`python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 -m robosuite_benchmark_b200.train --variant V.json --num_envs N`
(data parallel: G ranks x N envs, one replay shard per rank, gradients averaged over the ranks every update; rank 0 logs)

The reference's entry point (scripts/train.py:25-133) on the batched backend: same `--variant` JSON schema
(scripts/train.py:53-77; runs/*/variant.json), same log-directory naming ("{env}_{robots}_{controller}_SEED{seed}"), same outputs
(variant.json, progress.csv with the reference's 83 columns, params.pkl, debug.log).  Without --variant the flags of
util/arguments.py build the variant (defaults copied from there).
"""
from __future__ import annotations

import argparse
import datetime
import json
import os

import numpy as np


def build_parser():
    p = argparse.ArgumentParser()
    # robosuite args (util/arguments.py:23-84)
    p.add_argument("--env", type=str, default="Lift")
    p.add_argument("--robots", nargs="+", type=str, default=["Panda"])
    p.add_argument("--eval_horizon", type=int, default=500)
    p.add_argument("--expl_horizon", type=int, default=500)
    p.add_argument("--policy_freq", type=int, default=20)
    p.add_argument("--controller", type=str, default="OSC_POSE")
    p.add_argument("--reward_scale", type=float, default=1.0)
    p.add_argument("--hard_reset", action="store_true")
    p.add_argument("--env_config", type=str, default=None)
    # agent args (util/arguments.py:87-156)
    p.add_argument("--agent", type=str, default="SAC")
    p.add_argument("--qf_hidden_sizes", nargs="+", type=int, default=[256, 256])
    p.add_argument("--policy_hidden_sizes", nargs="+", type=int, default=[256, 256])
    p.add_argument("--gamma", type=float, default=0.99)
    p.add_argument("--policy_lr", type=float, default=3e-4)
    p.add_argument("--qf_lr", type=float, default=3e-4)
    p.add_argument("--soft_target_tau", type=float, default=5e-3)
    p.add_argument("--target_update_period", type=int, default=1)
    p.add_argument("--no_auto_entropy_tuning", action="store_true")
    # training args (util/arguments.py:159-202)
    p.add_argument("--variant", type=str, default=None)
    p.add_argument("--n_epochs", type=int, default=2000)
    p.add_argument("--trains_per_train_loop", type=int, default=1000)
    p.add_argument("--expl_ep_per_train_loop", type=int, default=10)
    p.add_argument("--steps_before_training", type=int, default=1000)
    p.add_argument("--batch_size", type=int, default=256)
    p.add_argument("--num_eval", type=int, default=10)
    p.add_argument("--log_dir", type=str, default="../log/runs/")
    p.add_argument("--seed", type=int, default=1)
    p.add_argument("--replay_buffer_size", type=int, default=int(1e6))
    # batched backend
    p.add_argument("--num_envs", type=int, default=1, help="environments stepped together on the GPU (1 = the reference's data flow)")
    p.add_argument("--device", type=str, default=None, help="default: cuda:LOCAL_RANK (one process per GPU under torchrun)")
    p.add_argument("--gemm", type=str, default="tcgen05", choices=["tcgen05", "cublas", "cublas_fp32"], help="GEMM back-end of the SAC update")
    return p


def variant_from_args(a):
    def env_kwargs(horizon, reward_scale):
        d = dict(env_name=a.env, robots=a.robots, horizon=horizon, control_freq=a.policy_freq, controller=a.controller,
                 reward_scale=reward_scale, hard_reset=a.hard_reset, ignore_done=True)
        if a.env_config is not None:
            d["env_configuration"] = a.env_config
        return d
    return dict(algorithm=a.agent, seed=a.seed, version="normal", replay_buffer_size=a.replay_buffer_size,
                qf_kwargs=dict(hidden_sizes=a.qf_hidden_sizes), policy_kwargs=dict(hidden_sizes=a.policy_hidden_sizes),
                algorithm_kwargs=dict(num_epochs=a.n_epochs, num_eval_steps_per_epoch=a.eval_horizon * a.num_eval,
                                      num_trains_per_train_loop=a.trains_per_train_loop,
                                      num_expl_steps_per_train_loop=a.expl_horizon * a.expl_ep_per_train_loop,
                                      min_num_steps_before_training=a.steps_before_training, expl_max_path_length=a.expl_horizon,
                                      eval_max_path_length=a.eval_horizon, batch_size=a.batch_size),
                trainer_kwargs=dict(discount=a.gamma, soft_target_tau=a.soft_target_tau, target_update_period=a.target_update_period,
                                    policy_lr=a.policy_lr, qf_lr=a.qf_lr, reward_scale=a.reward_scale,
                                    use_automatic_entropy_tuning=not a.no_auto_entropy_tuning),
                expl_environment_kwargs=env_kwargs(a.expl_horizon, a.reward_scale), eval_environment_kwargs=env_kwargs(a.eval_horizon, 1.0))


def run_experiment(args):
    from .algorithm import experiment
    if args.variant is not None:
        try:
            with open(args.variant) as f:
                variant = json.load(f)
        except FileNotFoundError:
            raise FileNotFoundError("Error opening specified variant json at: {}. Please check filepath and try again.".format(args.variant))
    else:
        variant = variant_from_args(args)
    ek = variant["expl_environment_kwargs"]
    tmp_file_prefix = "{}_{}_{}_SEED{}".format(ek["env_name"], "".join(ek["robots"]), ek["controller"], args.seed)
    stamp = datetime.datetime.now().strftime("%Y_%m_%d_%H_%M_%S")
    log_dir = os.path.join(args.log_dir, tmp_file_prefix, f"{tmp_file_prefix}_{stamp}_0000--s-0")     # written by rank 0 only
    np.random.seed(args.seed)                         # scripts/train.py:112-113 seeds from the flag, not from variant["seed"]
    import torch
    torch.manual_seed(args.seed)
    return experiment(variant, agent=variant.get("algorithm", "SAC"), num_envs=args.num_envs, device=args.device, log_dir=log_dir, seed=args.seed,
                      gemm=args.gemm), log_dir


if __name__ == "__main__":
    algo, log_dir = run_experiment(build_parser().parse_args())
    if algo.rank == 0:
        print("logs:", log_dir)
    if algo.world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()

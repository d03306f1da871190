"""`python -m robosuite_benchmark_b200.rollout --load_dir <run dir> [--num_episodes N] [--horizon H] [--num_envs E]`

The reference's `scripts/rollout.py` (scripts/rollout.py:86-160; util/rlkit_utils.py:168-300 `simulate_policy` / `evaluate_policy`) on
the batched backend: read `variant.json` + `params.pkl` of a run directory, rebuild the evaluation env from
`variant["eval_environment_kwargs"]`, roll out `evaluation/policy` (deterministic: tanh of the mean) and report the returns.
Episodes run in parallel (one env per episode).  No rendering / video (`--record_video` of the reference is out of scope: DESIGN.md §6).
`params.pkl` may be one written by this package (torch-pickled network objects) or one of the reference's committed snapshots (pickled
rlkit modules: read through `policy_io`, no rlkit needed).
"""
from __future__ import annotations

import argparse
import json
import os

import numpy as np


def load_policy(params_pkl: str, key: str = "evaluation/policy"):
    """-> TanhGaussianPolicy handle (unbound: weights on the host until first use) from a snapshot written by this package (torch-pickled
    network objects) or from one of the reference's committed snapshots (pickled rlkit modules: read through `policy_io`, no rlkit needed)."""
    import torch
    from .policy_io import load_params_pkl, mlp_weights
    from .sac import MakeDeterministic, TanhGaussianPolicy, register_safe_globals
    register_safe_globals()
    try:
        obj = torch.load(params_pkl, map_location="cpu", weights_only=True)[key]
        pol = obj.stochastic_policy if isinstance(obj, MakeDeterministic) else obj
        if not isinstance(pol, TanhGaussianPolicy):
            raise TypeError(type(pol))
        return pol
    except Exception:
        w = mlp_weights(load_params_pkl(params_pkl)[key])            # the reference's module pickles
        return policy_from_state_dict(w)


def policy_from_state_dict(w):
    """rlkit state-dict names (`fc0.weight` [out, in], ..., `last_fc`, `last_fc_log_std`) -> unbound TanhGaussianPolicy handle."""
    from .sac import TanhGaussianPolicy
    g = lambda k: np.asarray(w[k].detach().cpu().numpy() if hasattr(w[k], "detach") else w[k], np.float32)
    O, A = g("fc0.weight").shape[1], g("last_fc.weight").shape[0]
    pol = TanhGaussianPolicy(obs_dim=O, action_dim=A)
    W2 = np.concatenate([g("last_fc.weight").T, g("last_fc_log_std.weight").T if "last_fc_log_std.weight" in w else np.zeros((256, A), np.float32)], axis=1)
    b2 = np.concatenate([g("last_fc.bias"), g("last_fc_log_std.bias") if "last_fc_log_std.bias" in w else np.zeros(A, np.float32)])
    pol._host = {"W0": np.ascontiguousarray(g("fc0.weight").T), "b0": g("fc0.bias"), "W1": np.ascontiguousarray(g("fc1.weight").T), "b1": g("fc1.bias"),
                 "W2": np.ascontiguousarray(W2), "b2": b2}
    return pol


def evaluate_policy(env_config, policy, num_episodes=10, horizon=None, device="cuda:0", seed=0):
    """Returns of `num_episodes` deterministic episodes (util/rlkit_utils.py:208-300 without the printing / video): one env per episode,
    the policy forward kernel and the step kernel per control step, nothing else."""
    import torch
    import robosuite_benchmark_b200 as suite
    from .controllers import ALL_CONTROLLERS, load_controller_config
    if isinstance(policy, dict):
        policy = policy_from_state_dict(policy)
    cfg = dict(env_config)
    horizon = int(horizon or cfg.get("horizon", 500))
    cfg["horizon"] = horizon
    controller = cfg.pop("controller")                              # util/rlkit_utils.py:39-47
    ccfg = load_controller_config(default_controller=controller) if controller in ALL_CONTROLLERS else load_controller_config(custom_fpath=controller)
    env = suite.make(**cfg, has_renderer=False, has_offscreen_renderer=False, use_object_obs=True, use_camera_obs=False, reward_shaping=True,
                     controller_configs=ccfg, num_envs=max(2, num_episodes), batched=True, device=device, seed=seed, env_id_base=1 << 21)
    sim = env.sim
    assert policy.obs_dim == sim.obs_dim and policy.action_dim == sim.act_dim, \
        f"policy ({policy.obs_dim} -> {policy.action_dim}) does not fit the env ({sim.obs_dim} -> {sim.act_dim})"
    obs = sim.reset()
    act = torch.empty(sim.num_envs, sim.act_dim, device=sim.device)
    rew = torch.empty(sim.num_envs, device=sim.device)
    done = torch.empty(sim.num_envs, dtype=torch.uint8, device=sim.device)
    ret = torch.zeros(sim.num_envs, device=sim.device)
    for _ in range(horizon):
        policy.get_actions(obs, deterministic=True, out=act)
        sim.step(act, obs, rew, done)
        ret += rew
    return ret[:num_episodes].cpu().numpy()


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--load_dir", type=str, required=True, help="run directory holding variant.json and params.pkl")
    p.add_argument("--num_episodes", type=int, default=10)
    p.add_argument("--horizon", type=int, default=None)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--device", type=str, default="cuda:0")
    a = p.parse_args(argv)
    with open(os.path.join(a.load_dir, "variant.json")) as f:
        variant = json.load(f)
    policy = load_policy(os.path.join(a.load_dir, "params.pkl"))
    rets = evaluate_policy(variant["eval_environment_kwargs"], policy, a.num_episodes, a.horizon, a.device, a.seed)
    for i, r in enumerate(rets):
        print(f"Rollout episode {i}: return {r:.3f}")
    print(f"Average return over {len(rets)} episodes: {rets.mean():.3f} (std {rets.std():.3f})")
    return rets


if __name__ == "__main__":
    main()

"""`python -m robosuite_benchmark_b200.rollout --load_dir <run dir> [--num_episodes N] [--horizon H] [--num_envs E]`

The reference's `scripts/rollout.py` (scripts/rollout.py:86-160; util/rlkit_utils.py:168-300 `simulate_policy` / `evaluate_policy`) on
the batched backend: read `variant.json` + `params.pkl` of a run directory, rebuild the evaluation env from
`variant["eval_environment_kwargs"]`, roll out `evaluation/policy` (deterministic: tanh of the mean) and report the returns.
Episodes run in parallel (one env per episode).  No rendering / video (`--record_video` of the reference is out of scope: DESIGN.md §6).
`params.pkl` may be one written by this package (state dicts) or one of the reference's committed snapshots (pickled rlkit modules:
read through `policy_io`, no rlkit needed).
"""
from __future__ import annotations

import argparse
import json
import os
import pickle

import numpy as np


def load_policy_weights(params_pkl: str, key: str = "evaluation/policy"):
    from .policy_io import load_params_pkl, mlp_weights
    try:
        with open(params_pkl, "rb") as f:
            snap = pickle.load(f)                                   # this package's snapshots: plain state dicts
        sd = snap[key]
        return {k: np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, np.float64) for k, v in sd.items()}
    except Exception:
        return mlp_weights(load_params_pkl(params_pkl)[key])        # the reference's module pickles


def evaluate_policy(env_config, weights, num_episodes=10, horizon=None, device="cuda:0", seed=0):
    """Returns of `num_episodes` deterministic episodes (util/rlkit_utils.py:208-300 without the printing / video)."""
    import torch
    import robosuite_benchmark_b200 as suite
    from .controllers import ALL_CONTROLLERS, load_controller_config
    cfg = dict(env_config)
    horizon = int(horizon or cfg.get("horizon", 500))
    cfg["horizon"] = horizon
    controller = cfg.pop("controller")                              # util/rlkit_utils.py:39-47
    ccfg = load_controller_config(default_controller=controller) if controller in ALL_CONTROLLERS else load_controller_config(custom_fpath=controller)
    env = suite.make(**cfg, has_renderer=False, has_offscreen_renderer=False, use_object_obs=True, use_camera_obs=False, reward_shaping=True,
                     controller_configs=ccfg, num_envs=max(2, num_episodes), batched=True, device=device, seed=seed, env_id_base=1 << 21)
    sim = env.sim
    W = {k: torch.tensor(v, dtype=torch.float32, device=sim.device) for k, v in weights.items()}
    assert W["fc0.weight"].shape[1] == sim.obs_dim and W["last_fc.weight"].shape[0] == sim.act_dim, \
        f"policy ({W['fc0.weight'].shape[1]} -> {W['last_fc.weight'].shape[0]}) does not fit the env ({sim.obs_dim} -> {sim.act_dim})"
    obs = sim.reset()
    ret = torch.zeros(sim.num_envs, device=sim.device)
    for _ in range(horizon):
        h = torch.relu(obs @ W["fc0.weight"].T + W["fc0.bias"])
        k = 1
        while f"fc{k}.weight" in W:
            h = torch.relu(h @ W[f"fc{k}.weight"].T + W[f"fc{k}.bias"]); k += 1
        a = torch.tanh(h @ W["last_fc.weight"].T + W["last_fc.bias"])
        obs, r, _ = sim.step(a.contiguous())
        ret += r
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
    weights = load_policy_weights(os.path.join(a.load_dir, "params.pkl"))
    rets = evaluate_policy(variant["eval_environment_kwargs"], weights, a.num_episodes, a.horizon, a.device, a.seed)
    for i, r in enumerate(rets):
        print(f"Rollout episode {i}: return {r:.3f}")
    print(f"Average return over {len(rets)} episodes: {rets.mean():.3f} (std {rets.std():.3f})")
    return rets


if __name__ == "__main__":
    main()

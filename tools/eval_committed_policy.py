#!/usr/bin/env python3
"""roll a committed 2020 policy (runs/<family>-SEED*/.../params.pkl) out in the batched CUDA env and compare the
return with the run's own logged evaluation returns.  Needs /root/reference (this container) and a GPU (run under gpurun with the
weights exported first: `python tools/eval_committed_policy.py export` writes gpurun_out/policy_<run>.npz here).
  python tools/eval_committed_policy.py export [run]      (CPU, here)
  python tools/eval_committed_policy.py run [run]         (GPU box)"""
import sys, os, glob, csv
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
mode = sys.argv[1] if len(sys.argv) > 1 else "export"
run = sys.argv[2] if len(sys.argv) > 2 else "Lift-Panda-OSC-POSE-SEED17"
npz = os.path.join(ROOT, "tests", "golden", f"policy_{run}.npz")
if mode == "export":
    from robosuite_benchmark_b200.policy_io import load_params_pkl, mlp_weights
    path = glob.glob(f"/root/reference/runs/{run}/*/params.pkl")[0]
    w = mlp_weights(load_params_pkl(path)["trainer/policy"])
    rows = list(csv.DictReader(open(glob.glob(f"/root/reference/runs/{run}/*/progress.csv")[0])))
    logged = np.array([float(r["evaluation/Average Returns"]) for r in rows])
    os.makedirs(os.path.dirname(npz), exist_ok=True)
    np.savez_compressed(npz, logged=logged.astype(np.float32), **{k: v.astype(np.float32) for k, v in w.items()})
    print("wrote", npz, "logged eval return: last-50 mean %.1f max %.1f" % (logged[-50:].mean(), logged.max()))
else:
    import torch
    import robosuite_benchmark_b200 as suite
    from robosuite_benchmark_b200.policy_io import DeterministicPolicy
    d = dict(np.load(npz)); logged = d.pop("logged"); pol = DeterministicPolicy(d)
    E = 256; dev = torch.device("cuda", 0)
    cfg = suite.load_controller_config(default_controller="OSC_POSE")
    env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
    W = {k: torch.tensor(v, dtype=torch.float32, device=dev) for k, v in d.items()}
    obs = env.sim.reset(); ret = torch.zeros(E, device=dev)
    for k in range(500):
        h = torch.relu(obs @ W["fc0.weight"].T + W["fc0.bias"]); h = torch.relu(h @ W["fc1.weight"].T + W["fc1.bias"])
        a = torch.tanh(h @ W["last_fc.weight"].T + W["last_fc.bias"])
        obs, r, _ = env.sim.step(a.contiguous()); ret += r
    print(f"{run}: return in this sim over {E} episodes: mean {ret.mean().item():.1f} (sd {ret.std().item():.1f}, max {ret.max().item():.1f}); "
          f"logged by the reference run: last-50-epoch mean {logged[-50:].mean():.1f}, max {logged.max():.1f}")

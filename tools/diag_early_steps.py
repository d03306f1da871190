#!/usr/bin/env python3
"""Kernel time of the first 25 control steps after a reset when an UNTRAINED policy drives 4096 Lift envs (what bench.py --mode train collects per epoch), under both
OSC orientation rules, deterministic and sampled actions (developer tool, gpurun)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import robosuite_benchmark_b200 as suite
from robosuite_benchmark_b200.sac import ParamStore, TanhGaussianPolicy
dev = torch.device("cuda", 0); E = 4096
for mode in ("euler_transpose", "axis_angle"):
    for det in (True, False):
        for seed in (17, 18):
            cfg = suite.load_controller_config(default_controller="OSC_POSE"); cfg["orientation_delta"] = mode
            env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
            sim = env.sim; pol = TanhGaussianPolicy.of(ParamStore(42, 7, dev, seed=seed))
            act = torch.empty(E, 7, device=dev); rew = torch.empty(E, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); tot = 0.0
            for ep in range(4):
                obs = sim.reset(); torch.cuda.synchronize(); t0.record()
                for k in range(25):
                    pol.get_actions(obs, deterministic=det, out=act, step=ep * 25 + k); sim.step(act, obs, rew, done)
                t1.record(); torch.cuda.synchronize()
                if ep: tot += t0.elapsed_time(t1)
            print(f"{mode:16s} deterministic={det!s:5s} policy seed {seed}: {tot / 75:.3f} ms per control step (policy kernel + k_step), mean |action| {act.abs().mean().item():.2f}, counters {sim.counters()}, "
                  f"Newton iterations per substep {sim.newton_iterations().float().mean().item() / 25:.2f}", flush=True)
            env.close()

#!/usr/bin/env python3
"""Truncation counters of a full random-action episode of 4096 envs under candidate (ncon_max, nefc_max) limits, with the kernel time per control step each costs
(developer tool, gpurun):  python tools/limits_sweep.py Lift Panda OSC_POSE 16,64 20,64 20,72 24,80
RSB_SWEEP_POLICY=1: drive the envs with the committed Lift-Panda-OSC_POSE policy (tests/golden) instead of random actions -- grasping makes more contacts."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import robosuite_benchmark_b200 as suite
dev = torch.device("cuda", 0)
name, robots, ctrl = sys.argv[1], sys.argv[2].split("+"), sys.argv[3]
cfg = suite.load_controller_config(default_controller=ctrl)
pol = None
if os.environ.get("RSB_SWEEP_POLICY") == "1":
    import numpy as np
    from robosuite_benchmark_b200.rollout import policy_from_state_dict
    d = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "policy_Lift-Panda-OSC-POSE-SEED17.npz"))); d.pop("logged")
    pol = policy_from_state_dict(d)
for lim in sys.argv[4:]:
    nc, ne = map(int, lim.split(","))
    E = 4096
    env = suite.make(name, robots, controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True,
                     ncon_max=nc, nefc_max=ne)
    sim = env.sim
    obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
    sim.reset(obs=obs); first = None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(500):
        if pol is None: sim.random_actions(k, out=act)
        else: pol.get_actions(obs, deterministic=(k % 2 == 0), out=act, step=k)       # alternate the mean action and a sampled one
        if k == 400: t0.record()
        sim.step(act, obs, rew, done)
        if first is None and k % 10 == 9:
            c = sim.counters()
            if c["ncon_overflow"] or c["nefc_overflow"]: first = (k, dict(c))
    t1.record(); torch.cuda.synchronize()
    print("policy-driven" if pol is not None else "random actions", name, robots, ctrl, "limits", (nc, ne), "counters", sim.counters(), "first overflow by step", first, "| ms/control step (steps 400-499): %.3f" % (t0.elapsed_time(t1) / 100),
          "envs/block", sim.info("envs_per_block") if hasattr(sim, "info") else "", flush=True)
    sim.close(); del sim, env

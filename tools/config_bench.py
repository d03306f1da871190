#!/usr/bin/env python3
"""k_step throughput of the other BASELINE.json configs (developer tool, run under gpurun).  Not the bench line: bench.py measures configs[1]."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import robosuite_benchmark_b200 as suite
dev = torch.device("cuda", 0)
CFG = [("Lift", "Panda", "OSC_POSE", 4096), ("Door", "Panda", "JOINT_VELOCITY", 16384), ("Stack", "Sawyer", "OSC_POSE", 4096), ("TwoArmLift", ["Panda", "Panda"], "OSC_POSE", 4096)]
for name, robots, ctrl, E in CFG:
    cfg = suite.load_controller_config(default_controller=ctrl)
    kw = dict(env_configuration="single-arm-opposed") if name == "TwoArmLift" else {}
    env = suite.make(name, robots, controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True, **kw)
    sim = env.sim
    obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
    sim.reset(obs=obs)
    pre, steps = 100, 60
    for k in range(pre):
        sim.random_actions(k, out=act); sim.step(act, obs, rew, done)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        sim.random_actions(pre + k, out=act); ev[k][0].record(); sim.step(act, obs, rew, done); ev[k][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    print(f"{name}-{robots}-{ctrl}: {E} envs, {ms:.3f} ms/control step (steps {pre}-{pre+steps}) -> {E/ms*1000:.0f} control-steps/s; lanes {sim.info('lanes')}, "
          f"{sim.info('envs_per_block')} envs/CTA x {sim.info('blocks_per_sm')} CTA/SM, {sim.info('smem_bytes')} B smem/env, mean reward {rew.mean().item():.4f}")
    env.close() if hasattr(env, "close") else None
    del sim, env

#!/usr/bin/env python3
"""k_step time per control step across an episode (developer tool, run under gpurun): python tools/episode_profile.py [steps]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import robosuite_benchmark_b200 as suite
E = 4096; steps = int(sys.argv[1]) if len(sys.argv) > 1 else 520
dev = torch.device("cuda", 0)
cfg = suite.load_controller_config(default_controller="OSC_POSE")
env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
sim = env.sim
obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
sim.reset(obs=obs)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
rews = []
for k in range(steps):
    sim.random_actions(k, out=act)
    ev[k][0].record(); sim.step(act, obs, rew, done); ev[k][1].record()
    if (k + 1) % 500 == 0: sim.reset(obs=obs)
    if k % 50 == 49: rews.append(float(rew.mean().item()))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
for w in range(0, steps, 50):
    seg = ms[w:w + 50]
    print(f"steps {w:4d}-{w+len(seg)-1:4d}: mean {sum(seg)/len(seg):.3f} ms  min {min(seg):.3f} max {max(seg):.3f}  reward {rews[w//50] if w//50 < len(rews) else float('nan'):.4f}")
print(f"steady (steps 100+) mean {sum(ms[100:])/max(1,len(ms[100:])):.3f} ms")
print(f"overall mean {sum(ms)/len(ms):.3f} ms -> {E/(sum(ms)/len(ms))*1000:.0f} control-steps/s")

#!/usr/bin/env python3
"""Max ncon / nefc seen over a random-action rollout per family (developer tool, run under gpurun): are the per-task row limits tight?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import robosuite_benchmark_b200 as suite
dev = torch.device("cuda", 0)
FAMS = [("Lift", "Sawyer", "OSC_POSITION"), ("Lift", "Panda", "OSC_POSE"), ("Stack", "Sawyer", "OSC_POSE"), ("TwoArmLift", ["Panda", "Panda"], "OSC_POSE"), ("Door", "Panda", "JOINT_VELOCITY")]
for name, robots, ctrl in FAMS:
    cfg = suite.load_controller_config(default_controller=ctrl)
    E = 2048
    env = suite.make(name, robots, controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True,
                     ncon_max=48, nefc_max=160)
    sim = env.sim
    obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
    sim.reset(obs=obs); mc = me = 0; hist = []
    for k in range(481):
        sim.random_actions(k, out=act)
        if k % 20 == 0:
            st = sim.get_state(); d = sim.debug_substep(act, True)[:, :3].cpu().numpy(); sim.set_state(st)
            mc = max(mc, d[:, 0].max()); me = max(me, d[:, 1].max()); hist.append((int(d[:, 0].max()), int(d[:, 1].max()), float(np.percentile(d[:, 1], 99.9))))
        sim.step(act, obs, rew, done)
    print(name, robots, sim.counters(), "limits ncon_max", sim.info("ncon_max"), "nefc_max", sim.info("nefc_max"), "| seen max ncon", int(mc), "max nefc", int(me), "| (max ncon, max nefc, p99.9 nefc) per sample", hist)
    del sim, env

#!/usr/bin/env python3
"""Distribution of ncon / nefc / Newton iterations over the batch at a few points of an episode (developer tool, run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import robosuite_benchmark_b200 as suite
E = 4096; dev = torch.device("cuda", 0)
cfg = suite.load_controller_config(default_controller="OSC_POSE")
env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
sim = env.sim
obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
sim.reset(obs=obs)
for k in range(301):
    sim.random_actions(k, out=act)
    if k in (5, 60, 150, 300):
        st = sim.get_state()
        dbg = sim.debug_substep(act, True)
        d = dbg.cpu().numpy() if hasattr(dbg, "cpu") else np.asarray(dbg)
        ncon, nefc, it = d[:, 0], d[:, 1], d[:, 2]
        print(f"step {k}: ncon mean {ncon.mean():.2f} p50 {np.percentile(ncon,50):.0f} p99 {np.percentile(ncon,99):.0f} max {ncon.max():.0f} | nefc mean {nefc.mean():.1f} p99 {np.percentile(nefc,99):.0f} max {nefc.max():.0f} | iters mean {it.mean():.2f} p50 {np.percentile(it,50):.0f} p90 {np.percentile(it,90):.0f} p99 {np.percentile(it,99):.0f} max {it.max():.0f}")
        print("   iters hist", np.bincount(it.astype(int), minlength=13).tolist())
        print("   ncon hist", np.bincount(ncon.astype(int), minlength=17).tolist())
        if k == 150:
            hard = np.nonzero(it >= 8)[0]
            np.savez("gpurun_out/stragglers.npz", idx=hard, state=st.cpu().numpy()[hard], act=act.cpu().numpy()[hard], iters=it[hard], ncon=ncon[hard], nefc=nefc[hard])
        sim.set_state(st)
    sim.step(act, obs, rew, done)

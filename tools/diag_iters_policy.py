#!/usr/bin/env python3
"""Newton iterations per substep under the committed Lift policy's contact load: distribution over the batch at a few control steps, and against the number of
contacts (developer tool, gpurun).  RSB_SOLVER_ITERS etc. apply."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import robosuite_benchmark_b200 as suite
from robosuite_benchmark_b200.rollout import policy_from_state_dict
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "policy_Lift-Panda-OSC-POSE-SEED17.npz"))); d.pop("logged")
pol = policy_from_state_dict(d); dev = torch.device("cuda", 0); E = 4096
env = suite.make("Lift", "Panda", controller_configs=suite.load_controller_config(default_controller="OSC_POSE"), num_envs=E, batched=True, device=dev, seed=17, horizon=500,
                 control_freq=20, reward_shaping=True, ignore_done=True)
sim = env.sim; obs = sim.reset(); act = torch.empty(E, 7, device=dev); rew = torch.empty(E, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
print("solver option", sim.solver_option())
for k in range(400):
    pol.get_actions(obs, deterministic=True, out=act)
    if k in (50, 100, 200, 399):
        st = sim.get_state(); dbg = sim.debug_substep(act, True)[:, :3].cpu().numpy(); sim.set_state(st)        # ncon, nefc, iterations of the FIRST substep
    sim.step(act, obs, rew, done)
    if k in (50, 100, 200, 399):
        it = sim.newton_iterations().cpu().numpy() / 25.0
        q = np.percentile(it, [50, 90, 99, 100])
        byc = [(c, round(float(it[dbg[:, 0] == c].mean()), 2), int((dbg[:, 0] == c).sum())) for c in sorted(set(dbg[:, 0].astype(int))) if (dbg[:, 0] == c).sum() >= 20]
        print(f"step {k}: iterations per substep p50 {q[0]:.2f} p90 {q[1]:.2f} p99 {q[2]:.2f} max {q[3]:.2f}; envs >= 6/substep: {(it >= 6).mean():.3f}, >= 11: {(it >= 11).mean():.4f}; "
              f"mean max-per-CTA (28 consecutive envs) {np.mean([it[i:i + 28].max() for i in range(0, E, 28)]):.2f}; (ncon, mean iterations, envs): {byc}", flush=True)

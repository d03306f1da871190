#!/usr/bin/env python3
"""Third pass (after diag_policy_transfer.py / ...2.py): the committed policy transfers only when a rotation action turns the end effector the OTHER way.  Is the
sign really in the controller, or does the policy infer the gripper's attitude from JOINT observations whose sign convention differs in the arm authored here
(a mirrored joint axis)?  Under the POSITIVE (axis-angle) controller, negate sin(q_j) and qd_j of joint subsets in the observation (developer tool, gpurun)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import robosuite_benchmark_b200 as suite
from robosuite_benchmark_b200.rollout import policy_from_state_dict
run = "Lift-Panda-OSC-POSE-SEED17"
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"policy_{run}.npz"))); logged = d.pop("logged")
pol = policy_from_state_dict(d)
E, dev = 256, torch.device("cuda", 0)

def rollout(mode, joints=(), steps=500, seed=17):
    cfg = suite.load_controller_config(default_controller="OSC_POSE"); cfg["orientation_delta"] = mode
    env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=seed, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
    sim = env.sim
    obs = sim.reset(); ret = torch.zeros(E, device=dev); act = torch.empty(E, 7, device=dev); rew = torch.empty(E, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
    lifted = torch.zeros(E, device=dev)
    cols = [j for j in joints] + [14 + j for j in joints]                 # sin q_j and qd_j (cos q_j is even)
    for k in range(steps):
        o = obs
        if cols: o = obs.clone(); o[:, cols] = -o[:, cols]
        pol.get_actions(o, deterministic=True, out=act)
        sim.step(act, obs, rew, done); ret += rew
        lifted = torch.maximum(lifted, (obs[:, 34] > 0.84).float())
    env.close()
    return ret.mean().item(), ret.std().item() / np.sqrt(E), ret.max().item(), lifted.mean().item()

print(f"{run}: logged last-50-epoch mean {logged[-50:].mean():.1f}, max {logged.max():.1f}; {E} episodes per hypothesis")
H = [("euler_transpose", ()), ("axis_angle", ()), ("axis_angle", (0, 2, 4, 6)), ("axis_angle", (1, 3, 5)), ("axis_angle", tuple(range(7)))] + [("axis_angle", (j,)) for j in range(7)] + \
    [("axis_angle", (4, 6)), ("axis_angle", (4, 5, 6)), ("axis_angle", (5, 6)), ("euler_transpose", (0, 2, 4, 6))]
for mode, joints in H:
    m, se, mx, lf = rollout(mode, joints)
    print(f"  controller {mode:16s} joints negated in obs {str([j + 1 for j in joints]):24s} return mean {m:6.1f} +- {se:4.1f} max {mx:6.1f}  lifted {lf:.2f}", flush=True)

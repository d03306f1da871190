#!/usr/bin/env python3
"""Second pass of tools/diag_policy_transfer.py: WHICH negative orientation-delta convention did the committed policies train under?  The controller maps the scaled
rotation action delta to a goal rotation F(delta) R_ee.  Candidates are emulated outside the kernels by feeding the axis-angle vector of F(delta) (developer tool, gpurun)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import robosuite_benchmark_b200 as suite
from robosuite_benchmark_b200.rollout import policy_from_state_dict
run = sys.argv[1] if len(sys.argv) > 1 else "Lift-Panda-OSC-POSE-SEED17"
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"policy_{run}.npz"))); logged = d.pop("logged")
pol = policy_from_state_dict(d)
E, dev = 512, torch.device("cuda", 0)
cfg = suite.load_controller_config(default_controller="OSC_POSE")
cfg["orientation_delta"] = "axis_angle"     # the candidates below are expressed as axis-angle vectors fed to the axis-angle controller; the winner, euler2mat(d)^T,
                                            # has since become the shipped default ("euler_transpose", include/rsb_model.h RSB_ORI_DELTA_EULER_T)

def euler2mat(e):                       # mujoco-py / robosuite transform_utils.euler2mat
    ai, aj, ak = -e[:, 2], -e[:, 1], -e[:, 0]
    si, sj, sk, ci, cj, ck = torch.sin(ai), torch.sin(aj), torch.sin(ak), torch.cos(ai), torch.cos(aj), torch.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    m = torch.empty(e.shape[0], 3, 3, device=e.device, dtype=e.dtype)
    m[:, 2, 2] = cj * ck; m[:, 2, 1] = sj * sc - cs; m[:, 2, 0] = sj * cc + ss
    m[:, 1, 2] = cj * sk; m[:, 1, 1] = sj * ss + cc; m[:, 1, 0] = sj * cs - sc
    m[:, 0, 2] = -sj; m[:, 0, 1] = cj * si; m[:, 0, 0] = cj * ci
    return m

def logmap(R):
    tr = (R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]).double()
    ang = torch.acos(((tr - 1) / 2).clamp(-1, 1))
    v = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], 1).double()
    f = torch.where(ang > 1e-6, ang / (2 * torch.sin(ang).clamp_min(1e-12)), torch.full_like(ang, 0.5))
    return (v * f[:, None]).float()

def variant(kind):
    def f(a):
        delta = a[:, 3:6].clamp(-1, 1) * 0.5
        if kind == "axis-angle(-d)": R = None; a[:, 3:6] = -a[:, 3:6]; return a
        if kind == "euler2mat(-d)": R = euler2mat(-delta.double())
        elif kind == "euler2mat(d)^T": R = euler2mat(delta.double()).transpose(1, 2)
        elif kind == "euler2mat(-d)^T": R = euler2mat(-delta.double()).transpose(1, 2)
        elif kind == "euler2mat(d)": R = euler2mat(delta.double())
        a[:, 3:6] = logmap(R) / 0.5
        return a
    return f

def rollout(act_fn, steps=500, seed=17):
    env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=seed, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
    sim = env.sim
    obs = sim.reset(); ret = torch.zeros(E, device=dev); act = torch.empty(E, 7, device=dev); rew = torch.empty(E, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
    lifted = torch.zeros(E, device=dev)
    for k in range(steps):
        pol.get_actions(obs, deterministic=True, out=act)
        a = act_fn(act.clone()) if act_fn else act
        sim.step(a.contiguous(), obs, rew, done); ret += rew
        lifted = torch.maximum(lifted, (obs[:, 34] > 0.84).float())
    env.close()
    return ret.mean().item(), ret.std().item(), ret.max().item(), lifted.mean().item()

print(f"{run}: logged last-50-epoch mean {logged[-50:].mean():.1f}, last-200 mean {logged[-200:].mean():.1f}, max {logged.max():.1f}; {E} episodes per candidate")
for kind in ("euler2mat(-d)^T", "euler2mat(d)", "axis-angle(-d)", "euler2mat(-d)", "euler2mat(d)^T"):
    m, sd, mx, lf = rollout(variant(kind))
    print(f"  goal = {kind:16s} R_ee : return mean {m:6.1f} +- {sd / np.sqrt(E):4.1f} (sd {sd:5.1f}) max {mx:6.1f}  lifted {lf:.2f}", flush=True)

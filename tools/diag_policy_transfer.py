#!/usr/bin/env python3
"""Why does the committed Lift-Panda-OSC_POSE-SEED17 policy not transfer?  Roll it out in the CUDA env under observation / action convention
hypotheses applied OUTSIDE the kernels (developer tool, run under gpurun): a convention mismatch shows up as a jump of the return towards the logged 364."""
import itertools, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import robosuite_benchmark_b200 as suite
from robosuite_benchmark_b200.rollout import policy_from_state_dict
run = sys.argv[1] if len(sys.argv) > 1 else "Lift-Panda-OSC-POSE-SEED17"
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"policy_{run}.npz"))); logged = d.pop("logged")
pol = policy_from_state_dict(d)
E, dev = 128, torch.device("cuda", 0)
cfg = suite.load_controller_config(default_controller="OSC_POSE")
cfg["orientation_delta"] = "axis_angle"     # this study was run against the axis-angle controller; its finding (rotation sign) made "euler_transpose" the default

def rollout(obs_fn=None, act_fn=None, steps=500, seed=17):
    env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=seed, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
    sim = env.sim
    obs = sim.reset(); ret = torch.zeros(E, device=dev); act = torch.empty(E, 7, device=dev); rew = torch.empty(E, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
    lifted = torch.zeros(E, device=dev)
    for k in range(steps):
        o = obs_fn(obs.clone()) if obs_fn else obs
        pol.get_actions(o.contiguous(), deterministic=True, out=act)
        a = act_fn(act.clone()) if act_fn else act
        sim.step(a.contiguous(), obs, rew, done); ret += rew
        lifted = torch.maximum(lifted, (obs[:, 34] > 0.84).float())
    env.close()
    return ret.mean().item(), ret.max().item(), lifted.mean().item()

# obs layout (42): sin q 0:7, cos q 7:14, qd 14:21, eef_pos 21:24, eef_quat xyzw 24:28, grip q 28:30, grip qd 30:32, cube_pos 32:35, cube_quat xyzw 35:39, eef - cube 39:42
def f_grip_act(a): a[:, 6] = -a[:, 6]; return a
def f_grip_obs(o): o[:, 28:32] = -o[:, 28:32]; return o
def f_grip_swap(o): o[:, 28:30] = o[:, [29, 28]]; o[:, 30:32] = o[:, [31, 30]]; return o
def f_eef_wxyz(o): o[:, 24:28] = o[:, [27, 24, 25, 26]]; return o
def f_eef_neg(o): o[:, 24:28] = -o[:, 24:28]; return o
def f_cube_wxyz(o): o[:, 35:39] = o[:, [38, 35, 36, 37]]; return o
def f_rel_neg(o): o[:, 39:42] = -o[:, 39:42]; return o
def f_obj_first(o): return torch.cat([o[:, 32:42], o[:, 0:32]], 1)
def f_rot_act_neg(a): a[:, 3:6] = -a[:, 3:6]; return a
def f_pos_xy_neg(a): a[:, 0:2] = -a[:, 0:2]; return a
def chain(*fs):
    def g(x):
        for f in fs: x = f(x)
        return x
    return g
H = [("baseline", None, None), ("gripper action sign flipped", None, f_grip_act), ("gripper obs sign flipped", f_grip_obs, None), ("gripper obs + action flipped", f_grip_obs, f_grip_act),
     ("gripper joints swapped in obs", f_grip_swap, None), ("eef quat wxyz", f_eef_wxyz, None), ("eef quat negated", f_eef_neg, None), ("cube quat wxyz", f_cube_wxyz, None),
     ("gripper_to_cube negated", f_rel_neg, None), ("object-state first", f_obj_first, None), ("rotation action negated", None, f_rot_act_neg), ("xy action negated", None, f_pos_xy_neg),
     ("eef neg + grip act", f_eef_neg, f_grip_act), ("rel neg + grip act", f_rel_neg, f_grip_act)]
print(f"{run}: logged last-50-epoch mean {logged[-50:].mean():.1f}, max {logged.max():.1f}; {E} episodes per hypothesis")
for name, of, af in H:
    m, mx, lf = rollout(of, af)
    print(f"  {name:34s} return mean {m:7.1f} max {mx:7.1f}  fraction of episodes with the cube lifted > 4 cm at some step {lf:.2f}", flush=True)

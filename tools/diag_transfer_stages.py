#!/usr/bin/env python3
"""Where do the committed policies that do NOT transfer fail?  Per run: how far each stage of the task gets (reach / manipulate / succeed), from the observations of a
deterministic rollout (developer tool, gpurun; needs oracle/_ref/policies from tools/eval_committed_runs.py export).
  python tools/diag_transfer_stages.py [FILTER] [EPISODES]"""
import sys, os, glob, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import robosuite_benchmark_b200 as suite
from robosuite_benchmark_b200.controllers import load_controller_config
from robosuite_benchmark_b200.rollout import policy_from_state_dict
filt = sys.argv[1] if len(sys.argv) > 1 else ""
E = int(sys.argv[2]) if len(sys.argv) > 2 else 128
extra = json.loads(os.environ.get("RSB_EVAL_CONTROLLER_OVERRIDES", "{}"))
mk = json.loads(os.environ.get("RSB_EVAL_MAKE_OVERRIDES", "{}"))
dev = torch.device("cuda", 0)
for f in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "policies", "*.npz"))):
    run = os.path.basename(f)[:-4]
    if filt and filt not in run: continue
    d = dict(np.load(f)); logged = d.pop("logged"); cfg = json.loads(str(d.pop("env_kwargs"))); pol = policy_from_state_dict(d)
    cc = load_controller_config(default_controller=cfg.pop("controller"))
    if "kp" in extra: cc.pop("kv", None)
    cc.update(extra); cfg.update(mk)
    env = suite.make(**cfg, reward_shaping=True, controller_configs=cc, num_envs=E, batched=True, device=dev, seed=17)
    sim = env.sim; name = cfg["env_name"]
    obs = sim.reset(); act = torch.empty(E, sim.act_dim, device=dev); rew = torch.empty(E, device=dev); done = torch.empty(E, dtype=torch.uint8, device=dev)
    ret = torch.zeros(E, device=dev); st = {}
    def upd(k, v, mode):
        st[k] = v.clone() if k not in st else (torch.minimum(st[k], v) if mode == "min" else torch.maximum(st[k], v))
    o0 = obs.clone()
    for t in range(500):
        pol.get_actions(obs, deterministic=True, out=act); sim.step(act, obs, rew, done); ret += rew
        if name == "Lift":
            upd("min |eef-cube| [m]", obs[:, -3:].norm(dim=1), "min"); upd("max cube z - start [m]", obs[:, -8] - o0[:, -8], "max"); upd("max step reward", rew, "max")
        elif name == "Door":
            upd("min |eef-handle| [m]", obs[:, -5:-2].norm(dim=1), "min"); upd("max |handle_qpos| [rad]", obs[:, -1].abs(), "max"); upd("max hinge_qpos [rad]", obs[:, -2], "max")
        elif name == "Stack":
            upd("min |eef-cubeA| [m]", obs[:, -9:-6].norm(dim=1), "min"); upd("max cubeA z - start [m]", obs[:, -21] - o0[:, -21], "max"); upd("max step reward", rew, "max")
        elif name == "TwoArmLift":
            upd("min |g0-handle0| [m]", obs[:, -6:-3].norm(dim=1), "min"); upd("min |g1-handle1| [m]", obs[:, -3:].norm(dim=1), "min"); upd("max pot z - start [m]", obs[:, -23] - o0[:, -23], "max"); upd("max step reward", rew, "max")
    env.close()
    q = lambda v: "p10 %.3f p50 %.3f p90 %.3f" % tuple(np.percentile(v.cpu().numpy(), [10, 50, 90]))
    print(f"{run}: return {ret.mean().item():.1f} (logged last-50 {logged[-50:].mean():.1f})  start: " + (f"|eef-target| {o0[:, -3:].norm(dim=1).mean().item():.3f}" if name in ("Lift",) else "") )
    for k, v in st.items(): print(f"    {k:28s} {q(v)}")
    sys.stdout.flush()

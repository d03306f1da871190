#!/usr/bin/env python3
"""Transfer check over EVERY committed run of the families this package implements: the reference's 2020 policies -- trained against real robosuite +
MuJoCo -- rolled out deterministically in the batched CUDA env, next to the evaluation returns their own runs logged.
  python tools/eval_committed_runs.py export          (here: reads /root/reference/runs/*/.../{params.pkl,variant.json,progress.csv}, writes the git-ignored
                                                       oracle/_ref/policies/<run>.npz -- build products of the reference, like oracle/_ref/refpy.bin)
  python tools/eval_committed_runs.py run [EPISODES] [FILTER]   (GPU box: table on stdout)"""
import sys, os, glob, csv, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
OUT = os.path.join(ROOT, "oracle", "_ref", "policies")
FAMILIES = ("Lift", "Door", "Stack", "TwoArmLift", "PickPlaceCan", "PickPlaceMilk", "TwoArmPegInHole", "NutAssemblyRound", "TwoArmHandoff")
mode = sys.argv[1] if len(sys.argv) > 1 else "export"
if mode == "export":
    from robosuite_benchmark_b200.policy_io import load_params_pkl, mlp_weights
    os.makedirs(OUT, exist_ok=True)
    n = 0
    for run in sorted(os.listdir("/root/reference/runs")):
        if run.split("-")[0] not in FAMILIES:
            continue
        d = glob.glob(f"/root/reference/runs/{run}/*/params.pkl")
        if not d:
            continue
        base = os.path.dirname(d[0])
        w = mlp_weights(load_params_pkl(d[0])["trainer/policy"])
        variant = json.load(open(os.path.join(base, "variant.json")))
        rows = list(csv.DictReader(open(os.path.join(base, "progress.csv"))))
        logged = np.array([float(r["evaluation/Average Returns"]) for r in rows], np.float32)
        np.savez_compressed(os.path.join(OUT, run + ".npz"), logged=logged, env_kwargs=json.dumps(variant["eval_environment_kwargs"]),
                            **{k: v.astype(np.float32) for k, v in w.items()})
        n += 1
    print("exported", n, "runs to", OUT)
else:
    import torch
    from robosuite_benchmark_b200.rollout import evaluate_policy
    from robosuite_benchmark_b200.model import assets as _A
    for spec in filter(None, os.environ.get("RSB_EVAL_ASSET_MODULE", "").split(";")):         # e.g. "RETHINK_FINGER_STYLE=round1": module-level switches of model/assets.py
        k, v = spec.split("=", 1)
        try: v = json.loads(v)
        except ValueError: pass
        if isinstance(v, dict) and isinstance(getattr(_A, k, None), dict): getattr(_A, k).update(v)
        else: setattr(_A, k, v)
    for spec in filter(None, os.environ.get("RSB_EVAL_ASSET_OVERRIDES", "").split(";")):      # e.g. "Sawyer.grip_sign=[-1,1]": patch model/assets.py ROBOTS entries (what-if studies)
        k, v = spec.split("="); r, field = k.split(".")
        _A.ROBOTS[r][field] = json.loads(v)
    episodes = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    extra = json.loads(os.environ.get("RSB_EVAL_CONTROLLER_OVERRIDES", "{}"))          # e.g. '{"orientation_delta": "axis_angle"}' (applied through a temp controller json)
    print(f"{'run':46s} {'here: mean +- se':>18s} {'max':>7s} | {'logged: last-50 mean':>20s} {'min..max of last 50':>20s} {'run max':>8s} | ratio")
    fam = {}
    for f in sorted(glob.glob(os.path.join(OUT, "*.npz"))):
        run = os.path.basename(f)[:-4]
        if filt and filt not in run:
            continue
        d = dict(np.load(f)); logged = d.pop("logged"); cfg = json.loads(str(d.pop("env_kwargs")))
        if extra:
            from robosuite_benchmark_b200.controllers import load_controller_config
            cc = load_controller_config(default_controller=cfg["controller"])
            if "kp" in extra: cc.pop("kv", None)
            cc.update(extra)
            p = f"/tmp/ctrl_{run}.json"; json.dump(cc, open(p, "w")); cfg["controller"] = p
        try:
            ret = evaluate_policy(cfg, d, num_episodes=episodes, seed=17)
        except Exception as e:                                                      # a family the model builder does not cover
            print(f"{run:46s} not run: {type(e).__name__}: {str(e)[:90]}"); continue
        l50 = logged[-50:]
        print(f"{run:46s} {ret.mean():9.1f} +- {ret.std() / np.sqrt(len(ret)):5.1f} {ret.max():7.1f} | {l50.mean():20.1f} {l50.min():9.1f}..{l50.max():<9.1f} {logged.max():8.1f} | {ret.mean() / l50.mean():.2f}", flush=True)
        fam.setdefault(run.rsplit("-SEED", 1)[0], []).append((ret.mean(), l50.mean(), ret.max(), logged.max()))
    print("\nper family (mean over its seeds): here / logged last-50 mean; best episode here / best logged evaluation")
    for k, v in fam.items():
        v = np.array(v)
        print(f"  {k:40s} {v[:, 0].mean():7.1f} / {v[:, 1].mean():7.1f} = {v[:, 0].mean() / v[:, 1].mean():.2f}    {v[:, 2].max():7.1f} / {v[:, 3].max():7.1f}")

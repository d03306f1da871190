#!/usr/bin/env python3
"""tools/eval_committed_runs.py without a GPU: the committed policies rolled out in the fp64 CPU ORACLE (oracle/rsb_oracle.c), one episode per worker process.
Slower (about 1 s per 500-step episode and core) but free of GPU budget: what-if studies of the model DATA (assets, controller config) can run here, since the
oracle and the CUDA kernels are held to each other by the parity tests.
  python tools/eval_committed_runs_cpu.py [FILTER] [EPISODES] ; same RSB_EVAL_* environment switches as the GPU tool."""
import sys, os, glob, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
from multiprocessing import Pool


def _setup():
    from robosuite_benchmark_b200.model import assets as _A
    for spec in filter(None, os.environ.get("RSB_EVAL_ASSET_MODULE", "").split(";")):
        k, v = spec.split("=", 1)
        try: v = json.loads(v)
        except ValueError: pass
        if isinstance(v, dict) and isinstance(getattr(_A, k, None), dict): getattr(_A, k).update(v)
        else: setattr(_A, k, v)
    for spec in filter(None, os.environ.get("RSB_EVAL_ASSET_OVERRIDES", "").split(";")):
        k, v = spec.split("="); r, field = k.split(".")
        _A.ROBOTS[r][field] = json.loads(v)


def episode(job):
    f, ep, stats = job
    _setup()
    from oracle.oracle import OracleEnv
    from robosuite_benchmark_b200.controllers import load_controller_config
    from robosuite_benchmark_b200.model.tasks import build_task
    from robosuite_benchmark_b200.policy_io import DeterministicPolicy
    d = dict(np.load(f)); d.pop("logged"); cfg = json.loads(str(d.pop("env_kwargs")))
    pol = DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()})
    cc = load_controller_config(default_controller=cfg["controller"])
    extra = json.loads(os.environ.get("RSB_EVAL_CONTROLLER_OVERRIDES", "{}"))
    if "kp" in extra: cc.pop("kv", None)
    cc.update(extra)
    kw = {k: cfg[k] for k in ("env_configuration",) if k in cfg}
    m, t = build_task(cfg["env_name"], cfg["robots"], cc, horizon=cfg.get("horizon", 500), ignore_done=True, **kw)
    nc, ne = (int(x) for x in os.environ.get("RSB_EVAL_LIMITS", "48,160").split(","))       # generous limits: the oracle is not bound by shared memory
    orc = OracleEnv(m, t, ncon_max=nc, nefc_max=ne)
    o, ret = orc.reset(seed=17, env_id=ep), 0.0
    o0 = np.array(o); extra_stats = {}
    for _ in range(500):
        o, r, _ = orc.step(pol(np.asarray(o)))
        ret += r
        if stats and cfg["env_name"] == "Door":
            extra_stats["min_handle_dist"] = min(extra_stats.get("min_handle_dist", 9), float(np.linalg.norm(o[-5:-2])))
            extra_stats["max_handle_q"] = max(extra_stats.get("max_handle_q", 0), abs(float(o[-1]))); extra_stats["max_hinge"] = max(extra_stats.get("max_hinge", 0), float(o[-2]))
        if stats and cfg["env_name"] == "TwoArmLift":
            extra_stats["max_pot_dz"] = max(extra_stats.get("max_pot_dz", -9), float(o[-23] - o0[-23]))
            extra_stats["min_g0h0"] = min(extra_stats.get("min_g0h0", 9), float(np.linalg.norm(o[-6:-3]))); extra_stats["min_g1h1"] = min(extra_stats.get("min_g1h1", 9), float(np.linalg.norm(o[-3:])))
            extra_stats["max_r"] = max(extra_stats.get("max_r", 0), r)
            extra_stats["final_pot_dz"] = float(o[-23] - o0[-23])
        if stats and cfg["env_name"] == "Stack":
            extra_stats["min_gA"] = min(extra_stats.get("min_gA", 9), float(np.linalg.norm(o[-9:-6]))); extra_stats["max_A_dz"] = max(extra_stats.get("max_A_dz", -9), float(o[-21] - o0[-21])); extra_stats["max_r"] = max(extra_stats.get("max_r", 0), r)
        if stats and cfg["env_name"] == "Lift":
            extra_stats["max_r"] = max(extra_stats.get("max_r", 0), r)
            extra_stats["min_cube_dist"] = min(extra_stats.get("min_cube_dist", 9), float(np.linalg.norm(o[-3:])))
            extra_stats["max_cube_dz"] = max(extra_stats.get("max_cube_dz", -9), float(o[-8] - o0[-8]))
    return os.path.basename(f)[:-4], ret, extra_stats


if __name__ == "__main__":
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    episodes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    stats = os.environ.get("RSB_EVAL_STATS") == "1"
    files = [f for f in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "policies", "*.npz"))) if all(x in os.path.basename(f) for x in filt.split("+"))]
    jobs = [(f, e, stats) for f in files for e in range(episodes)]
    with Pool(int(os.environ.get("RSB_EVAL_PROCS", os.cpu_count()))) as p:
        res = p.map(episode, jobs, chunksize=1)
    fam = {}
    for f in files:
        run = os.path.basename(f)[:-4]; logged = np.load(f)["logged"]
        r = np.array([x[1] for x in res if x[0] == run]); l50 = logged[-50:]
        line = f"{run:46s} {r.mean():8.1f} +- {r.std() / np.sqrt(len(r)):5.1f} max {r.max():6.1f} | logged {l50.mean():6.1f} (max {logged.max():5.1f}) | {r.mean() / l50.mean():.2f}"
        if stats:
            ks = sorted({k for x in res if x[0] == run for k in x[2]})
            line += "  " + " ".join(f"{k} p50 {np.median([x[2][k] for x in res if x[0] == run]):.3f}" for k in ks)
        print(line, flush=True)
        fam.setdefault(run.rsplit("-SEED", 1)[0], []).append((r.mean(), l50.mean(), r.max(), logged.max()))
    for k, v in fam.items():
        v = np.array(v)
        print(f"  {k:40s} {v[:, 0].mean():7.1f} / {v[:, 1].mean():7.1f} = {v[:, 0].mean() / v[:, 1].mean():.2f}    {v[:, 2].max():7.1f} / {v[:, 3].max():7.1f}")

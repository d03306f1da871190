#!/usr/bin/env python3
"""Grid search over model DATA switches (model/assets.py) scored by how well the committed policies' returns in the fp64 CPU oracle match what their own runs
logged (mean of the last 5 logged evaluation returns; runs whose policy is the last epoch's).  loss = mean over runs of (log(here + 10) - log(logged + 10))^2.
  python tools/fit_assets_cpu.py FILTER EPISODES 'SPEC' ['SPEC' ...]        SPEC as in RSB_EVAL_ASSET_MODULE"""
import sys, os, glob, json, subprocess
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
filt, episodes, specs = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
files = [f for f in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "policies", "*.npz"))) if all(x in os.path.basename(f) for x in filt.split("+"))]
ref = {os.path.basename(f)[:-4]: float(np.load(f)["logged"][-5:].mean()) for f in files}
print("reference (mean of the last 5 logged evaluation returns):", {k.split("-SEED")[0][-12:] + k.split("SEED")[1]: round(v) for k, v in ref.items()})
for spec in specs:
    env = dict(os.environ, RSB_EVAL_ASSET_MODULE=spec)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "eval_committed_runs_cpu.py"), filt, str(episodes)], env=env, capture_output=True, text=True).stdout
    here = {l.split()[0]: float(l.split()[1]) for l in out.splitlines() if l and not l.startswith(" ") and "SEED" in l.split()[0]}
    loss = np.mean([(np.log(here[k] + 10) - np.log(ref[k] + 10)) ** 2 for k in ref])
    print(f"loss {loss:6.3f}  {spec:90s} " + " ".join(f"{here[k]:4.0f}" for k in ref), flush=True)

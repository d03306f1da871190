#!/usr/bin/env python3
"""Put a learning curve measured here next to the one the reference's committed run logged (run HERE: needs /root/reference/runs).
  python tools/merge_curve.py gpurun_out/r2_train_Lift_progress.csv Lift-Panda-OSC-POSE-SEED17 profiles/r2_train_lift_seed17_num_envs5.csv"""
import csv, glob, sys
mine = list(csv.DictReader(open(sys.argv[1])))
ref = list(csv.DictReader(open(glob.glob(f"/root/reference/runs/{sys.argv[2]}/*/progress.csv")[0])))
keep = ["Epoch", "evaluation/Average Returns", "exploration/Average Returns", "trainer/QF1 Loss", "trainer/Policy Loss", "trainer/Alpha",
        "time/exploration sampling (s)", "time/evaluation sampling (s)", "time/training (s)", "time/epoch (s)"]
with open(sys.argv[3], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep + ["reference run: evaluation/Average Returns", "reference run: time/epoch (s)"])
    for i, r in enumerate(mine):
        w.writerow([r[k] for k in keep] + ([ref[i]["evaluation/Average Returns"], ref[i]["time/epoch (s)"]] if i < len(ref) else ["", ""]))
import numpy as np
a = np.array([float(r["evaluation/Average Returns"]) for r in mine]); b = np.array([float(r["evaluation/Average Returns"]) for r in ref[:len(mine)]])
q = len(a) // 4
print("mean evaluation return per quarter of the run, here vs logged:", [(round(a[i * q:(i + 1) * q].mean(), 1), round(b[i * q:(i + 1) * q].mean(), 1)) for i in range(4)],
      "| wall time here %.0f s, logged %.0f s" % (sum(float(r["time/epoch (s)"]) for r in mine), sum(float(r["time/epoch (s)"]) for r in ref[:len(mine)])))

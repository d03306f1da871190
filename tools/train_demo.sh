#!/bin/bash
# Train SAC on Lift-Panda-OSC_POSE with the committed variant (shortened) on the batched backend and print the learning curve (run under gpurun)
mkdir -p gpurun_out
timeout ${2:-1400} python -m robosuite_benchmark_b200.train --variant ${3:-tools/variant_Lift_short.json} --seed 17 --num_envs ${1:-5} --log_dir gpurun_out/train > gpurun_out/train_lift.log 2>&1
echo "train rc=$?"; tail -3 gpurun_out/train_lift.log
f=$(find gpurun_out/train -name progress.csv | head -1); cp "$f" gpurun_out/train_progress.csv
python - <<'PY'
import csv
rows = list(csv.DictReader(open("gpurun_out/train_progress.csv")))
print(len(rows), "epochs")
for r in rows[::10] + rows[-1:]:
    print(r["Epoch"], "eval ret %.1f expl ret %.1f  time/epoch %.2f s (expl %.2f eval %.2f train %.2f)" % (float(r["evaluation/Average Returns"]), float(r["exploration/Average Returns"]),
          float(r["time/epoch (s)"]), float(r["time/exploration sampling (s)"]), float(r["time/evaluation sampling (s)"]), float(r["time/training (s)"])))
PY

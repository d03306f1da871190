#!/bin/bash
# round 2, trip ah: learning curves from scratch under the defaults that the transfer study established (reference data flow: 5 envs, committed variants shortened)
mkdir -p gpurun_out
for v in DoorJV Lift; do
  rm -rf gpurun_out/train_$v
  timeout 700 python -m robosuite_benchmark_b200.train --variant tools/variant_${v}_short.json --seed 17 --num_envs 5 --log_dir gpurun_out/train_$v > gpurun_out/train_$v.log 2>&1
  echo "$v rc=$?"
  f=$(find gpurun_out/train_$v -name progress.csv | head -1); cp "$f" gpurun_out/r2_train_${v}_progress.csv
  python - "$v" <<'PY'
import csv, sys
rows = list(csv.DictReader(open(f"gpurun_out/r2_train_{sys.argv[1]}_progress.csv")))
print(len(rows), "epochs")
for r in rows[::10] + rows[-1:]:
    print(r["Epoch"], "eval ret %.1f expl ret %.1f  time/epoch %.2f s" % (float(r["evaluation/Average Returns"]), float(r["exploration/Average Returns"]), float(r["time/epoch (s)"])))
PY
  rm -rf gpurun_out/train_$v
done

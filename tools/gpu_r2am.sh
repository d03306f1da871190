#!/bin/bash
# round 2, trip am: final regression after the Door fit (push door, XML hinge values): GPU suite, transfer table (96 episodes per policy), default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/eval_committed_runs.py run 96 2>&1 | grep -v Warn > gpurun_out/r2_policy_transfer_all.txt; tail -15 gpurun_out/r2_policy_transfer_all.txt
timeout 400 python bench.py > gpurun_out/r2_bench_n1_am.json 2> gpurun_out/bench_am.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_am.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], {k: round(v["steps_per_s"]) for k, v in d["other_configs"].items()}, round(d["trained_policy"]["steps_per_s"]), d["sac"]["b128"]["updates_per_s"])
PY

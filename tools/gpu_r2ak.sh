#!/bin/bash
# round 2, trip ak: final single-GPU regression (whole GPU suite + default bench line with the trained-policy leg)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2_bench_n1_ak.json 2> gpurun_out/bench_ak.err; tail -3 gpurun_out/bench_ak.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_ak.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["gpu_launches"])
print(d["trained_policy"])
PY

#!/bin/bash
# One GPU round trip (run under gpurun): GPU tests, a short bench, the launch list and one full ncu capture of k_step.
# usage: tools/gpu_round.sh TAG [quick]
TAG=${1:-x}; MODE=${2:-full}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py --steps 20 --warmup 5 --no-cpu --no-sac > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_$TAG.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), d.get("kernel_info"), d.get("clocks"))
PY
if [ "$MODE" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --no-cpu --no-sac > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
fi

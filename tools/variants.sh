#!/bin/bash
# bench k_step under env-var variants (run under gpurun): tools/variants.sh "RSB_EPB=28" "RSB_LOCKSTEP=0" ...
mkdir -p gpurun_out
for v in "$@"; do
  env $v python bench.py --steps 20 --warmup 5 --no-cpu --no-sac > gpurun_out/var.log 2>&1
  python - "$v" <<'PY'
import json, sys
for l in open("gpurun_out/var.log"):
    if l.startswith("{"):
        d = json.loads(l); print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), d.get("kernel_info"))
PY
done

#!/bin/bash
# k_step time vs number of CTAs (28 envs each): is instruction fetch a per-SM or a chip-wide limit?  (run under gpurun)
mkdir -p gpurun_out
for k in "$@"; do
  n=$((28 * k))
  RSB_EPB=28 python bench.py --steps 20 --warmup 5 --no-cpu --no-sac --envs $n > gpurun_out/var.log 2>&1
  python - "$k" <<'PY'
import json, sys
for l in open("gpurun_out/var.log"):
    if l.startswith("{"):
        d = json.loads(l); print("ctas", sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "kernel_ms", round(d["roofline"]["kernel_ms"], 3))
PY
done

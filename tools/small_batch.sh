#!/bin/bash
# k_step latency for small batches under 16- vs 32-lane groups (run under gpurun)
for n in 8 1024 2048; do for l in 16 32; do
  RSB_LANES=$l python bench.py --steps 20 --warmup 5 --no-cpu --no-sac --envs $n > gpurun_out/var.log 2>&1
  python - "$n" "$l" <<'PY'
import json, sys
for line in open("gpurun_out/var.log"):
    if line.startswith("{"):
        d = json.loads(line); print("envs", sys.argv[1], "lanes", sys.argv[2], "kernel_ms", round(d["roofline"]["kernel_ms"], 3), "value", round(d["value"]), d["kernel_info"])
PY
done; done

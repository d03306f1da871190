#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_r2k.json 2> gpurun_out/bench_n8_r2k.err; echo "bench8 rc=$?"; tail -3 gpurun_out/bench_n8_r2k.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_n8_r2k.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3))
        print("sac", {k: (round(v["updates_per_s"]) if isinstance(v, dict) else v) for k, v in d["sac"].items()})
        print("train", json.dumps(d["train"])[:900])
        print("others", {k: round(v["steps_per_s_all_gpus"]) for k, v in d["other_configs"].items()})
PY

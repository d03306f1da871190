#!/bin/bash
# Round-2 final multi-GPU trip (run with gpurun --gpus 8): 2-GPU test, then the driver's scaling commands at N = 2 and N = 8, train mode at N = 8.
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_multi.py -q -rs -s > gpurun_out/pytest_multi_final.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_multi_final.log | tail -4
for n in 2 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_n${n}_final.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --mode train --steps 8 --warmup 2 > gpurun_out/r2_train_n8_final.json 2> gpurun_out/train_n8.err; echo "train n=8 rc=$?"
python - <<'PY'
import json
def first(f):
    return json.loads([l for l in open(f) if l.startswith("{")][0])
for n in (2, 8):
    d = first(f"gpurun_out/r2_bench_n{n}_final.json")
    print(f"N={n} value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "trunc", d["truncation"])
    print("  sac", {k: (round(v["updates_per_s"]), round(v["us_per_update"], 1)) for k, v in d["sac"].items() if isinstance(v, dict)})
    print("  train", {k: round(d["train"][k]) for k in ("env_steps_per_s", "updates_per_s_in_loop", "sampling_env_steps_per_s", "training_updates_per_s")})
    print("  others", {k: round(v["steps_per_s_all_gpus"]) for k, v in d["other_configs"].items()})
t = first("gpurun_out/r2_train_n8_final.json"); print("train-mode N=8", round(t["value"]), t["ms_per_step"], round(t["train"]["sampling_env_steps_per_s"]), round(t["train"]["training_updates_per_s"]))
PY

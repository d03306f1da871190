#!/bin/bash
# Round-2 final single-GPU trip: whole GPU suite, smoke, default bench line (the driver's command), train-mode line, per-config lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --mode train --steps 8 --warmup 2 > gpurun_out/r2_train_n1_final.json 2> gpurun_out/train_final.err; echo "train rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_reference_arm.json 2> gpurun_out/reference.err; echo "reference rc=$?"
for c in door stack twoarmlift; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 --quick --no-sac > gpurun_out/r2_bench_${c}_n1.json 2> gpurun_out/bench_$c.err; echo "$c rc=$?"; done
python - <<'PY'
import json
def first(f):
    return json.loads([l for l in open(f) if l.startswith("{")][0])
d = first("gpurun_out/r2_bench_n1_final.json")
print("bench", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "full-episode", round(d["full_episode"]["steps_per_s"]), "cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"])
print("sac", {k: (round(v["updates_per_s"]), round(v["us_per_update"], 1)) for k, v in d["sac"].items() if isinstance(v, dict)})
print("train-in-bench", {k: d["train"][k] for k in ("env_steps_per_s", "updates_per_s_in_loop", "sampling_env_steps_per_s", "training_updates_per_s")})
t = first("gpurun_out/r2_train_n1_final.json"); print("train-mode", round(t["value"]), t["ms_per_step"], round(t["train"]["sampling_env_steps_per_s"]), round(t["train"]["training_updates_per_s"]), t["train"]["truncation"])
for c in ("door", "stack", "twoarmlift"):
    x = first(f"gpurun_out/r2_bench_{c}_n1.json"); print(c, round(x["value"]), round(x["e2e"]["value"]), x["roofline"]["algorithmic_bytes_per_env_step"], x["truncation"])
print("reference arm", first("gpurun_out/r2_reference_arm.json")["value"])
PY

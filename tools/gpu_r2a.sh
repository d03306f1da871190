#!/bin/bash
# Round-2 GPU trip A (run under gpurun): all GPU tests (not stopping at the first failure), smoke, default bench, train-mode bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_collector.py::test_no_contact_or_row_truncation_over_a_full_episode > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2a.log
timeout 900 python -m pytest tests/test_gpu_collector.py -q -k "truncation_over" > gpurun_out/pytest_r2a_trunc.log 2>&1; echo "pytest-trunc rc=$?"; tail -8 gpurun_out/pytest_r2a_trunc.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r2a.log
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2a.err
timeout 600 python bench.py --mode train --steps 8 --warmup 2 > gpurun_out/train_r2a.json 2> gpurun_out/train_r2a.err; echo "train rc=$?"; tail -3 gpurun_out/train_r2a.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_r2a.json", "gpurun_out/train_r2a.json"):
    try:
        for l in open(f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3))
                for k in ("full_episode", "other_configs", "train", "sac"):
                    if d.get(k): print(" ", k, json.dumps(d[k])[:1500])
    except Exception as e:
        print(f, "unreadable", e)
PY

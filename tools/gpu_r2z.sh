#!/bin/bash
# round 2, trip z: after the orientation-delta convention + Lift limits (18, 62): whole GPU suite, default bench, policy-driven limits sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
(RSB_SWEEP_POLICY=1 timeout 600 python tools/limits_sweep.py Lift Panda OSC_POSE 18,62 24,80 32,96) 2>&1 | grep -v Warning | tee gpurun_out/r2_limits_sweep_lift_policy.txt
timeout 300 python tools/eval_committed_policy.py run > gpurun_out/r2_eval_committed_policy.txt 2>&1; tail -2 gpurun_out/r2_eval_committed_policy.txt
timeout 600 python bench.py > gpurun_out/r2_bench_n1_z.json 2> gpurun_out/bench_z.err; cut -c1-600 gpurun_out/r2_bench_n1_z.json

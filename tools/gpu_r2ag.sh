#!/bin/bash
# round 2, trip ag: after the evidence-driven model changes (JV law, Rethink sign, recalled door, pot, base spacing): whole GPU suite, transfer table, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 1500 python tools/eval_committed_runs.py run 256 2>&1 | grep -v Warn > gpurun_out/r2_policy_transfer_all.txt; tail -16 gpurun_out/r2_policy_transfer_all.txt
timeout 300 python tools/diag_transfer_stages.py Lift-Sawyer-OSC 128 2>&1 | grep -v Warn | tee gpurun_out/r2_transfer_stages_sawyer.txt
timeout 600 python bench.py > gpurun_out/r2_bench_n1_ag.json 2> gpurun_out/bench_ag.err; cut -c1-300 gpurun_out/r2_bench_n1_ag.json

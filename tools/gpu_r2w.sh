#!/bin/bash
# round 2, trip w: after the OSC orientation-delta convention change -- whole GPU suite, committed-policy transfer record, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/diag_policy_transfer2.py > gpurun_out/r2_policy_transfer.txt 2>&1; tail -8 gpurun_out/r2_policy_transfer.txt
timeout 300 python tools/eval_committed_policy.py run > gpurun_out/r2_eval_committed_policy.txt 2>&1; tail -6 gpurun_out/r2_eval_committed_policy.txt
timeout 600 python bench.py > gpurun_out/r2_bench_n1_w.json 2> gpurun_out/bench_w.err; tail -c 1500 gpurun_out/r2_bench_n1_w.json

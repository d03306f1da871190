#!/bin/bash
# Round-end GPU pass (run under gpurun): the whole GPU test suite, the default bench, the launch list of a short bench, and one full ncu capture of
# the SAC update's tensor-core GEMM.  usage: tools/gpu_final.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_full_$TAG.log 2>&1; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32 -s 400 -c 3 -o gpurun_out/prof_tc_gemm_$TAG -f python tools/sac_rate.py --updates 20 > gpurun_out/ncu_tc_$TAG.log 2>&1; echo "ncu rc=$?"
tail -1 gpurun_out/bench_full_$TAG.log | cut -c1-400

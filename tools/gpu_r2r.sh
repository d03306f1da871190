#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_sac.py tests/test_gpu_train_loop.py tests/test_gpu_collector.py -q -x -k "not truncation_over" > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_r2r.log
for tma in 1 0; do RSB_GEMM_TMA=$tma timeout 300 python tools/sac_rate.py 2>&1 | grep "^b128 \|^b4096 " | cut -c1-110 | sed "s/^/TMA=$tma /"; done | tee gpurun_out/sac_rate_r2r.log
RSB_PDL=0 timeout 300 python tools/sac_timeline.py 4096 > gpurun_out/sac_timeline_b4096_nopdl_tma.txt 2>&1; grep -v Warn gpurun_out/sac_timeline_b4096_nopdl_tma.txt | tail -30

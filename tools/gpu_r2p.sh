#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sac.py tests/test_gpu_train_loop.py -q -x > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2p.log
for tma in 1 0; do RSB_GEMM_TMA=$tma timeout 300 python tools/sac_rate.py 2>&1 | grep "^b128 \|^b4096 " | cut -c1-110 | sed "s/^/TMA=$tma /"; done | tee gpurun_out/sac_rate_r2p.log
timeout 300 python tools/diag_tc_gemm.py --time 2>&1 | grep "^time" | tee gpurun_out/gemm_time_r2p.log

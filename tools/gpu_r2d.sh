#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sac.py tests/test_gpu_tc_gemm.py tests/test_gpu_reference_dropin.py -q -rs -s > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2d.log | tail -12
for pdl in 1 0; do RSB_PDL=$pdl timeout 300 python tools/sac_rate.py 2>&1 | tail -4 | sed "s/^/PDL=$pdl /"; done | tee gpurun_out/sac_rate_r2d.log

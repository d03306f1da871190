#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sac.py tests/test_gpu_train_loop.py tests/test_gpu_reference_dropin.py -q -rs -s -x > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2g.log | tail -25
for pdl in 1 0; do RSB_PDL=$pdl timeout 300 python tools/sac_rate.py 2>&1 | grep "^b128\|^b4096" | cut -c1-110 | sed "s/^/PDL=$pdl /"; done | tee gpurun_out/sac_rate_r2g.log

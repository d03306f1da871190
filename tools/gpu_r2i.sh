#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sac.py tests/test_gpu_train_loop.py -q -rs -s -x > gpurun_out/pytest_r2i.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2i.log | tail -12
timeout 300 python tools/sac_rate.py 2>&1 | grep "^b128\|^b4096" | cut -c1-110 | tee gpurun_out/sac_rate_r2i.log
RSB_PDL=0 timeout 300 python tools/sac_timeline.py 128 > gpurun_out/sac_timeline_b128_nopdl.txt 2>&1; grep -v Warn gpurun_out/sac_timeline_b128_nopdl.txt | tail -30

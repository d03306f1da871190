#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py -q -rs -x > gpurun_out/pytest_r2o.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2o.log | tail -25

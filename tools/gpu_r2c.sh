#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collector.py tests/test_gpu_sac.py tests/test_gpu_reference_dropin.py -q -rs -s -k "truncation_over or tf32_operand or dropin" > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2c.log | tail -12
timeout 900 python tools/diag_iters.py > gpurun_out/iters_r2c.log 2>&1; echo "iters rc=$?"; cat gpurun_out/iters_r2c.log | cut -c1-400

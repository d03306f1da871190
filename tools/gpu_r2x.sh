#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/limits_sweep.py Lift Panda OSC_POSE 16,64 18,64 20,64 16,72 20,72 20,80 24,80 2>&1 | grep -v Warning | tee gpurun_out/r2_limits_sweep_lift.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12

#!/bin/bash
mkdir -p gpurun_out
L="16,64 18,63 18,62 20,62 20,60"
(timeout 600 python tools/limits_sweep.py Lift Panda OSC_POSE $L; RSB_SWEEP_POLICY=1 timeout 600 python tools/limits_sweep.py Lift Panda OSC_POSE $L; timeout 600 python tools/limits_sweep.py Lift Panda JOINT_VELOCITY $L) 2>&1 | grep -v Warning | tee gpurun_out/r2_limits_sweep_lift2.txt

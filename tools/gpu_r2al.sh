#!/bin/bash
# k_step under the trained policy's contact load: sensitivity to the Newton iteration cap, the CTA-wide solver lockstep, envs per block and lanes per env
mkdir -p gpurun_out
run() { echo "== $1"; env $1 RSB_SWEEP_POLICY=1 timeout 200 python tools/limits_sweep.py Lift Panda OSC_POSE 18,62 2>&1 | grep -v Warn | sed 's/first overflow.*| ms/| ms/'; }
(run "RSB_NOP=1"; run "RSB_SOLVER_ITERS=4"; run "RSB_SOLVER_ITERS=2"; run "RSB_LOCKSTEP=0"; run "RSB_EPB=14"; run "RSB_LANES=32") | tee gpurun_out/r2_policy_load_knobs.txt

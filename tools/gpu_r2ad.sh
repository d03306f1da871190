#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; RSB_EVAL_ASSET_MODULE="$1" timeout 600 python tools/eval_committed_runs.py run 128 Door-Panda 2>&1 | grep -v Warn | grep "^Door\|^  \|not run"; }
(run 'DOOR_STYLE=robosuite_recalled'
 run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"bolt": false}'
 run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"inertia": [0.0005, 0.0005, 0.0005]}') | tee gpurun_out/r2_policy_transfer_door.txt

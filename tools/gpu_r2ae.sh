#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; RSB_EVAL_ASSET_MODULE="$1" timeout 600 python tools/eval_committed_runs.py run 96 Door-Panda 2>&1 | grep -v Warn | grep "^  \|not run"; }
(run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"bolt_pos": [-0.04, 0.044, 0], "bolt_size": [0.03, 0.01, 0.008]}'
 run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"bolt_pos": [-0.075, 0.10, 0], "bolt_size": [0.075, 0.015, 0.02]}'
 run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"bolt_pos": [-0.04, 0.044, 0], "bolt_size": [0.03, 0.01, 0.008], "inertia": [0.0005, 0.0005, 0.0005]}'
 run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"bolt": false, "inertia": [0.0005, 0.0005, 0.0005]}'
 run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"bolt": false, "inertia": [0.005, 0.005, 0.005]}'
 run 'DOOR_STYLE=robosuite_recalled;DOOR_LATCH={"bolt": false, "stiffness": 0.5}'
 ) | tee gpurun_out/r2_policy_transfer_door2.txt

#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $2 :: $1"; RSB_EVAL_ASSET_MODULE="$1" timeout 600 python tools/eval_committed_runs.py run 96 $2 2>&1 | grep -v Warn | grep "^  \|not run"; }
(P='POT={"thickness": 0.025, "handle_z": 0.06, "bar_half": 0.055, "side_bars": true}'
 run "$P" TwoArmLift-PandaPanda
 run "$P;TWO_ARM_BASE_Y=0.56" TwoArmLift-PandaPanda
 run "TWO_ARM_BASE_Y=0.56" TwoArmLift-PandaPanda
 run 'POT={"side_bars": true}' TwoArmLift-PandaPanda
 run 'POT={"thickness": 0.025}' TwoArmLift-PandaPanda
 run "" Door-Sawyer
 ) | tee gpurun_out/r2_policy_transfer_pot.txt

#!/bin/bash
mkdir -p gpurun_out
for style in narrow_tall round1; do
  echo "== Rethink fingers: $style (grip sign fixed)"
  RSB_EVAL_ASSET_MODULE="RETHINK_FINGER_STYLE=$style" timeout 600 python tools/eval_committed_runs.py run 128 Sawyer-OSC 2>&1 | grep -v Warn | grep "^  \|not run"
done | tee gpurun_out/r2_policy_transfer_sawyer.txt

#!/bin/bash
# 100-epoch SAC runs of further committed variants (run under gpurun); learning curves copied to gpurun_out/train_progress_<name>.csv
for v in DoorJV LiftSawyer TwoArmLift; do
  rm -rf gpurun_out/train
  echo "=== $v"; tools/train_demo.sh 5 600 tools/variant_${v}_short.json | tail -13
  cp gpurun_out/train_progress.csv gpurun_out/train_progress_$v.csv
done

#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/diag_transfer_stages.py Door-Panda-OSC 128; timeout 300 python tools/diag_transfer_stages.py Door-Panda-JOINT 128; timeout 200 python tools/diag_transfer_stages.py Lift-Sawyer-OSC-POSE-SEED17 128;  timeout 200 python tools/diag_transfer_stages.py Lift-Panda-OSC-POSE-SEED17 128; timeout 200 python tools/diag_transfer_stages.py TwoArmLift-PandaPanda-OSC-POSE-SEED17 128; timeout 200 python tools/diag_transfer_stages.py Stack-Sawyer-OSC-POSE-SEED17 128) 2>&1 | grep -v Warn | tee gpurun_out/r2_transfer_stages.txt

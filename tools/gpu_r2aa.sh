#!/bin/bash
# which JOINT_VELOCITY law were the committed JV policies trained under?  (tools/eval_committed_runs.py with controller overrides)
mkdir -p gpurun_out
for ov in '{"kp": 4.0, "ki_ratio": 0, "kd_ratio": 0, "kp_scale_by_actuator_range": false}' \
          '{"kp": 3.0, "ki_ratio": 0, "kd_ratio": 0, "kp_scale_by_actuator_range": false}' \
          '{"kp": 3.0, "ki_ratio": 0, "kd_ratio": 0, "kp_scale_by_actuator_range": true}' \
          '{"kp": 4.0, "ki_ratio": 0, "kd_ratio": 0, "kp_scale_by_actuator_range": false, "velocity_limits": null}' \
          '{"kp": 4.0, "ki_ratio": 0, "kd_ratio": 0, "kp_scale_by_actuator_range": false, "output_max": 1.0, "output_min": -1.0}'; do
  echo "== overrides $ov"
  RSB_EVAL_CONTROLLER_OVERRIDES="$ov" timeout 500 python tools/eval_committed_runs.py run 128 Panda-JOINT-VELOCITY 2>&1 | grep -v Warn | grep "^  \|not run" 
done | tee gpurun_out/r2_policy_transfer_jv.txt

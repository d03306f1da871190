#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_collector.py tests/test_gpu_sac.py -q -rs -s -k "policy_rollout or policy_forward or fused_ring or tf32_operand or graph_update" > gpurun_out/pytest_r2n.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2n.log | tail -14
timeout 600 python bench.py --mode train --steps 8 --warmup 2 > gpurun_out/train_r2n.json 2> gpurun_out/train_r2n.err; echo "train rc=$?"; python -c "
import json
t=json.loads([l for l in open('gpurun_out/train_r2n.json') if l.startswith('{')][0]); print(round(t['value']), t['ms_per_step'], t['train']['phase_s'], round(t['train']['sampling_env_steps_per_s']), round(t['train']['training_updates_per_s']))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_policy_act -c 4 python bench.py --mode train --steps 1 --warmup 1 --train-steps-per-epoch 5 --updates-per-epoch 10 2>&1 | grep -A3 "k_policy_act" | grep "gpu__time" | head -4

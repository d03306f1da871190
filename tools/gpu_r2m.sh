#!/bin/bash
# Round-2 profiling trip (1 GPU): default bench line, train-mode line, launch lists (bench + train loop) and ncu --set full captures (kept under 64 MiB in total).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2m.err
timeout 600 python bench.py --mode train --steps 8 --warmup 2 > gpurun_out/train_r2m.json 2> gpurun_out/train_r2m.err; echo "train rc=$?"
# launch lists (per-launch times under ncu are serialised / cold-cache: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --quick > gpurun_out/ncu_list_bench.log 2>&1; echo "ncu list bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_train.csv python bench.py --mode train --steps 1 --warmup 1 --train-steps-per-epoch 5 --updates-per-epoch 10 > gpurun_out/ncu_list_train.log 2>&1; echo "ncu list train rc=$?"
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 110 -c 1 -o gpurun_out/r2_kstep_lift python bench.py --steps 3 --warmup 3 --no-cpu --quick --no-sac > gpurun_out/ncu_kstep.log 2>&1; echo "ncu kstep rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_policy_act|k_path_stats" -c 2 -o gpurun_out/r2_collect python bench.py --mode train --steps 1 --warmup 1 --train-steps-per-epoch 5 --updates-per-epoch 10 > gpurun_out/ncu_collect.log 2>&1; echo "ncu collect rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_gemm_tf32|k_q_losses|k_policy_head|k_sac_begin|k_adam" -s 203 -c 22 -o gpurun_out/r2_sac_b128 python tools/sac_timeline.py 128 > gpurun_out/ncu_sac.log 2>&1; echo "ncu sac rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out

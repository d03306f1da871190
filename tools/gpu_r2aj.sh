#!/bin/bash
# round 2, trip aj (8 GPUs): final scaling lines after the model changes
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 > gpurun_out/r2_bench_n8_aj.json 2> gpurun_out/bench_n8_aj.err; cut -c1-250 gpurun_out/r2_bench_n8_aj.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 8 --mode train > gpurun_out/r2_train_n8_aj.json 2> gpurun_out/train_n8_aj.err; cut -c1-250 gpurun_out/r2_train_n8_aj.json

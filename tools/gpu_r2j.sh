#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py -q -rs -s -x > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2j.log | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --mode train --steps 6 --warmup 2 > gpurun_out/train_n2_r2j.json 2> gpurun_out/train_n2_r2j.err; echo "train rc=$?"; tail -3 gpurun_out/train_n2_r2j.err; cut -c1-1500 gpurun_out/train_n2_r2j.json

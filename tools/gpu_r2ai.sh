#!/bin/bash
# round 2, trip ai (2 GPUs): smoke(), the 2-GPU test, reference arm, 2-GPU bench + train-mode line after the model changes
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/r2_bench_n2_ai.json 2> gpurun_out/bench_n2_ai.err; cut -c1-250 gpurun_out/r2_bench_n2_ai.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --mode train > gpurun_out/r2_train_n2_ai.json 2> gpurun_out/train_n2_ai.err; cut -c1-250 gpurun_out/r2_train_n2_ai.json

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 tools/diag_dp_clocks.py 2>&1 | grep "^rank"
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/sac_rate.py --updates 300 2>&1 | grep "^b128 \|^b128_nccl" | cut -c1-100

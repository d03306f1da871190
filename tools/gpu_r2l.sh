#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -rs -s -x > gpurun_out/pytest_r2l.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_r2l.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/sac_rate.py --updates 300 > gpurun_out/sac_rate_n2.log 2>&1; echo "rate rc=$?"; grep "^b128\|^b4096\|^allreduce" gpurun_out/sac_rate_n2.log | cut -c1-110

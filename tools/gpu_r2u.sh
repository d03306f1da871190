#!/bin/bash
mkdir -p gpurun_out
RSB_PDL=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/sac_timeline.py 128 > gpurun_out/sac_timeline_b128_world2_nopdl.txt 2>&1; grep -v "Warn\|warn" gpurun_out/sac_timeline_b128_world2_nopdl.txt | tail -32

#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0; do RSB_PDL=$pdl timeout 300 python tools/sac_rate.py 2>&1 | grep "^b128 \|^b4096 " | cut -c1-120 | sed "s/^/PDL=$pdl /"; done | tee gpurun_out/sac_rate_r2e.log

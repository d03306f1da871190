#!/bin/bash
mkdir -p gpurun_out
RSB_PDL=0 timeout 300 python tools/sac_timeline.py 4096 > gpurun_out/sac_timeline_b4096_nopdl.txt 2>&1; grep -v Warn gpurun_out/sac_timeline_b4096_nopdl.txt | tail -34

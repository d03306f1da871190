#!/bin/bash
# First GPU trip after round 2: everything that was added after that round's GPU minutes were spent.
#   gpurun --timeout 900 -- 'bash tools/gpu_r3_first.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_zz_pickplace.py -q -p no:cacheprovider > gpurun_out/r3_newfam_tests.log 2>&1
tail -5 gpurun_out/r3_newfam_tests.log
for c in pickplacecan peginhole nutassemblyround handoff; do
  python bench.py --config $c --steps 20 --warmup 3 --no-sac --no-train --no-cpu > gpurun_out/r3_bench_$c.json 2> gpurun_out/r3_bench_$c.err
done
python tools/eval_committed_runs.py run 256 PickPlace > gpurun_out/r3_policy_transfer_pickplace.txt 2>&1
python tools/eval_committed_runs.py run 256 TwoArmPegInHole > gpurun_out/r3_policy_transfer_peginhole.txt 2>&1
python tools/eval_committed_runs.py run 256 NutAssemblyRound > gpurun_out/r3_policy_transfer_nutassembly.txt 2>&1
python tools/eval_committed_runs.py run 256 TwoArmHandoff > gpurun_out/r3_policy_transfer_handoff.txt 2>&1
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r3_all_gpu_tests.log 2>&1; tail -3 gpurun_out/r3_all_gpu_tests.log

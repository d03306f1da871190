#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_collector.py::test_no_contact_or_row_truncation_over_a_full_episode > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_r2b.log
timeout 900 python tools/limits_stats.py > gpurun_out/limits_r2b.log 2>&1; echo "limits rc=$?"; cat gpurun_out/limits_r2b.log | cut -c1-600

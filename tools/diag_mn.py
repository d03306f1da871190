"""Which descriptor-offset assignment fits the MN-major swizzled layout?  (developer tool, run under gpurun)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import diag_tc_gemm as d
from robosuite_benchmark_b200 import gemm
dev = torch.device("cuda:0")
for swap in (0, 1):
    d.lib().rsb_gemm_debug_mn_swap(swap)
    for c in d.cases():
        for nt, sp in ((32, 1), (64, 1), (128, 1), (0, 0), (64, 2)):
            err, *_ = d.run_case(c, dev, nt, sp)
            print(f"mn_swap={swap} {c[0]:26s} n_tile={nt:3d} splits={sp} mode A={d.lib().rsb_gemm_debug_last_tma() & 3} B={d.lib().rsb_gemm_debug_last_tma() >> 2} err {err:.2e} timeouts {gemm.timeouts()}")

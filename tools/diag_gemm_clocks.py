"""Phase clocks of CTA (0,0,0) of k_gemm_tf32 for a few shapes (developer tool, run under gpurun)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from robosuite_benchmark_b200 import gemm
from robosuite_benchmark_b200.backend import lib
dev = torch.device("cuda:0")
L = lib()
names = {0: "entry", 1: "tmem+bar ready, dep wait", 2: "first copies issued", 3: "first chunk landed", 4: "products issued", 5: "accumulator complete", 8: "acc parked in smem", 9: "barrier passed", 10: "C written", 6: "released", 7: "exit"}
for (M, N, K, nb, tag) in ((8192, 256, 256, 2, "Q L2 B=4096"), (4096, 256, 256, 1, "dH1p B=4096"), (256, 256, 256, 2, "Q L2 B=128")):
    a = torch.randn(nb, M, K, device=dev); w = torch.randn(nb, K, N, device=dev); b = torch.randn(nb, N, device=dev); o = torch.empty(nb, M, N, device=dev)
    for tma in (1, 0):
        L.rsb_gemm_debug_tma(tma)
        for _ in range(3):
            gemm.gemm_tf32(a, w, o, bias=b, relu=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                gemm.gemm_tf32(a, w, o, bias=b, relu=True)
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        clk = (C.c_longlong * 12)()
        L.rsb_gemm_debug_clocks(clk)
        base = clk[0]
        seq = [0, 1, 2, 3, 4, 5, 8, 9, 10, 6, 7]
        plan = (C.c_int * 8)(); L.rsb_gemm_plan(M, N, K, nb, 0, 0, plan)
        print(f"{tag} tma={tma} mode={L.rsb_gemm_debug_last_tma()} n_tile={plan[0]} splits={plan[1]} stages={plan[3]} ctas={plan[7]}: {e0.elapsed_time(e1) * 1000 / 20:.1f} us per launch in a graph; CTA0 cycles: " +
              ", ".join(f"{names[i]} +{clk[i] - base}" for i in seq))
L.rsb_gemm_debug_tma(-1)

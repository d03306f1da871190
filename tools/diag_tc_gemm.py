"""Diagnostic for csrc/rsb_tc_gemm.cu on a B200: error of the tcgen05 TF32 GEMM against fp64 for the SAC update's shapes and operand layouts,
for both orders of the descriptor byte offsets (`--swap 0|1`; run each in its own process), plus per-call time next to cuBLAS TF32.
    python tools/diag_tc_gemm.py --swap 0 [--time]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from robosuite_benchmark_b200 import gemm  # noqa: E402
from robosuite_benchmark_b200.backend import lib  # noqa: E402


def cases():
    # (name, nb, M, N, K, a_transposed, b_transposed, epilogue)
    return [("fwd 128x256x256", 1, 128, 256, 256, False, False, "bias_relu"), ("fwd twin 256x256x49", 2, 256, 256, 49, False, False, "bias_relu"),
            ("fwd 256x14x256", 1, 256, 14, 256, False, False, "bias"), ("fwd twin 128x1x256", 2, 128, 1, 256, False, False, "bias"),
            ("dX twin 256x256x256", 2, 256, 256, 256, False, True, "mask"), ("dX 128x49x256 acc", 1, 128, 49, 256, False, True, "acc"),
            ("dW twin 256x256x128", 2, 256, 256, 128, True, False, "none"), ("dW 42x256x128", 1, 42, 256, 128, True, False, "none"),
            ("outer 256x256x1", 2, 256, 256, 1, False, True, "mask"), ("big 4096x256x256", 1, 4096, 256, 256, False, False, "bias_relu"),
            ("dW big 256x256x4096", 2, 256, 256, 4096, True, False, "none"), ("odd 130x70x100", 3, 130, 70, 100, False, False, "bias_relu")]


def run_case(c, dev, n_tile=0, splits=0):
    lib().rsb_gemm_debug_splits(splits)
    name, nb, M, N, K, ta, tb, epi = c
    g = torch.Generator(device="cpu").manual_seed(hash(name) % 1000)
    A = torch.randn((nb, K, M) if ta else (nb, M, K), generator=g).to(dev)
    B = torch.randn((nb, N, K) if tb else (nb, K, N), generator=g).to(dev)
    a, b = (A.transpose(1, 2) if ta else A), (B.transpose(1, 2) if tb else B)
    bias = torch.randn(nb, N, generator=g).to(dev) if epi.startswith("bias") else None
    mask = torch.randn(nb, M, N, generator=g).to(dev) if epi == "mask" else None
    out = torch.randn(nb, M, N, generator=g).to(dev)
    ref = a.double() @ b.double()
    if bias is not None:
        ref = ref + bias.double()[:, None, :]
    if epi == "bias_relu":
        ref = ref.clamp_min(0)
    if mask is not None:
        ref = ref * (mask > 0)
    if epi == "acc":
        ref = ref + out.double()
    gemm.gemm_tf32(a, b, out, bias=bias, relu=(epi == "bias_relu"), mask=mask, accumulate=(epi == "acc"), n_tile=n_tile)
    bound = a.abs().double() @ b.abs().double()
    err = ((out.double() - ref).abs() / (bound + 1e-3)).max().item()
    return err, a, b, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--swap", type=int, default=0)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--chain", action="store_true", help="per-launch time of 40 dependent launches inside one CUDA graph")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib().rsb_gemm_debug_swap_offsets(args.swap)
    if args.chain:
        return chain(dev)
    worst = 0.0
    for c in cases():
        for nt, sp in ((0, 0), (16, 1), (16, 4), (32, 2), (128, 4)):
            err, a, b, out = run_case(c, dev, nt, sp)
            to = gemm.timeouts()
            worst = max(worst, err)
            print(f"swap={args.swap} n_tile={nt:3d} splits={sp} {c[0]:28s} rel err {err:.2e} timeouts {to}", flush=True)
    print(f"swap={args.swap} WORST {worst:.3e} -> {'OK' if worst < 3e-3 else 'WRONG'}")
    if args.time and worst < 3e-3:
        torch.backends.cuda.matmul.allow_tf32 = True
        for c in cases():
            err, a, b, out = run_case(c, dev)
            ref = torch.empty_like(out)
            def t(fn, n=200):
                for _ in range(20):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fn()
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n * 1e3
            us_tc = t(lambda: gemm.gemm_tf32(a, b, out))
            us_cb = t(lambda: torch.bmm(a, b, out=ref))
            fl = 2.0 * a.shape[0] * a.shape[1] * a.shape[2] * b.shape[2]
            print(f"time {c[0]:28s} tcgen05 {us_tc:8.2f} us ({fl / us_tc * 1e-6:7.2f} TFLOP/s)   cuBLAS tf32 {us_cb:8.2f} us (back-to-back launches, no graph)")


def chain(dev):
    """In-graph cost of one product of the update's dependency chain: 40 launches into one CUDA graph, each reading the previous output."""
    import ctypes as C
    L = lib()
    torch.backends.cuda.matmul.allow_tf32 = True
    st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for M, K in ((128, 256), (256, 49)):
        a, w, o = torch.randn(M, K, device=dev), torch.randn(K, 256, device=dev), torch.zeros(M, 256, device=dev)
        for _ in range(3):
            gemm.gemm_tf32(a, w, o)
        clk = (C.c_longlong * 12)()
        L.rsb_gemm_debug_clocks(clk)
        names = ("copies issued", "tmem ready", "chunk 0 landed", "products issued", "accumulator complete", "parked", "barrier", "C written", "released", "exit")
        print(f"clocks {M}x256x{K}: " + ", ".join(f"{n} {clk[i] - clk[0]}" for n, i in zip(names, (1, 2, 3, 4, 5, 8, 9, 10, 6, 7))), flush=True)
    for name, M, K in (("128x256x256", 128, 256), ("256x256x256", 256, 256), ("256x256x49", 256, 49), ("4096x256x256", 4096, 256)):
        x = [torch.randn(M, K, device=dev) * 0.1, torch.zeros(M, 256, device=dev), torch.zeros(M, 256, device=dev)]
        w0, w1, b = torch.randn(K, 256, device=dev) * 0.05, torch.randn(256, 256, device=dev) * 0.05, torch.zeros(256, device=dev)
        def tc(i):
            src, dst = (x[0], x[1]) if i == 0 else (x[1 + (i + 1) % 2], x[1 + i % 2])
            gemm.gemm_tf32(src, w0 if i == 0 else w1, dst, bias=b, relu=True)
        def tc_plus_small(i):
            tc(i)
            L.rsb_relu_bwd(C.c_void_p(x[1 + i % 2].data_ptr()), C.c_void_p(x[1 + i % 2].data_ptr()), 256, st())
        def cb(i):
            src, dst = (x[0], x[1]) if i == 0 else (x[1 + (i + 1) % 2], x[1 + i % 2])
            torch.mm(src, w0 if i == 0 else w1, out=dst)
        def cb_plus_bias(i):
            cb(i)
            L.rsb_bias_relu(C.c_void_p(x[1 + i % 2].data_ptr()), C.c_void_p(b.data_ptr()), M, 256, 1, 1, M * 256, 256, st())
        for label, fn in (("tcgen05 (bias+relu fused)", tc), ("tcgen05 + tiny kernel", tc_plus_small), ("cuBLAS mm", cb), ("cuBLAS mm + bias_relu", cb_plus_bias)):
            n = 40
            s = torch.cuda.Stream(dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                for i in range(3):
                    fn(i)
            torch.cuda.current_stream(dev).wait_stream(s); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(n):
                    fn(i)
            for _ in range(5):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            print(f"chain {name:14s} {label:28s} {e0.elapsed_time(e1) / 20 / n * 1e3:7.2f} us per step", flush=True)


if __name__ == "__main__":
    main()

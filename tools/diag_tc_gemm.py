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


def run_case(c, dev, n_tile=0):
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
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib().rsb_gemm_debug_swap_offsets(args.swap)
    worst = 0.0
    for c in cases():
        for nt in (0, 16, 128):
            err, a, b, out = run_case(c, dev, nt)
            to = gemm.timeouts()
            worst = max(worst, err)
            print(f"swap={args.swap} n_tile={nt:3d} {c[0]:28s} rel err {err:.2e} timeouts {to}", flush=True)
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


if __name__ == "__main__":
    main()

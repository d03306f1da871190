"""GPU parity of the tcgen05 TF32 GEMM (csrc/rsb_tc_gemm.cu, include/rsb_gemm.h) against fp64 on the SAC update's shapes and operand layouts.

Tolerance: TF32 keeps 10 mantissa bits of each operand, so |C - C64| <= 2 * 2^-10 * (|A| @ |B|) elementwise (written below as 3e-3, with
1e-3 absolute slack in the denominator); a wrong shared-memory layout or descriptor gives errors of order 1."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_tcgen05_gemm_matches_fp64_on_every_sac_shape(dev):
    import diag_tc_gemm as d
    from robosuite_benchmark_b200 import gemm
    for c in d.cases():
        for nt, splits in ((0, 0), (16, 1), (16, 4), (64, 2), (128, 4), (32, 1)):     # tile width x CTAs sharing a tile along K (0 = chosen)
            err, *_ = d.run_case(c, dev, nt, splits)
            assert gemm.timeouts() == 0, c
            assert err < 3e-3, (c, nt, splits, err)
    d.lib().rsb_gemm_debug_splits(0)


def test_tma_staging_is_bit_identical_to_the_cp_async_gather(dev):
    """Operands with one unit stride and 16-byte multiples otherwise are staged by TMA with the 128-byte swizzle -- K-major (contraction index contiguous)
    or MN-major (transposed activations, weights stored [in, out]) --, the rest by the strided cp.async gather into the no-swizzle K-major layout: three
    shared-memory layouts / descriptor kinds feeding the SAME products in the same order -- the outputs must be bit-identical, for every shape / tile
    width / split factor, and TMA must actually engage on the eligible operands (A in bits 0-1, B in bits 2-3; 1 = K-major, 2 = MN-major)."""
    import torch
    import diag_tc_gemm as d
    from robosuite_benchmark_b200 import gemm
    L = d.lib()
    engaged = {}
    for c in d.cases():
        for nt, splits in ((0, 0), (16, 4), (64, 2), (128, 1), (32, 1)):
            L.rsb_gemm_debug_tma(1)
            err, a, b, out_tma = d.run_case(c, dev, nt, splits)
            engaged[c[0]] = L.rsb_gemm_debug_last_tma()
            L.rsb_gemm_debug_tma(0)
            _, _, _, out_gather = d.run_case(c, dev, nt, splits)
            assert L.rsb_gemm_debug_last_tma() == 0
            assert gemm.timeouts() == 0, c
            assert err < 3e-3, (c, nt, splits, err)
            assert torch.equal(out_tma, out_gather), (c, nt, splits, (out_tma - out_gather).abs().max().item())
    L.rsb_gemm_debug_tma(-1); L.rsb_gemm_debug_splits(0)
    assert engaged["fwd 128x256x256"] == 1 | (2 << 2) and engaged["big 4096x256x256"] == 1 | (2 << 2)        # X K-major, W [in, out] MN-major
    assert engaged["dX twin 256x256x256"] == 1 | (1 << 2)                                                     # dY and W^T: both K-major
    assert engaged["dW twin 256x256x128"] == 2 | (2 << 2) and engaged["dW big 256x256x4096"] == 2 | (2 << 2)  # X^T and dY: both MN-major
    assert engaged["odd 130x70x100"] == 1 and engaged["fwd twin 256x256x49"] == 2 << 2                        # 280-byte rows of B / 196-byte rows of A: gather


def test_tcgen05_gemm_is_deterministic_and_graph_capturable(dev):
    import torch
    from robosuite_benchmark_b200 import gemm
    a, b = torch.randn(2, 256, 256, device=dev), torch.randn(2, 256, 256, device=dev)
    o1, o2 = torch.empty(2, 256, 256, device=dev), torch.empty(2, 256, 256, device=dev)
    gemm.gemm_tf32(a, b, o1)
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        gemm.gemm_tf32(a, b, o2)
    torch.cuda.current_stream(dev).wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        gemm.gemm_tf32(a, b, o2)
    o2.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    assert gemm.timeouts() == 0


def test_contraction_over_strided_blocks_equals_the_sum_of_products(dev):
    """stack_k: dX = dH_1 W_1^T + dH_2 W_2^T (the twin Q networks' input gradient) as one product over K = 2 x 256."""
    import torch
    from robosuite_benchmark_b200 import gemm
    dH = torch.randn(2, 256, 256, device=dev)                     # [net, 2B, H]; the product reads rows [0, B)
    W = torch.randn(2, 49, 256, device=dev)                       # [net, in, H]
    out = torch.empty(128, 49, device=dev)
    gemm.gemm_tf32(dH[:, :128], W.transpose(1, 2), out, stack_k=True)
    ref = sum(dH[i, :128].double() @ W[i].double().t() for i in range(2))
    bound = sum(dH[i, :128].abs().double() @ W[i].abs().double().t() for i in range(2))
    assert gemm.timeouts() == 0
    assert ((out.double() - ref).abs() / (bound + 1e-3)).max().item() < 3e-3

"""Torch-tensor front end of the tcgen05 TF32 GEMM (include/rsb_gemm.h, csrc/rsb_tc_gemm.cu).

`gemm_tf32(a, b, out, ...)` computes out[b] = epilogue(a[b] @ b[b]) for 2-D or 3-D fp32 CUDA tensors with ARBITRARY strides on `a` and `b`
(transposed / expanded / sliced views are read in place); `out` needs a contiguous last dimension.  Stands where `torch.mm/bmm`
(cuBLAS) stood in the SAC update (reference: rlkit's Linear layers under SACTrainer.train_from_torch, util/rlkit_custom.py:238).
There is no fallback: without the CUDA library the call raises.
"""
from __future__ import annotations

import ctypes as C

from .backend import lib, _stream_ptr

RELU, ACCUMULATE = 1, 2


def _p(t):
    return C.c_void_p(t.data_ptr())


def gemm_tf32(a, b, out, bias=None, relu=False, mask=None, accumulate=False, n_tile=0, stack_k=False):
    """out = epi(a @ b).  a: [(nb,) M, K], b: [(nb,) K, N], out: [(nb,) M, N]; bias: [(nb,) N]; mask: like out (ReLU backward: zero where mask <= 0).
    stack_k=True: a [nblk, M, KB], b [nblk, KB, N], out [M, N] = sum over the blocks of a[i] @ b[i] as ONE product with K = nblk * KB (KB % 64 == 0)."""
    k_block = a_kbs = b_kbs = 0
    if stack_k:
        assert a.dim() == 3 and b.dim() == 3 and out.dim() == 2 and a.shape[0] == b.shape[0] and a.shape[2] == b.shape[1] and a.shape[2] % 64 == 0
        nb, M, K, N = 1, a.shape[1], a.shape[0] * a.shape[2], b.shape[2]
        k_block, a_kbs, b_kbs = int(a.shape[2]), a.stride(0), b.stride(0)
        sa, sb = (a.stride(1), a.stride(2), 0), (b.stride(1), b.stride(2), 0)
        o3 = out.unsqueeze(0)
        bias3 = None if bias is None else bias.unsqueeze(0)
        mask3 = None if mask is None else mask.unsqueeze(0)
    else:
        if a.dim() == 2:
            a3, b3, o3 = a.unsqueeze(0), b.unsqueeze(0), out.unsqueeze(0)
            bias3 = None if bias is None else bias.unsqueeze(0)
            mask3 = None if mask is None else mask.unsqueeze(0)
        else:
            a3, b3, o3, bias3, mask3 = a, b, out, bias, mask
        nb, M, K = a3.shape
        N = b3.shape[2]
        assert b3.shape[0] == nb and b3.shape[1] == K, (a3.shape, b3.shape)
        sa, sb = (a3.stride(1), a3.stride(2), a3.stride(0)), (b3.stride(1), b3.stride(2), b3.stride(0))
    assert tuple(o3.shape) == (nb, M, N), (tuple(o3.shape), (nb, M, N))
    assert o3.stride(2) == 1 or N == 1, "out needs a contiguous last dimension"
    for t in (a, b, out):
        assert t.is_cuda and t.dtype.is_floating_point and t.element_size() == 4
    if bias3 is not None:
        assert tuple(bias3.shape) == (nb, N) and (bias3.stride(1) == 1 or N == 1)
    if mask3 is not None:
        assert tuple(mask3.shape) == (nb, M, N) and (mask3.stride(2) == 1 or N == 1)
    rc = lib().rsb_gemm_tf32(_p(a), sa[0], sa[1], sa[2], _p(b), sb[0], sb[1], sb[2], _p(o3), o3.stride(1), o3.stride(0), M, N, K, nb,
                             None if bias3 is None else _p(bias3), 0 if bias3 is None else bias3.stride(0),
                             None if mask3 is None else _p(mask3), 0 if mask3 is None else mask3.stride(1), 0 if mask3 is None else mask3.stride(0),
                             (RELU if relu else 0) | (ACCUMULATE if accumulate else 0), int(n_tile), k_block, a_kbs, b_kbs, _stream_ptr(a.device))
    if rc != 0:
        raise RuntimeError(lib().rsb_sac_last_error().decode())
    return out


def timeouts():
    """Device-side watchdog counter (0 on a healthy run); synchronises."""
    return int(lib().rsb_gemm_timeouts())

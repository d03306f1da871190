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


def gemm_tf32(a, b, out, bias=None, relu=False, mask=None, accumulate=False, n_tile=0):
    """out = epi(a @ b).  a: [(nb,) M, K], b: [(nb,) K, N], out: [(nb,) M, N]; bias: [(nb,) N]; mask: like out (ReLU backward: zero where mask <= 0)."""
    if a.dim() == 2:
        a3, b3, o3 = a.unsqueeze(0), b.unsqueeze(0), out.unsqueeze(0)
        bias3 = None if bias is None else bias.unsqueeze(0)
        mask3 = None if mask is None else mask.unsqueeze(0)
    else:
        a3, b3, o3, bias3, mask3 = a, b, out, bias, mask
    nb, M, K = a3.shape
    N = b3.shape[2]
    assert b3.shape[0] == nb and b3.shape[1] == K and tuple(o3.shape) == (nb, M, N), (a3.shape, b3.shape, o3.shape)
    assert o3.stride(2) == 1 or N == 1, "out needs a contiguous last dimension"
    for t in (a3, b3, o3):
        assert t.is_cuda and t.dtype.is_floating_point and t.element_size() == 4
    if bias3 is not None:
        assert tuple(bias3.shape) == (nb, N) and (bias3.stride(1) == 1 or N == 1)
    if mask3 is not None:
        assert tuple(mask3.shape) == (nb, M, N) and (mask3.stride(2) == 1 or N == 1)
    rc = lib().rsb_gemm_tf32(_p(a3), a3.stride(1), a3.stride(2), a3.stride(0), _p(b3), b3.stride(1), b3.stride(2), b3.stride(0),
                             _p(o3), o3.stride(1), o3.stride(0), M, N, K, nb,
                             None if bias3 is None else _p(bias3), 0 if bias3 is None else bias3.stride(0),
                             None if mask3 is None else _p(mask3), 0 if mask3 is None else mask3.stride(1), 0 if mask3 is None else mask3.stride(0),
                             (RELU if relu else 0) | (ACCUMULATE if accumulate else 0), int(n_tile), _stream_ptr(a.device))
    if rc != 0:
        raise RuntimeError(lib().rsb_sac_last_error().decode())
    return out


def timeouts():
    """Device-side watchdog counter (0 on a healthy run); synchronises."""
    return int(lib().rsb_gemm_timeouts())

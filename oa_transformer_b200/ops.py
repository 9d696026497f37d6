"""Python-side operator wrappers: torch tensors in, liboat C-ABI calls out (on the current CUDA stream).

Each wrapper only validates dtype/contiguity, fills the POD argument struct and launches; no arithmetic is done in
PyTorch here.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, check, lib, ptr, stream_ptr

ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_RELU = 0, 1, 2, 3


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a 2-D tensor with unit inner stride"
    return t.stride(0)


def gemm(A, B, *, a_major=0, b_major=0, alpha=1.0, bias=None, scale_cols=0, scale=1.0, act=ACT_NONE, aux=None,
         residual=None, out_f32=None, out_bf16=None, out2_bf16=None, accumulate=False, split_k=0):
    """C[M,N] = alpha * A(M,K) . B(N,K)^T with the fused epilogue of oat_gemm_bf16 (include/oat.h).

    a_major=0: A is [M,K]; a_major=1: A is stored [K,M] (M contiguous). Same for B with N.
    """
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    if a_major == 0:
        M, K = A.shape
    else:
        K, M = A.shape
    if b_major == 0:
        N, Kb = B.shape
    else:
        Kb, N = B.shape
    assert K == Kb, "contraction mismatch: %d vs %d" % (K, Kb)
    a = GemmArgs()
    a.A, a.lda, a.a_major = ptr(A), _ld(A), a_major
    a.B, a.ldb, a.b_major = ptr(B), _ld(B), b_major
    a.M, a.N, a.K = M, N, K
    a.alpha = alpha
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    a.bias = ptr(bias)
    a.scale_cols, a.scale = scale_cols, scale
    a.act = act
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and aux.shape == (M, N)
        a.aux_bf16, a.ld_aux = ptr(aux), _ld(aux)
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N)
        a.residual, a.ldr = ptr(residual), _ld(residual)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape == (M, N)
        a.out_f32, a.ld_f32 = ptr(out_f32), _ld(out_f32)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.shape == (M, N)
        a.out_bf16, a.ld_bf16 = ptr(out_bf16), _ld(out_bf16)
    if out2_bf16 is not None:
        assert out2_bf16.dtype == torch.bfloat16 and out2_bf16.shape == (M, N)
        a.out2_bf16, a.ld2 = ptr(out2_bf16), _ld(out2_bf16)
    a.accumulate = 1 if accumulate else 0
    a.split_k = split_k
    check(lib().oat_gemm_bf16(ctypes.byref(a), stream_ptr()), "oat_gemm_bf16")

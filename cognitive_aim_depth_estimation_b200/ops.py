"""Thin Python wrappers over the C-ABI entry points (one wrapper per entry point, no arithmetic here).

Tensors are torch CUDA tensors used purely as device-memory handles; every wrapper validates dtype /
contiguity, forwards raw pointers and raises `CogAimError` on a non-zero status.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

EPI_BIAS_BF16 = 0
EPI_GELU_BF16 = 1
EPI_RESID_F32 = 2
EPI_PATCH_F32 = 3
EPI_ROWSTATS = 4
EPI_COLSUM = 5
EPI_F32 = 6


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")


def stats_partials(n: int) -> int:
    return 2 * ((n + 127) // 128)


def gemm(A, W, epilogue, out=None, *, M=None, N=None, K=None, lda=None, ldw=None, batch=1, a_batch_stride=0,
         w_batch_stride=0, ldo=None, out_batch_stride=0, bias=None, ls=None, pos=None, patches_per_img=0,
         scale_log2=0.0, part_a=None, part_b=None, col_max=None, col_rinv=None):
    """C = epilogue(A @ W^T) on tcgen05 tensor cores (csrc/gemm.cu). A:[M,K] bf16, W:[N,K] bf16."""
    lib = _lib.load()
    _req(A, torch.bfloat16, "A")
    _req(W, torch.bfloat16, "W")
    M = A.shape[-2] if M is None else M
    K = A.shape[-1] if K is None else K
    N = W.shape[-2] if N is None else N
    lda = A.stride(-2) if lda is None else lda
    ldw = W.stride(-2) if ldw is None else ldw
    if out is not None and ldo is None:
        ldo = out.stride(-2)
    for t, n in ((bias, "bias"), (ls, "ls"), (pos, "pos"), (part_a, "part_a"), (part_b, "part_b"),
                 (col_max, "col_max"), (col_rinv, "col_rinv")):
        _req(t, torch.float32, n)
    st = lib.ca_gemm_bf16(ptr(A), ptr(W), M, N, K, lda, ldw, batch, a_batch_stride, w_batch_stride, epilogue,
                          ptr(out), ldo or 0, out_batch_stride, ptr(bias), ptr(ls), ptr(pos), patches_per_img,
                          float(scale_log2), ptr(part_a), ptr(part_b), ptr(col_max), ptr(col_rinv), stream_ptr())
    check(st, "ca_gemm_bf16")
    return out


LOG2E = math.log2(math.e)

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

PATCH_ROW_STRIDE = 592
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def attention(qkv, out, B, T, H):
    """out[B*T, H*64] = softmax(QK^T/8)V per (image, head) (csrc/attention.cu)."""
    _req(qkv, torch.bfloat16, "qkv")
    _req(out, torch.bfloat16, "out")
    check(_lib.load().ca_attention_bf16(ptr(qkv), ptr(out), B, T, H, stream_ptr()), "ca_attention_bf16")
    return out


def patchify_f32(images, patches):
    _req(images, torch.float32, "images")
    _req(patches, torch.bfloat16, "patches")
    B, _, S, _ = images.shape
    check(_lib.load().ca_patchify_f32(ptr(images), ptr(patches), B, S, stream_ptr()), "ca_patchify_f32")
    return patches


def preprocess_u8(images_hwc, patches, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    import ctypes as C
    _req(images_hwc, torch.uint8, "images")
    _req(patches, torch.bfloat16, "patches")
    B, S = images_hwc.shape[0], images_hwc.shape[1]
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    check(_lib.load().ca_preprocess_u8(ptr(images_hwc), ptr(patches), B, S, m, s, stream_ptr()), "ca_preprocess_u8")
    return patches


def cls_rows(x, cls, pos, B, T, D):
    check(_lib.load().ca_cls_rows(ptr(x), ptr(cls), ptr(pos), B, T, D, stream_ptr()), "ca_cls_rows")


def layernorm(x, gamma, beta, out, eps=1e-6):
    _req(x, torch.float32, "x")
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    check(_lib.load().ca_layernorm(ptr(x), ptr(gamma), ptr(beta), ptr(out), int(out.dtype == torch.bfloat16), rows, D,
                                   eps, stream_ptr()), "ca_layernorm")
    return out


def focal_input(tokens, pe, rowscale, xin, B, N, D):
    check(_lib.load().ca_focal_input(ptr(tokens), ptr(pe), ptr(rowscale), ptr(xin), B, N, D, stream_ptr()),
          "ca_focal_input")
    return xin

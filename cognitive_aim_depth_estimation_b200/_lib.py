"""ctypes binding of libcogaim_b200.so (C-ABI declared in include/cogaim_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing or a call fails, a
RuntimeError carrying `ca_last_error()` is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libcogaim_b200.so"

_lib = None

c_f32p = C.c_void_p  # device pointers travel as integers
c_ptr = C.c_void_p


class CogAimError(RuntimeError):
    pass


# name -> (argtypes); every function returns int status except where noted
_SIGNATURES = {
    "ca_version": [],
    "ca_device_check": [C.c_int],
    "ca_gemm_bf16": [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong,
                     C.c_longlong, C.c_int, c_ptr, C.c_int, C.c_longlong, c_ptr, c_ptr, c_ptr, C.c_int, C.c_float,
                     c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "ca_attention_bf16": [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr],
    "ca_patchify_f32": [c_ptr, c_ptr, C.c_int, C.c_int, c_ptr],
    "ca_preprocess_u8": [c_ptr, c_ptr, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), c_ptr],
    "ca_cls_rows": [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr],
    "ca_layernorm": [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_float, c_ptr],
    "ca_focal_input": [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr],
}


def exported_symbols():
    """Every symbol include/cogaim_b200.h declares (used by the CPU-side ABI test)."""
    return ["ca_last_error", *_SIGNATURES.keys()]


def load():
    """Load the library once; raise loudly if it has not been built (`__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CogAimError(
            f"{LIB_PATH} is missing: build it with `python -m cognitive_aim_depth_estimation_b200.build` "
            "(there is no CPU / PyTorch fallback for the Cognitive-Aim forward path)")
    lib = C.CDLL(str(LIB_PATH))
    lib.ca_last_error.restype = C.c_char_p
    lib.ca_last_error.argtypes = []
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().ca_last_error()
        raise CogAimError(f"{what} failed with status {status}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)

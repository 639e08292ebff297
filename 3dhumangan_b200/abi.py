"""ctypes binding of the C ABI declared in include/hg3d.h.

The product path has NO fallback: if `lib3dhg_sm100a.so` is missing, fails to load, or an entry
point returns non-zero, a RuntimeError is raised (mirroring TORCH_CHECK -> RuntimeError in the
reference's own native ops, lib/components/ops/bias_act.cpp:34-51).  All pointers are raw device
pointers taken from torch tensors; the library never allocates device memory and never
synchronises the device; the CUDA stream is passed explicitly (torch's current stream).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib3dhg_sm100a.so")
_lib = None

# name -> (restype, argtypes).  Keep in sync with include/hg3d.h (tests check every symbol).
SIGNATURES = {
    "hg_last_error": (c_char_p, []),
    "hg_abi_version": (c_int, []),
    "hg_check_device": (c_int, []),
    "hg_packed_weight_bytes": (c_size_t, [c_int, c_int, c_int]),
    "hg_pack_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "hg_linear": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
}


def lib():
    """Load the shared library once; raise loudly when it is absent (no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python 3dhumangan_b200/build.py` "
                "(this package has no CPU or eager-PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().hg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (or NULL).  Tensors must be CUDA + contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("hg3d: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("hg3d: expected a contiguous tensor")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_device(t=None):
    if not torch.cuda.is_available():
        raise RuntimeError("hg3d: no CUDA device visible; the sm_100a kernels cannot run (no fallback)")
    check(lib().hg_check_device(), "hg_check_device")


# ----------------------------------------------------------------------------------------------
# thin typed wrappers (shape checks live in C; these only marshal)
# ----------------------------------------------------------------------------------------------
def packed_weight_bytes(N, K, Nb):
    return int(lib().hg_packed_weight_bytes(N, K, Nb))


def pack_weight(W, Nb=None, scale=1.0, scale_dev=None, out=None):
    """W [N,K] fp32 (row stride may exceed K) -> packed bf16 hi/lo operand image (uint8 tensor)."""
    assert W.dim() == 2 and W.dtype == torch.float32 and W.stride(1) == 1
    N, K = W.shape
    if Nb is None:
        Nb = min(256, (N + 15) // 16 * 16)
    nbytes = packed_weight_bytes(N, K, Nb)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    with torch.cuda.device_of(W):
        check(lib().hg_pack_weight(c_void_p(W.data_ptr()), N, K, W.stride(0), ptr(scale_dev), float(scale), Nb,
                                   ptr(out), out.numel(), stream()), "hg_pack_weight")
    return out, Nb


def linear(X, Wimg, Nb, N, bias=None, passes=3, out=None):
    """Y = X @ W^T + bias with the packed weight image; X [M,K] fp32 row-major."""
    assert X.dim() == 2 and X.dtype == torch.float32 and X.stride(1) == 1
    M, K = X.shape
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=X.device)
    with torch.cuda.device_of(X):
        check(lib().hg_linear(c_void_p(X.data_ptr()), X.stride(0), M, K, ptr(Wimg), Nb, N, ptr(bias),
                              ptr(out), out.stride(0), passes, stream()), "hg_linear")
    return out

"""ctypes loader for libmspde.so (the C-ABI declared in include/mspde.h).

There is no CPU fallback: if the library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmspde.so")
_lib = None


class MspdeError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MspdeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the pCN hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mspde_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc: int, what: str = "mspde call"):
    """0 ok; >0 LAPACK-style info -> numpy.linalg.LinAlgError (what Simulation.run catches,
    /root/reference/mcmc/image_cupy.py:969); <0 CUDA/NCCL failure -> MspdeError."""
    if rc == 0:
        return
    if rc > 0:
        raise np.linalg.LinAlgError(f"{what}: factorisation failed, info={rc}")
    raise MspdeError(f"{what}: {lib().mspde_last_error().decode()}")


def ptr(t) -> ctypes.c_void_p:
    """Raw device (or host) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, np.ndarray):
        return ctypes.c_void_p(t.ctypes.data)
    return ctypes.c_void_p(t.data_ptr())


def current_stream() -> ctypes.c_void_p:
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


GEMM_4M, GEMM_3M = 0, 1


def set_gemm_mode(mode: int):
    """Process-wide complex product scheme of the GEMM kernel (include/mspde.h: MSPDE_GEMM_4M / MSPDE_GEMM_3M)."""
    check(lib().mspde_set_gemm_mode(int(mode)), "set_gemm_mode")


def get_gemm_mode() -> int:
    return int(lib().mspde_get_gemm_mode())


def set_gemm_pre_min(v: int):
    """Threshold on min(m, n) from which 3M products use pre-summed operand planes (include/mspde.h)."""
    check(lib().mspde_set_gemm_pre_min(int(v)), "set_gemm_pre_min")


def get_gemm_pre_min() -> int:
    return int(lib().mspde_get_gemm_pre_min())


def release_gemm_workspaces() -> int:
    """Waits for the device and frees every packed-operand workspace the GEMM kept (include/mspde.h); returns the bytes freed."""
    L = lib()
    L.mspde_release_gemm_workspaces.restype = ctypes.c_longlong
    return int(L.mspde_release_gemm_workspaces())


def set_qr_block(width: int):
    """64 or 128: block-reflector width of the Householder QR (include/mspde.h: mspde_set_qr_block)."""
    check(lib().mspde_set_qr_block(int(width)), "set_qr_block")


def get_qr_block() -> int:
    return int(lib().mspde_get_qr_block())

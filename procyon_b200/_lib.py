"""ctypes binding of libprocyon_b200.so (the C ABI declared in include/procyon_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libprocyon_b200.so"

_lib = None


class ProcyonB200Error(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing or os.environ.get("PROCYON_B200_NO_BUILD"):
            raise ProcyonB200Error(f"{LIB_PATH} is missing: run `python -m procyon_b200.build` (there is no CPU fallback)")
        from . import build as _build

        _build.build()
    _lib = ctypes.CDLL(str(LIB_PATH))
    _lib.pcy_last_error.restype = ctypes.c_char_p
    _lib.pcy_launch_count.restype = ctypes.c_longlong
    _lib.pcy_reset_launch_count.restype = None
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pcy_last_error().decode("utf-8", "replace")
        raise ProcyonB200Error(f"{what} failed with status {rc}: {msg}")


def ptr(t) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ProcyonB200Error("procyon_b200 kernels need CUDA tensors (no CPU fallback exists)")


c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_float = ctypes.c_float

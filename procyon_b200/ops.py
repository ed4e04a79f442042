"""Thin Python wrappers over the op-level C ABI (device pointers in, device pointers out).

Used by the host-side mirrors of the reference modules and by the parity tests; every function
enqueues on torch's current CUDA stream.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import c_float, c_i64, c_int, check, ptr, stream_ptr

ACT_NONE, ACT_GELU, ACT_SWIGLU = 0, 1, 2


def _bf16_2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{name} must be bfloat16, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D with unit inner stride")
    return t


def linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *, residual: torch.Tensor | None = None,
           act: int = ACT_NONE, scale: float = 1.0, scale_ncols: int = 0, out: torch.Tensor | None = None,
           out_fp32: bool = False, force: str | None = None, rms_weight: torch.Tensor | None = None,
           rms_eps: float = 1e-5) -> torch.Tensor:
    """out = epi(a @ w.T) — see pcy_linear_bf16 in include/procyon_b200.h.

    `bias` must be fp32. `force` in {None, "tc", "skinny"} pins the kernel (tests).
    """
    lib = _lib.load()
    _lib.require_cuda(a, w, bias, residual, out)
    a = _bf16_2d(a, "a")
    w = _bf16_2d(w, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    else:
        out_fp32 = out.dtype == torch.float32
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("bias must be float32 (pack it once at load time)")
    if residual is not None:
        residual = _bf16_2d(residual, "residual")
    args = [ptr(a), c_i64(a.stride(0)), ptr(w), c_i64(w.stride(0)), ptr(out), c_i64(out.stride(0)), c_int(M), c_int(N),
            c_int(K), ptr(bias), ptr(residual), c_i64(residual.stride(0) if residual is not None else 0), c_int(act),
            c_float(scale), c_int(scale_ncols), c_int(1 if out_fp32 else 0)]
    if force is None and rms_weight is None:
        check(lib.pcy_linear_bf16(*args, stream_ptr(a.device)), "pcy_linear_bf16")
    else:
        force_tc = 1 if force == "tc" else 0
        if force is None:
            force_tc = 0
        check(lib.pcy_linear_bf16_ex(*args, c_int(force_tc), ptr(rms_weight), c_float(rms_eps), stream_ptr(a.device)),
              "pcy_linear_bf16_ex")
    return out


def pack_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _lib.require_cuda(gate, up)
    F, K = gate.shape
    packed = torch.empty((2 * F, K), device=gate.device, dtype=torch.bfloat16)
    check(lib.pcy_pack_gate_up(ptr(gate.contiguous()), ptr(up.contiguous()), ptr(packed), c_int(F), c_int(K),
                               stream_ptr(gate.device)), "pcy_pack_gate_up")
    return packed


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """torch.nn.LayerNorm over the last dim of bf16 rows [n, d] (pcy_layernorm_bf16); weight / bias bf16."""
    lib = _lib.load()
    _lib.require_cuda(x)
    x = _bf16_2d(x, "x").contiguous()
    w = weight.to(device=x.device, dtype=torch.bfloat16).contiguous()
    b = bias.to(device=x.device, dtype=torch.bfloat16).contiguous()
    y = torch.empty_like(x)
    check(lib.pcy_layernorm_bf16(ptr(x), ptr(w), ptr(b), ptr(y), c_i64(x.shape[0]), c_int(x.shape[1]), c_float(eps),
                                 stream_ptr(x.device)), "pcy_layernorm_bf16")
    return y

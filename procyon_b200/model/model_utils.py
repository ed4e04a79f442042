"""Projector MLPs and small helpers — host-side mirror of procyon/model/model_utils.py.

create_mlp            reference: procyon/model/model_utils.py:13-41   (same nn.Sequential layout => same keys)
compute_conflict_matrix                                   :135-146
left_pad_tensors                                          :151-170
lm_loss               HF LlamaForCausalLM shifted cross-entropy (used through pmc_llama.py:576), fused + chunked
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from .._lib import c_i64, c_int, check, ptr, stream_ptr


class FusedMLP(nn.Sequential):
    """nn.Sequential([Linear, Dropout, GELU] x (n-1) + [Linear]) whose forward runs the fused kernels.

    Each Linear+bias(+GELU) is one GEMM launch (tcgen05 for M > 16 rows, weight-streaming below). Dropout is the
    identity: the hot path is inference (`model.eval()`); training-mode dropout is not reproduced.
    """

    def _packed(self, lin: nn.Linear, device):
        key = (lin.weight.data_ptr(), lin.weight._version, None if lin.bias is None else lin.bias._version, str(device))
        cache = getattr(lin, "_pcy_cache", None)
        if cache is None or cache[0] != key:
            w = lin.weight.detach().to(device=device, dtype=torch.bfloat16).contiguous()
            b = None
            if lin.bias is not None:
                # bias lives in the module dtype in the reference (bf16 after .bfloat16()); keep that rounding
                b = lin.bias.detach().to(device=device, dtype=torch.bfloat16).float().contiguous()
            lin._pcy_cache = (key, w, b)
            cache = lin._pcy_cache
        return cache[1], cache[2]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(x)
        lead = x.shape[:-1]
        h = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                j = i + 1
                act = ops.ACT_NONE
                while j < len(mods) and not isinstance(mods[j], nn.Linear):
                    if isinstance(mods[j], nn.GELU):
                        act = ops.ACT_GELU
                    j += 1
                w, b = self._packed(m, h.device)
                if h.shape[0] == 0:
                    h = torch.empty((0, w.shape[0]), device=h.device, dtype=torch.bfloat16)
                else:
                    h = ops.linear(h, w, b, act=act)
                i = j
            else:
                i += 1
        return h.reshape(*lead, h.shape[-1])


def create_mlp(n_layers, in_features, out_features, hidden_features=256, dropout_rate=0.25):
    """Same module layout as the reference (state_dict keys 0,3,6,... for n_layers > 1; '0' without bias for 1)."""
    if n_layers == 1:
        return FusedMLP(nn.Linear(in_features, out_features, bias=False))
    layers = []
    for i in range(n_layers):
        in_size = hidden_features if i > 0 else in_features
        if i < n_layers - 1:
            layers.append(nn.Linear(in_size, hidden_features))
            if dropout_rate is not None:
                layers.append(nn.Dropout(dropout_rate))
            layers.append(nn.GELU())
        else:
            layers.append(nn.Linear(in_size, out_features))
    return FusedMLP(*layers)


def compute_conflict_matrix(id1, id2):
    """[i, j] is True when samples i and j share their first id but not their second: a false negative of in-batch
    contrastive learning (procyon/model/model_utils.py:135-146)."""
    same_first = id1[:, None] == id1[None, :]
    same_second = id2[:, None] == id2[None, :]
    return same_first & ~same_second


def left_pad_tensors(tensors: List[torch.Tensor], pad_value=0):
    """Left-pad 1-D id tensors to a common length; the mask is float32 0/1 as in the reference."""
    max_length = max(t.size(0) for t in tensors)
    padded, masks = [], []
    for t in tensors:
        pad = max_length - t.size(0)
        padded.append(torch.cat([torch.full((pad,), pad_value, dtype=t.dtype), t]))
        masks.append(torch.cat([torch.zeros(pad), torch.ones(t.size(0))]))
    return torch.stack(padded), torch.stack(masks)


def lm_loss(text_encoder, hidden: torch.Tensor, labels: torch.Tensor, chunk_rows: int = 512) -> torch.Tensor:
    """HF causal-LM loss: mean CE of logits[:, :-1] against labels[:, 1:] with ignore_index=-100, fp32.

    hidden = post-final-norm states bf16 [B,S,d]. The (B*S, V) logits are never materialised: rows with a live
    target are gathered, pushed through the LM head `chunk_rows` at a time (tcgen05 GEMM, fp32 logits) and
    reduced by the fused cross-entropy kernel.
    """
    lib = _lib.load()
    B, S, d = hidden.shape
    dev = hidden.device
    tgt = labels.to(dev)[:, 1:].reshape(-1)
    rows = hidden[:, :-1, :].reshape(-1, d)
    live = (tgt != -100).nonzero(as_tuple=True)[0]
    acc = torch.zeros(2, device=dev, dtype=torch.float32)  # [sum of row losses, number of rows]
    if live.numel() > 0:
        rows = rows.index_select(0, live).contiguous()
        tgt32 = tgt.index_select(0, live).to(torch.int32).contiguous()
        w = text_encoder._lm_head_bf16(dev)
        V = w.shape[0]
        for r0 in range(0, rows.shape[0], chunk_rows):
            r1 = min(rows.shape[0], r0 + chunk_rows)
            logits = ops.linear(rows[r0:r1], w, out_fp32=True)
            check(lib.pcy_cross_entropy_rows(ptr(logits), ptr(tgt32[r0:r1]), c_int(r1 - r0), c_int(V),
                                             c_i64(logits.stride(0)), ptr(acc), stream_ptr(dev)),
                  "pcy_cross_entropy_rows")
    return acc[0] / acc[1]

"""InfoNCE with in-batch negatives — host-side mirror of procyon/model/contrastive.py:95-204.

Forward only (the build has no backward): normalisation, similarity, mask and both cross-entropies run in
libprocyon_b200.so; the cross-rank exchange is one NCCL all-gather per side through torch.distributed.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from .. import _lib
from .._lib import c_float, c_int, check, ptr, stream_ptr


def _normalize(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    x = x.float().contiguous()
    out = torch.empty_like(x)
    check(lib.pcy_normalize_rows(ptr(x), ptr(out), c_int(x.shape[0]), c_int(x.shape[1]), stream_ptr(x.device)),
          "pcy_normalize_rows")
    return out


class InfoNCEInBatch(nn.Module):
    def __init__(self, input_embed_dim, use_projection=True, all_gather_version=False):
        super().__init__()
        if use_projection:
            raise NotImplementedError("use_projection_cl=True is not used by ProCyon-Full (llama3-full.yml:101)")
        self.input_embed_dim = input_embed_dim
        self.use_projection = use_projection
        self.all_gather_version = all_gather_version
        self.temperature = nn.Parameter(0.07 * torch.ones([]))
        self.ret_projection, self.protein_z_projection = None, None

    def forward(self, c_input, negatives_mask=None):
        lib = _lib.load()
        with torch.no_grad():
            self.temperature.clamp_(0.001, 0.5)
        zs = _normalize(c_input["positive"]["sequence"])
        zt = _normalize(c_input["positive"]["text"])
        _lib.require_cuda(zs, zt)
        b, d = zs.shape
        if zt.shape[0] != b:
            raise ValueError("InfoNCE needs as many text as sequence embeddings")
        gathered = self.all_gather_version and dist.is_available() and dist.is_initialized()
        if gathered:
            W, rank = dist.get_world_size(), dist.get_rank()
            all_s = torch.empty((W * b, d), device=zs.device, dtype=torch.float32)
            all_t = torch.empty((W * b, d), device=zs.device, dtype=torch.float32)
            dist.all_gather_into_tensor(all_s, zs)
            dist.all_gather_into_tensor(all_t, zt)
            G, off = W * b, rank * b
        else:
            # one rank: G = b, offset 0.  (With a mask the reference dereferences a local that only the gathered branch
            # defines, contrastive.py:186, and raises NameError; the loss is well defined — rows of the (b, b) mask
            # multiplied into the logits — so it is computed instead of failing.  Listed in DESIGN.md.)
            all_s, all_t, G, off = zs, zt, b, 0
        return infonce_loss(zs, zt, all_s, all_t, negatives_mask, off, float(self.temperature.detach()))


def infonce_loss(zs, zt, all_s, all_t, negatives_mask, rank_off: int, temperature: float) -> torch.Tensor:
    """One rank's loss on L2-normalised fp32 embeddings: zs / zt [b, d] local, all_s / all_t [G, d] the rank-ordered
    all-gather, negatives_mask [G, G] (0/1, multiplied into this rank's rows of the logits like the reference does,
    contrastive.py:195-196) or None, targets rank_off + i.  One launch sequence in libprocyon_b200.so
    (`pcy_infonce_loss`)."""
    lib = _lib.load()
    _lib.require_cuda(zs, zt, all_s, all_t)
    b, d = zs.shape
    G = all_s.shape[0]
    mask = None
    if negatives_mask is not None:
        mask = negatives_mask.to(device=zs.device, dtype=torch.uint8).contiguous()
        if mask.shape != (G, G):
            raise ValueError(f"negatives_mask must be {(G, G)}, got {tuple(mask.shape)}")
    scratch = torch.empty(2 * b * G, device=zs.device, dtype=torch.float32)
    loss = torch.empty(1, device=zs.device, dtype=torch.float32)
    check(lib.pcy_infonce_loss(ptr(zs), ptr(zt), ptr(all_s.contiguous()), ptr(all_t.contiguous()), ptr(mask),
                               ptr(scratch), ptr(loss), c_int(b), c_int(G), c_int(d), c_int(rank_off),
                               c_float(temperature), stream_ptr(zs.device)), "pcy_infonce_loss")
    return loss[0]

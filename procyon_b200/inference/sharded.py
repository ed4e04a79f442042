"""Multi-GPU sharding of the two paths that shard naturally: protein encode and retrieval scoring.

Protein encode mirrors the reference's eval-time pattern (procyon/training/trainIT.py:1595-1610 with
SequentialDistributedSampler, procyon/data/samplers.py:154-198): rank r encodes the contiguous block
[r*ceil(N/W), (r+1)*ceil(N/W)) of the proteins and the pooled embeddings are all-gathered (one NCCL all-gather of
(ceil(N/W), d) rows per rank), so the concatenation is already in protein order. The tail block is padded by
wrapping around to the first proteins, like the sampler, and trimmed after the gather.
No collective sits on the data path of the encode itself; Llama decode is never sharded (replicas only).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    per = (n + world - 1) // world
    return rank * per, min(n, (rank + 1) * per), per


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def encode_proteins_sharded(encode_fn: Callable[[torch.Tensor], torch.Tensor], tokens: torch.Tensor, group=None,
                            gather: bool = True) -> torch.Tensor:
    """tokens [N,T] (same on every rank) -> pooled embeddings [N,d] on every rank (or the local block if not gather).

    encode_fn maps a token block [n,T] to [n,d] (e.g. `lambda t: model.forward_sequences(t)["shared"]`).
    """
    W, rank = _world(group)
    N = tokens.shape[0]
    lo, hi, per = shard_bounds(N, W, rank)
    idx = torch.arange(lo, lo + per) % max(N, 1)  # wrap-around padding of the last block
    local = encode_fn(tokens[idx.to(tokens.device)]) if per > 0 else encode_fn(tokens[:0])
    if W == 1 or not gather:
        return local[: hi - lo] if not gather else local[:N]
    out = torch.empty((W * per, local.shape[1]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:N]


def sharded_scores(score_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], queries: torch.Tensor,
                   database: torch.Tensor, group=None) -> torch.Tensor:
    """Full (Q,N) score matrix on every rank from a row-sharded database; score_fn(q, db_block) -> (Q, n_block)."""
    W, rank = _world(group)
    N = database.shape[0]
    lo, hi, per = shard_bounds(N, W, rank)
    Q = queries.shape[0]
    local = torch.zeros((Q, per), device=queries.device, dtype=torch.float32)
    if hi > lo:
        local[:, : hi - lo] = score_fn(queries, database[lo:hi])
    if W == 1:
        return local[:, :N]
    gathered = torch.empty((W * Q, per), device=queries.device, dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)
    return gathered.view(W, Q, per).permute(1, 0, 2).reshape(Q, W * per)[:, :N]

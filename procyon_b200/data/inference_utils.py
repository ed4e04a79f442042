"""Retrieval scoring — host-side mirror of the scoring functions in procyon/data/inference_utils.py:921-999.

`get_proteins_from_embedding` / `get_proteins_from_batched_embeddings` keep the reference signatures; the cosine
scores come from one streaming kernel (`pcy_cosine_scores`: normalise + dot in a single pass over the database,
which stays resident on the GPU instead of being re-normalised and re-uploaded on every query, reference :958-961).
`ShardedProteinIndex` is the multi-GPU form: the database is row-sharded over the ranks of a process group, every
rank scores its shard and the per-shard scores are all-gathered (NCCL) into the full (Q, N) matrix.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import _lib
from .._lib import c_i64, c_int, check, ptr, stream_ptr


def cosine_scores(query_embeddings: torch.Tensor, protein_embeds: torch.Tensor,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(Q,d) x (N,d) -> (Q,N) fp32 cosine similarities on the device (F.normalize semantics, eps 1e-12)."""
    lib = _lib.load()
    dev = protein_embeds.device if protein_embeds.is_cuda else query_embeddings.device
    if dev.type != "cuda":
        raise _lib.ProcyonB200Error("cosine_scores needs the database or the query on a CUDA device")
    db = protein_embeds.to(dev)
    if db.dtype not in (torch.float32, torch.bfloat16):
        db = db.float()
    db = db.contiguous()
    q = query_embeddings.detach().to(device=dev, dtype=torch.float32).reshape(-1, db.shape[1]).contiguous()
    Q, N, d = q.shape[0], db.shape[0], db.shape[1]
    if out is None:
        out = torch.empty((Q, N), device=dev, dtype=torch.float32)
    check(lib.pcy_cosine_scores(ptr(q), ptr(db), c_int(1 if db.dtype == torch.bfloat16 else 0), ptr(out), c_int(Q),
                                c_int(N), c_int(d), c_i64(out.stride(0)), stream_ptr(dev)), "pcy_cosine_scores")
    return out


FUSED_TOPK_MAX = 32  # pcy_retrieval_scores_topk ranks inside the scoring launch up to this k
_tickets = {}


def _ticket(dev: torch.device) -> torch.Tensor:
    """Per-device workspace of the fused ranking: the 'last CTA' counter (zero on entry, re-zeroed by the kernel) and
    the per-CTA candidate lists."""
    key = str(dev)
    if key not in _tickets:
        import ctypes

        f = _lib.load().pcy_retrieval_workspace_bytes
        f.restype = ctypes.c_int64
        with torch.cuda.device(dev):
            _tickets[key] = torch.zeros(int(f()) // 4 + 1, device=dev, dtype=torch.int32)
    return _tickets[key]


def retrieval_scores_topk(query_embeddings: torch.Tensor, protein_embeds: torch.Tensor, k: int, index_base: int = 0,
                          scores_out: Optional[torch.Tensor] = None):
    """Cosine scores (Q, N) AND their k <= 32 best rows per query in one launch (`pcy_retrieval_scores_topk`).
    Returns (scores fp32 [Q, N], top_val fp32 [Q, k], top_idx int32 [Q, k] = row + index_base, -1 where N < k)."""
    lib = _lib.load()
    dev = protein_embeds.device if protein_embeds.is_cuda else query_embeddings.device
    if dev.type != "cuda":
        raise _lib.ProcyonB200Error("retrieval scoring needs the database or the query on a CUDA device")
    if not 1 <= k <= FUSED_TOPK_MAX:
        raise ValueError(f"fused top-k supports 1 <= k <= {FUSED_TOPK_MAX}, got {k}")
    db = protein_embeds.to(dev)
    if db.dtype not in (torch.float32, torch.bfloat16):
        db = db.float()
    db = db.contiguous()
    q = query_embeddings.detach().to(device=dev, dtype=torch.float32).reshape(-1, db.shape[1]).contiguous()
    Q, N, d = q.shape[0], db.shape[0], db.shape[1]
    scores = scores_out if scores_out is not None else torch.empty((Q, N), device=dev, dtype=torch.float32)
    top_val = torch.empty((Q, k), device=dev, dtype=torch.float32)
    top_idx = torch.empty((Q, k), device=dev, dtype=torch.int32)
    check(lib.pcy_retrieval_scores_topk(ptr(q), ptr(db), c_int(1 if db.dtype == torch.bfloat16 else 0), ptr(scores),
                                        c_int(Q), c_int(N), c_int(d), c_i64(scores.stride(0)), c_int(k),
                                        c_int(index_base), ptr(top_val), ptr(top_idx), ptr(_ticket(dev)),
                                        stream_ptr(dev)), "pcy_retrieval_scores_topk")
    return scores, top_val, top_idx


def get_proteins_from_embedding(protein_embeds: torch.Tensor, model_out: Optional[Dict] = None,
                                query_embeddings: Optional[torch.Tensor] = None, protein_ids=None, top_k: int = 20):
    """Returns a DataFrame (uniprot_id, name, sim_score) of the top_k hits (all proteins when top_k is None)."""
    import pandas as pd

    assert model_out is not None or query_embeddings is not None
    if model_out is not None:
        assert query_embeddings is None
        query_embeddings = model_out["contrastive_out"]["positive"]["text"][0, :].unsqueeze(0).detach()
    if top_k is not None and top_k <= FUSED_TOPK_MAX:
        # scores + ranking in one launch; only 2 * top_k numbers come back to the host
        _, top_val, top_idx = retrieval_scores_topk(query_embeddings, protein_embeds, top_k)
        keep = (top_idx[0] >= 0).cpu()
        top = top_idx[0].cpu()[keep].tolist()
        sim_sub = top_val[0].cpu()[keep].tolist()
    else:  # whole ranking (top_k=None) or more hits than the fused ranking holds: library sort of the device scores
        sims = cosine_scores(query_embeddings, protein_embeds)[0]
        sort_inds = torch.argsort(sims, descending=True)
        if top_k is not None:
            sort_inds = sort_inds[:top_k]
        top = sort_inds.cpu().tolist()
        sim_sub = sims[sort_inds].cpu().tolist()
    if protein_ids is None:
        return pd.DataFrame({"index": top, "sim_score": sim_sub})
    return pd.DataFrame({"uniprot_id": protein_ids["protein_id"].iloc[top], "name": protein_ids["name"].iloc[top],
                         "sim_score": sim_sub})


@torch.no_grad()
def get_proteins_from_batched_embeddings(protein_embeds: torch.Tensor,
                                         query_embeddings: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert query_embeddings is not None
    return cosine_scores(query_embeddings, protein_embeds).squeeze().detach().cpu().float()


class ShardedProteinIndex:
    """Row-sharded protein-embedding database for multi-GPU retrieval (BASELINE config 4).

    Rank r keeps rows [r*ceil(N/W), (r+1)*ceil(N/W)) resident on its GPU. `scores(q)` returns the full (Q, N) cosine
    matrix on every rank: local streaming kernel + one all-gather of (Q, N/W) fp32 per rank. Works without a process
    group (W = 1).
    """

    def __init__(self, protein_embeds: torch.Tensor, device, group=None):
        import torch.distributed as dist

        self.group = group
        self.W = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.W > 1 else 0
        self.N, self.d = protein_embeds.shape
        self.per = (self.N + self.W - 1) // self.W
        lo, hi = self.rank * self.per, min(self.N, (self.rank + 1) * self.per)
        self.local = protein_embeds[lo:hi].to(device=device).contiguous()
        if self.local.dtype not in (torch.float32, torch.bfloat16):
            self.local = self.local.float()
        self.device = torch.device(device)

    def scores(self, query_embeddings: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        q = query_embeddings.reshape(-1, self.d)
        Q = q.shape[0]
        local = torch.zeros((Q, self.per), device=self.device, dtype=torch.float32)
        if self.local.shape[0] > 0:
            cosine_scores(q, self.local, out=local[:, : self.local.shape[0]])
        if self.W == 1:
            return local[:, : self.N]
        gathered = torch.empty((self.W * Q, self.per), device=self.device, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, local.contiguous(), group=self.group)
        return gathered.view(self.W, Q, self.per).permute(1, 0, 2).reshape(Q, self.W * self.per)[:, : self.N]

    def topk(self, query_embeddings: torch.Tensor, k: int = 20):
        """(values [Q, k], global row indices [Q, k]) of the k best database rows per query.  Every rank ranks its own
        shard inside the scoring launch and only the k candidates per rank are exchanged (2 * W * k numbers per query
        instead of N scores), then merged by value (`pcy_topk_merge`)."""
        import torch.distributed as dist

        k = min(k, self.N)
        if k > FUSED_TOPK_MAX:
            return torch.topk(self.scores(query_embeddings), k, dim=-1)
        q = query_embeddings.reshape(-1, self.d)
        Q = q.shape[0]
        if self.local.shape[0] > 0:
            _, val, idx = retrieval_scores_topk(q, self.local, k, index_base=self.rank * self.per)
        else:
            val = torch.full((Q, k), float("-inf"), device=self.device)
            idx = torch.full((Q, k), -1, device=self.device, dtype=torch.int32)
        if self.W == 1:
            return val, idx.to(torch.int64)
        all_val = torch.empty((self.W, Q, k), device=self.device, dtype=torch.float32)
        all_idx = torch.empty((self.W, Q, k), device=self.device, dtype=torch.int32)
        dist.all_gather_into_tensor(all_val, val.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=self.group)
        cand_val = all_val.permute(1, 0, 2).reshape(Q, self.W * k).contiguous()
        cand_idx = all_idx.permute(1, 0, 2).reshape(Q, self.W * k).contiguous()
        out_val = torch.empty((Q, k), device=self.device, dtype=torch.float32)
        out_idx = torch.empty((Q, k), device=self.device, dtype=torch.int32)
        check(_lib.load().pcy_topk_merge(ptr(cand_val), ptr(cand_idx), c_int(Q), c_int(self.W * k), c_int(k),
                                         ptr(out_val), ptr(out_idx), stream_ptr(self.device)), "pcy_topk_merge")
        return out_val, out_idx.to(torch.int64)


# ---- QA inference (procyon/data/inference_utils.py:581-655) --------------------------------------------------------
from ..training.train_utils import get_qa_logits_inference  # noqa: E402,F401  (same home as in the reference)


class ProCyonQAInference:
    """Yes/no question answering on top of UnifiedProCyon.forward: `qa(model_inputs)["pred"]` holds the probabilities
    over the vocabulary at the answer position (reference: `ProCyonQAInference.fwd_pass`).  Only the B answer rows
    go through the LM head."""

    def __init__(self, model, device=None):
        model.eval()
        self.device = device
        self.model = model.to(device) if device is not None else model
        self.yes_token, self.no_token = model.yes_token, model.no_token

    @torch.no_grad()
    def __call__(self, *args, **kwargs):
        return self.fwd_pass(*args, **kwargs)

    def fwd_pass(self, model_inputs, aaseq_type: str = "protein", output_attentions=None):
        out = self.model(model_inputs, return_mlm=False, retrieval=False, get_full_labels=True, aaseq_type=aaseq_type,
                         crop_off=True, output_attentions=output_attentions)
        pred, _ = get_qa_logits_inference(out, answer_token=self.model.answer_idx)
        return {"pred": pred, "y": None, "out": out}


# ---- batching of single-query model inputs (procyon/data/inference_utils.py:848-884) ------------------------------
def merge_model_input_dicts(dict_list):
    """Concatenates single-query model-input dicts (as built by `create_input_retrieval`) into one batch: `data`
    entries are concatenated, every `input` index list is shifted past the indices already in the batch and appended
    as a new row, `instructions` are concatenated.  Like the reference, the first dict is extended in place."""
    if len(dict_list) == 1:
        return dict_list[0]
    mega = dict_list[0]
    for d in dict_list[1:]:
        for k, v in d["data"].items():
            if v is None:
                continue
            if isinstance(v, list):
                mega["data"][k] = mega["data"][k] + v
            else:
                mega["data"][k] = torch.cat([mega["data"][k], v])
        for k, v in d["input"].items():
            if v is None:
                continue
            shift = max(max(r) for r in mega["input"][k]) + 1  # (the reference's np.max needs equally long rows)
            mega["input"][k] += [[val + shift for val in v[0]]]
        mega["instructions"] += d["instructions"]
    return mega

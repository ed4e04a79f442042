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


def get_proteins_from_embedding(protein_embeds: torch.Tensor, model_out: Optional[Dict] = None,
                                query_embeddings: Optional[torch.Tensor] = None, protein_ids=None, top_k: int = 20):
    """Returns a DataFrame (uniprot_id, name, sim_score) of the top_k hits (all proteins when top_k is None)."""
    import pandas as pd

    assert model_out is not None or query_embeddings is not None
    if model_out is not None:
        assert query_embeddings is None
        query_embeddings = model_out["contrastive_out"]["positive"]["text"][0, :].unsqueeze(0).detach()
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
        s = self.scores(query_embeddings)
        return torch.topk(s, min(k, self.N), dim=-1)


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

"""Oracle: fusion glue of UnifiedProCyon (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates create_mlp's forward (procyon/model/model_utils.py:13-41), the soft-token splice
(procyon/model/model_unified.py:1135-1175), label masking (:521-538, mask_before :39-60), the retrieval head
(:556-579), InfoNCEInBatch (procyon/model/contrastive.py:120-204) and cosine scoring
(procyon/data/inference_utils.py:955-970).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def _rounder(act_round: str):
    if act_round == "bf16":
        return lambda t: t.to(torch.bfloat16).to(torch.float32)
    return lambda t: t


def mlp_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, act_round: str = "none") -> torch.Tensor:
    """create_mlp in eval mode: Linear (+GELU between layers); keys '0','3','6',... (Dropout/GELU take indices)."""
    rnd = _rounder(act_round)
    idx = sorted({int(k.split(".")[0]) for k in sd})
    h = rnd(x.float())
    for n, i in enumerate(idx):
        h = h @ sd[f"{i}.weight"].float().t()
        if f"{i}.bias" in sd:
            h = h + rnd(sd[f"{i}.bias"].float())
        if n < len(idx) - 1:
            h = F.gelu(h)
        h = rnd(h)
    return h


def splice(ids, table, prot_idx, soft, ret_idx, roll_num=0, struct_idx=None, struct_tokens=None, drug_idx=None,
           drug_tokens=None):
    z = table[ids].clone()
    if soft is not None:
        m = ids == prot_idx
        assert int(m.sum()) == soft.shape[0]
        z[m] = soft.to(z.dtype)
    if struct_tokens:
        m = ids == struct_idx
        for i in range(ids.shape[0]):
            if m[i].sum() > 0:
                z[i, m[i]] = struct_tokens[i].to(z.dtype)
    if drug_tokens is not None:
        m = ids == drug_idx
        z[m] = drug_tokens.to(z.dtype)
    ret = ids == ret_idx
    if roll_num != 0:
        ret = ret.roll(roll_num, 1)
    return z, ret


def make_labels(ids, pad_id, special_ids: List[int], answer_idx, train_qa_full_lm=False):
    lab = ids.clone()
    mask = lab == pad_id
    for s in special_ids:
        mask |= lab == s
    mask[:, -1] = True
    if not train_qa_full_lm:
        for i in range(lab.shape[0]):
            pos = (lab[i] == answer_idx).nonzero()
            last = int(pos.max())
            mask[i, : last + 1] = True
    return torch.where(mask, -100, lab)


def infonce(zs, zt, temperature=0.07, all_s=None, all_t=None, mask=None, rank=0):
    zs, zt = F.normalize(zs.float(), dim=-1), F.normalize(zt.float(), dim=-1)
    b = zs.shape[0]
    if all_s is None:
        sim_st = zs @ zt.t() / temperature
        sim_ts = sim_st.t()
        tgt = torch.arange(b)
    else:
        sim_st = zs @ all_t.t() / temperature
        sim_ts = zt @ all_s.t() / temperature
        tgt = rank * b + torch.arange(b)
        if mask is not None:
            rows = mask[rank * b + torch.arange(b)].float()
            sim_st, sim_ts = sim_st * rows, sim_ts * rows
    return (F.cross_entropy(sim_st, tgt) + F.cross_entropy(sim_ts, tgt)) / 2.0


def conflict_matrix(text_ids, prot_ids, aaseq_code=0, dset_ids=None, protein_dataset_id=4):
    """`negatives_mask` of a retrieval batch, procyon/model/model_unified.py:596-684, on the (already all-gathered) id
    vectors: True = usable negative.  A pair conflicts when it shares the text id but not the protein id, or (same
    amino-acid-sequence type) shares the protein id but not the text id; with `dataset_id`s, text conflicts only count
    inside one dataset — and the reference then clears them for every pair whose two samples are both PPI or both
    non-PPI (`ppi_dset_matrix` compares the two PPI flags for EQUALITY, :670-678), a quirk kept here."""
    n = prot_ids.shape[0]
    code = torch.full((n,), aaseq_code)
    same_type = code[None, :] == code[:, None]

    def conflict(a, b):
        return (a[None, :] == a[:, None]) & ~(b[None, :] == b[:, None])

    text_c = conflict(text_ids, prot_ids)
    prot_c = same_type & conflict(prot_ids, text_ids)
    if dset_ids is not None:
        text_c = (dset_ids[None, :] == dset_ids[:, None]) & text_c
        ppi = dset_ids == protein_dataset_id
        text_c[ppi[None, :] == ppi[:, None]] = False
    return ~(text_c | prot_c)


def cosine_scores(q, db):
    return F.normalize(q.float(), dim=-1) @ F.normalize(db.float(), dim=-1).t()

"""Host-side glue on the hot path, same names and semantics as procyon/training/train_utils.py.

Only the functions the fusion forward path calls are mirrored here:
  batched_split_long_seq   (reference: procyon/training/train_utils.py:1497-1596)
  reverse_batched_split    (reference: :1599-1649)
  concat_tensor_dict, unwrap_model, barrier   (reference: :1694-1729, :1754-1756)
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


def batched_split_long_seq(
    toks: torch.Tensor,
    padding_idx: int,
    eos_idx: int,
    long_protein_strategy: str = "split",
    max_protein_len: int = 1024,
):
    """Chunk proteins longer than `max_protein_len` residues into extra batch rows.

    Same return contract as the reference: (new_toks[:, :max_len+2], batch_keys, eos_loc); for
    'truncate' batch_keys and eos_loc are None.  Differences, all deliberate: the input tensor is not
    mutated, and the common no-split case costs a single device sync instead of one `.item()` per row.
    """
    if long_protein_strategy == "truncate":
        if toks.shape[1] > max_protein_len + 2:
            new_toks = toks[:, : max_protein_len + 2].clone()
            no_pad = new_toks[:, -1] != padding_idx
            # the reference writes `new_toks[:, no_pad] = eos_idx` (indexes columns with a row mask); the evident
            # intent — end truncated rows with EOS — is what is implemented here.
            new_toks[no_pad, -1] = eos_idx
        else:
            new_toks = toks
        return new_toks, None, None
    if long_protein_strategy != "split":
        raise ValueError(f"unknown long_protein_strategy {long_protein_strategy!r}")

    B, W = toks.shape
    is_eos = toks == eos_idx
    eos_pos = is_eos.int().argmax(dim=1)
    n_eos = is_eos.sum(dim=1)
    stats = torch.stack([eos_pos.max(), n_eos.min(), n_eos.max()]).tolist()  # one sync
    if stats[1] != 1 or stats[2] != 1:
        raise ValueError("every protein row must contain exactly one EOS token")
    if stats[0] <= max_protein_len + 1:
        keys = torch.arange(B, dtype=torch.int64)
        return toks[:, : max_protein_len + 2], keys, eos_pos.tolist() if B <= 4096 else eos_pos.cpu().tolist()

    # slow path (proteins > max_protein_len residues): do the bookkeeping on the host
    dev = toks.device
    t = toks.detach().cpu().clone()
    eos_list = eos_pos.cpu().tolist()
    cls_idx = int(t[0, 0])
    batch_keys = list(range(B))
    extra = []
    for i in range(B):
        if eos_list[i] <= max_protein_len + 1:
            continue
        num_add = eos_list[i] // (max_protein_len + 1)
        for j in range(num_add):
            bot = (j + 1) * max_protein_len + 1
            new = torch.full((W,), padding_idx, dtype=t.dtype)
            tail = t[i, bot:]
            new[1 : tail.shape[0] + 1] = tail
            new[0] = cls_idx
            if j < num_add - 1:
                new[max_protein_len + 1] = eos_idx
                new[max_protein_len + 2 :] = padding_idx
            extra.append(new)
            batch_keys.append(i)
        t[i, max_protein_len + 2 :] = padding_idx
        t[i, max_protein_len + 1] = eos_idx
    new_toks = torch.cat([t, torch.stack(extra)], dim=0)[:, : max_protein_len + 2]
    return new_toks.to(dev), torch.tensor(batch_keys, dtype=torch.int64), eos_list


def reverse_batched_split(protein_embeds: torch.Tensor, batch_keys: torch.Tensor, eos_locs: List[int]):
    """Stitch chunk rows back into one token sequence per protein (drops the inner CLS/EOS rows)."""
    max_ind = int(batch_keys.max())
    d = protein_embeds.shape[-1]
    T = protein_embeds.shape[1]
    max_size = max(eos_locs) + 1
    out = []
    for i in range(max_ind + 1):
        idx = (batch_keys == i).nonzero(as_tuple=True)[0].sort()[0]
        if idx.numel() == 0:
            continue
        cp = protein_embeds[idx.to(protein_embeds.device)]
        keep = torch.ones(cp.shape[0], T, dtype=torch.bool)
        keep[:-1, -1] = False
        keep[1:, 0] = False
        cp = cp.reshape(-1, d)[keep.flatten().to(cp.device)]
        if cp.shape[0] < max_size:
            cp = torch.cat([cp, torch.zeros(max_size - cp.shape[0], d, device=cp.device, dtype=cp.dtype)], dim=0)
        else:
            cp = cp[:max_size]
        out.append(cp)
    return torch.stack(out)


def concat_tensor_dict(Ld, dict_keys=None):
    if dict_keys is None:
        dict_keys = Ld[0].keys()
    out = {}
    for k in dict_keys:
        vals = [d[k] for d in Ld]
        if isinstance(vals[0], dict):
            out[k] = {kk: torch.cat([v[kk] for v in vals], dim=0) for kk in vals[0]}
        else:
            out[k] = torch.cat(vals, dim=0)
    return out


def unwrap_model(model):
    return model.module if hasattr(model, "module") else model


def barrier():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()


# ---- QA / retrieval scoring of a forward pass ------------------------------------------
def decompose_dataset_name(name: str):
    """'{aaseq_type}_{text_type}_{relation...}' -> its three parts (procyon/training/train_utils.py:1741-1748)."""
    parts = name.split("_")
    return parts[0], parts[1], "_".join(parts[2:])


def get_final_tokens(text_toks: torch.Tensor, padding_token: int) -> torch.Tensor:
    """Index of the last token before EOS in right-padded rows (train_utils.py:1095-1101): S - n_pad - 2."""
    n_pad = (text_toks == padding_token).sum(dim=-1)
    return text_toks.shape[1] - n_pad - 2


def get_after_answer_tokens(text_toks: torch.Tensor, answer_token: int, get_final: bool = True) -> torch.Tensor:
    """Position right after the (last) [ANSWER] token of every row (train_utils.py:1104-1117), without the
    per-row Python loop of the reference."""
    is_ans = text_toks == answer_token
    if not get_final:
        return is_ans.nonzero()[:, 1] + 1
    if not bool(is_ans.any(dim=1).all()):
        raise ValueError("a row without an [ANSWER] token")  # the reference fails on .max() of an empty tensor
    cols = torch.arange(text_toks.shape[1], device=text_toks.device)
    return (is_ans * cols).max(dim=1).values + 1


def _answer_positions(model_out, padding_token=None, answer_token=None):
    y_all = model_out["text_toks"].detach()
    if padding_token is not None:
        idx = get_final_tokens(y_all, padding_token)
    elif answer_token is not None:
        idx = get_after_answer_tokens(y_all, answer_token)
    else:
        raise ValueError("One of padding_token or answer_token for get_qa_metrics must not be None")
    return y_all, idx


def get_qa_scores(model_out, padding_token=None, answer_token=None):
    """(pred_toks, y_toks): the model's argmax token at the answer position and the label there
    (train_utils.py:1048-1070).  The prediction for position i is read at i - 1 (causal shift).  Only the B rows
    that matter go through the LM head (argmax of softmax == argmax of logits)."""
    y_all, idx = _answer_positions(model_out, padding_token, answer_token)
    rows = torch.arange(y_all.shape[0], device=y_all.device)
    y_toks = y_all[rows, idx]
    logits = model_out["outputs"].logits_at(idx - 1)
    return logits.argmax(dim=-1).cpu(), y_toks.cpu()


def get_qa_logits_inference(model_out, padding_token=None, answer_token=None):
    """(probabilities [B, V] at the answer position, labels) as procyon/data/inference_utils.py:581-604."""
    y_all, idx = _answer_positions(model_out, padding_token, answer_token)
    rows = torch.arange(y_all.shape[0], device=y_all.device)
    probs = model_out["outputs"].logits_at(idx - 1).softmax(dim=-1)
    return probs.cpu(), y_all[rows, idx].cpu()


def get_qa_metrics_from_preds(pred_toks, y_toks, yes_token: int, no_token: int, padding_token=None):
    """accuracy and macro-F1 of yes/no predictions (train_utils.py:1167-1190)."""
    from sklearn.metrics import f1_score

    n_yes_no = int((y_toks == yes_token).sum() + (y_toks == no_token).sum())
    assert n_yes_no == y_toks.numel(), \
        f"Tokens in y other than yes/no ({y_toks.numel() - n_yes_no} of {y_toks.numel()})"
    acc = (pred_toks == y_toks).float().mean()
    f1 = f1_score(y_toks.numpy(), pred_toks.numpy(), average="macro")
    return acc, f1


def get_qa_metrics(model_out, yes_token: int, no_token: int, padding_token=None, answer_token=None):
    """train_utils.py:1120-1164."""
    pred, y = get_qa_scores(model_out, padding_token=padding_token, answer_token=answer_token)
    return get_qa_metrics_from_preds(pred, y, yes_token, no_token)


def get_retrieval_scores_inbatch(cdict):
    """Cosine scores of the in-batch pairs: diagonal = positives, off-diagonal = negatives
    (train_utils.py:996-1019)."""
    import torch.nn.functional as F

    s = F.normalize(cdict["positive"]["sequence"].detach().float().cpu(), dim=-1)
    t = F.normalize(cdict["positive"]["text"].detach().float().cpu(), dim=-1)
    scores = s @ t.t()
    n = scores.shape[0]
    off = ~torch.eye(n, dtype=torch.bool)
    return torch.diagonal(scores).clone(), scores[off]


def get_cl_metrics(pos_scores, neg_scores):
    """(n_pos, n_neg, AUROC, AUPRC) of positive vs negative scores (train_utils.py:966-978)."""
    import numpy as np
    from sklearn.metrics import average_precision_score, roc_auc_score

    pos, neg = np.asarray(pos_scores), np.asarray(neg_scores)
    labels = np.concatenate([np.ones(len(pos)), np.zeros(len(neg))])
    both = np.concatenate([pos, neg])
    return len(pos), len(neg), roc_auc_score(labels, both), average_precision_score(labels, both)

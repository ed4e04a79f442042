"""Forward half of the reference trainer's loss functions:
`ProCyonTrainer.compute_lm_loss` / `compute_retrieval_loss` (procyon/training/trainIT.py:1195-1304) — the model call,
the per-task loss weight and the batch metrics.  The optimiser, DeepSpeed engine, wandb and logger plumbing of the
trainer are out of scope; callers that want the log lines pass `log=`.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

from .train_utils import (decompose_dataset_name, get_cl_metrics, get_qa_metrics, get_retrieval_scores_inbatch)

# text-type ids used to keep negatives within one knowledge source (procyon/data/constants.py:666-680)
DATASET_ID = {"go": 0, "pfam": 1, "disgenet": 2, "reactome": 3, "protein": 4, "omim": 5, "drugbank": 6,
              "drugbank:moa": 6, "drugbank:indication": 6, "gtop": 7, "ec": 8, "uniprot": 9, "peptide": 10}


def compute_lm_loss(model, inputs, task_type: str, args, dataset_key: Optional[str] = None,
                    caption_loss_rescale: Optional[Dict[str, float]] = None,
                    log: Optional[Callable[[str, float], None]] = None) -> torch.Tensor:
    """LM loss of a QA or caption batch times its task weight (trainIT.py:1195-1262).

    `args` supplies `qa_loss_weight` / `caption_loss_weight` (TrainArgs); metrics go to `log(name, value)`."""
    prefix = "batch_train" if dataset_key is None else f"{dataset_key}_batch_train"
    task_prefix = f"{prefix}_{task_type}"
    aaseq_type, text_type, _ = decompose_dataset_name(dataset_key) if dataset_key else ("protein", None, None)
    out = model(inputs, return_mlm=False, retrieval=False, get_full_labels=True, aaseq_type=aaseq_type,
                crop_off=(task_type == "caption"))
    loss = out["outputs"].loss
    if task_type == "qa":
        acc, f1 = get_qa_metrics(out, yes_token=model.yes_token, no_token=model.no_token,
                                 answer_token=model.answer_idx)
        if log:
            log(f"{task_prefix}_acc", float(acc))
            log(f"{task_prefix}_f1", float(f1))
        weight = args.qa_loss_weight
    elif task_type == "caption":
        weight = args.caption_loss_weight
        if caption_loss_rescale is not None:
            weight = weight * caption_loss_rescale[f"{aaseq_type}_{text_type}"]
    else:
        raise ValueError(f"task_type {task_type!r}: expected 'qa' or 'caption'")
    if log:
        log(f"{task_prefix}_ppl", math.exp(float(loss)))
        log(f"{task_prefix}_loss", float(loss))
    return loss * weight


def compute_retrieval_loss(model, inputs, args, model_args=None, task_type: str = "retrieval",
                           dataset_key: Optional[str] = None,
                           log: Optional[Callable[[str, float], None]] = None) -> torch.Tensor:
    """In-batch contrastive loss of a retrieval batch times `retrieval_loss_weight` (trainIT.py:1264-1304)."""
    aaseq_type, text_type, _ = decompose_dataset_name(dataset_key) if dataset_key else ("protein", None, None)
    if model_args is not None and getattr(model_args, "filter_negatives_by_id_contrastive", False) and text_type:
        inputs["dataset_id"] = torch.full((len(inputs["input"]["text"]),), DATASET_ID[text_type], dtype=torch.long)
    out = model(inputs, return_mlm=False, retrieval=True, aaseq_type=aaseq_type)
    loss = out["contrastive_loss"]
    if log:
        prefix = f"batch_train_{task_type}" if dataset_key is None else f"{dataset_key}_batch_train_{task_type}"
        pos, neg = get_retrieval_scores_inbatch(out["contrastive_out"])
        _, _, auroc, auprc = get_cl_metrics(pos.numpy(), neg.numpy())
        log(f"{prefix}_loss", float(loss.mean()))
        log(f"{prefix}_auroc", float(auroc))
        log(f"{prefix}_auprc", float(auprc))
    return loss * args.retrieval_loss_weight

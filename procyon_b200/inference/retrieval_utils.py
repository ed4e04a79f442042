"""Inference loops — host-side mirror of procyon/inference/retrieval_utils.py.

load_model_onto_device  reference :74-106   (from_pretrained -> .bfloat16() -> .eval() -> .to(device))
do_retrieval            reference :109-201  (model(inputs, retrieval=True) -> cosine scores over the protein DB)
The string/template side (`create_input_retrieval`, pandas lookups) stays with the reference's data package; here
`do_retrieval` takes the already-built model-input dict.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from ..data.inference_utils import get_proteins_from_embedding
from ..model.model_unified import UnifiedProCyon


def load_model_onto_device(checkpoint_dir: Optional[str] = None, **model_kwargs):
    ckpt = checkpoint_dir or os.getenv("CHECKPOINT_PATH")
    from .. import compat

    compat.install()
    data_args = torch.load(os.path.join(ckpt, "data_args.pt"), weights_only=False)
    model, _ = UnifiedProCyon.from_pretrained(checkpoint_dir=ckpt, **model_kwargs)
    model.bfloat16()
    model.eval()
    if not torch.cuda.is_available():
        raise RuntimeError("procyon_b200 needs a CUDA device (there is no CPU path)")
    device = torch.device("cuda")
    model.to(device)
    return model, device, data_args


def load_protein_target_embeddings(checkpoint_dir: str):
    """`protein_target_embeddings.pkl` = torch.save((Tensor[N,d], ids)) (reference :61-64)."""
    emb, ids = torch.load(os.path.join(checkpoint_dir, "protein_target_embeddings.pkl"), weights_only=False)
    return emb.float(), ids


@torch.no_grad()
def do_retrieval(model: UnifiedProCyon, input_dict, all_protein_embeddings: torch.Tensor, protein_ids=None,
                 top_k: Optional[int] = 20, aaseq_type: str = "protein"):
    model_out = model(inputs=input_dict, retrieval=True, aaseq_type=aaseq_type)
    return get_proteins_from_embedding(all_protein_embeddings, model_out, protein_ids=protein_ids, top_k=top_k)

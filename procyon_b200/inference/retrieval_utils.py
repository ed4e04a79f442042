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


@torch.no_grad()
def build_protein_target_embeddings(model: UnifiedProCyon, protein_tokens: torch.Tensor, batch_size: int = 256):
    """The retrieval database: `forward_sequences(tokens)["shared"]` for every protein (what the reference's
    evaluation loop gathers, procyon/training/trainIT.py:1596-1610), encoded in contiguous blocks per rank and
    all-gathered when torch.distributed is initialised.  Returns fp32 [N, d] on the model's device."""
    from .sharded import encode_proteins_sharded

    dev = model.input_embeddings.weight.device
    enc = lambda t: model.forward_sequences(t)["shared"]
    outs = []
    for i in range(0, protein_tokens.shape[0], batch_size * max(1, _world_size())):
        outs.append(encode_proteins_sharded(enc, protein_tokens[i:i + batch_size * max(1, _world_size())].to(dev)))
    return torch.cat(outs).float()


def _world_size() -> int:
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def save_protein_target_embeddings(checkpoint_dir: str, embeddings: torch.Tensor, protein_ids) -> str:
    """Writes `protein_target_embeddings.pkl` in the layout the reference loads (`torch.load` -> (Tensor[N,d], ids),
    procyon/inference/retrieval_utils.py:61-64)."""
    path = os.path.join(checkpoint_dir, "protein_target_embeddings.pkl")
    torch.save((embeddings.detach().float().cpu(), protein_ids), path)
    return path

"""Embedding-table readers — host-side mirror of the loaders in procyon/data/data_utils.py:365-398 that
`UnifiedProCyon.__init__` calls (model_unified.py:189-211, 269-297) for ProCyon-Full checkpoints:

  load_aaseq_embeddings             pre-computed protein / domain / peptide embeddings (ESM2 pooled), re-ordered to the
                                    row order of `<aaseq_type>_info_filtered.pkl`
  load_protein_struct_embeddings    GearNet structure embeddings, one row per protein index
  load_drug_structure_embeddings    drug structure embeddings, one row per drug index

The files are the reference's own (`DATA_DIR/generated_data/node_embeddings/...`, `DATA_DIR/integrated_data/v1/...`);
only reading and re-ordering happens here, on the host, once per model construction.
"""
from __future__ import annotations

import os
import pickle
from typing import Optional

import torch


def data_dir(explicit: Optional[str] = None) -> str:
    d = explicit or os.environ.get("DATA_DIR")
    if not d:
        raise RuntimeError("DATA_DIR is not set (the reference reads it from the environment / .env, "
                           "procyon/data/data_utils.py:19-26)")
    return d


def _load_tensor(path: str) -> torch.Tensor:
    with open(path, "rb") as fh:
        return torch.load(fh, map_location="cpu", weights_only=False)


def load_protein_struct_embeddings(protein_struct_embeddings_path: str) -> torch.Tensor:
    return _load_tensor(protein_struct_embeddings_path)


def load_drug_structure_embeddings(drug_struct_embeddings_path: str) -> torch.Tensor:
    return _load_tensor(drug_struct_embeddings_path)


def load_aaseq_embeddings(aaseq_embeddings_path: str, aaseq_embeddings_idmap_path: str, aaseq_type: str,
                          data_dir_override: Optional[str] = None) -> torch.Tensor:
    """Row i of the result is the embedding of the sequence whose `index` in `<type>_info_filtered.pkl` is i.

    The embedding file is in the order of its id map (a pickled list of "<id> <description...>" strings); the info
    table maps `<type>_id` -> `index`. Every id of the map must be in the table (KeyError otherwise, as in the
    reference's `.loc`)."""
    import pandas as pd

    emb = _load_tensor(aaseq_embeddings_path)
    info = pd.read_pickle(os.path.join(data_dir(data_dir_override),
                                       f"integrated_data/v1/{aaseq_type}/{aaseq_type}_info_filtered.pkl"))
    with open(aaseq_embeddings_idmap_path, "rb") as fh:
        id_map = [s.split(" ")[0] for s in pickle.load(fh)]
    index_of = dict(zip(info[f"{aaseq_type}_id"].tolist(), info["index"].tolist()))
    idx = torch.tensor([index_of[a] for a in id_map], dtype=torch.int64)  # table index of every embedding row
    # rows sorted by table index (stable, like pandas' sort_values on the re-indexed frame)
    order = torch.sort(idx, stable=True).indices
    return emb[order]

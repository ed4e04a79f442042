"""`ModelArgs` fields that steer the fusion forward path (mirror of procyon/training/training_args_IT.py:26-651).

Only the fields `UnifiedProCyon` reads are declared (defaults copied from the reference dataclass; see SURVEY
Appendix B). Pickled reference `ModelArgs` instances (checkpoint `model_args.pt`) carry many more attributes:
unpickling restores `__dict__` wholesale, so they load into this class unchanged once `procyon_b200.compat`
has aliased the reference's dotted path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional


@dataclass
class ModelArgs:
    # ---- protein encoder ----
    protein_encoder_num_params: str = "650m"
    protein_encoder_debug: bool = False
    protein_tokenizer_name: str = "ESM-1b"
    max_protein_len: int = 1024
    long_protein_strategy: str = "split"
    protein_pooling_opt: str = "max"
    protein_enc_batch_limit: Optional[int] = None
    protein_pooling_correction_option: bool = False
    freeze_protein_encoder: Optional[str] = None
    protein_task_spc_lora: bool = False
    protein_task_spc_lora_num: int = 2
    protein_lora_parameters: str = "default"
    aaseq_lora_alpha: int = 8
    aaseq_lora_r: int = 8
    aaseq_adapter_rank: int = 8
    lora_specific_style: str = "specific"
    # ---- pre-computed amino-acid-sequence embeddings (ProCyon-Full) ----
    use_aaseq_embeddings: bool = False
    freeze_aaseq_embeddings: bool = False
    protein_seq_embeddings_path: Optional[str] = None
    domain_embeddings_path: Optional[str] = None
    peptide_embeddings_path: Optional[str] = None
    protein_embeddings_idmap_path: Optional[str] = None  # default: DATA_DIR/generated_data/node_embeddings/protein/...
    domain_embeddings_idmap_path: Optional[str] = None
    peptide_embeddings_idmap_path: Optional[str] = None
    # ---- structure / drug soft tokens ----
    use_protein_struct: bool = False
    protein_struct_dropout: float = 0.5
    protein_struct_embeddings_path: Optional[str] = None
    use_drug_embeddings: bool = False
    drug_struct_embeddings_path: Optional[str] = None
    # ---- text encoder ----
    text_encoder_fname: str = "llama-3-8b"
    text_encoder_debug: bool = False
    max_text_len: int = 1024
    freeze_text_encoder: Optional[str] = None
    attention_type: str = "vanilla"
    model_splitting: bool = False
    n_model_pieces: int = 2
    use_lora: bool = False
    text_task_spc_lora: bool = False
    text_task_spc_lora_num: int = 2
    streaming_llm_max_gen_len: int = 50
    # ---- projectors ----
    num_layers_token_projector: int = 1
    num_layers_shared_projector: int = 1
    num_layers_lm_projector: int = 1
    hidden_size_token_projector: int = 256
    hidden_size_shared_projector: int = 256
    hidden_size_lm_projector: int = 256
    # ---- retrieval / contrastive ----
    ret_token_access: str = "all"
    roll_num: int = -1
    cl_method: str = "infonce"
    use_projection_cl: bool = False
    contrastive_global: bool = False
    filter_negatives_by_id_contrastive: bool = False
    negative_sampling_strategy_retrieval: str = "in_batch"
    # ---- language-model loss ----
    causal_qa: bool = True
    train_qa_full_lm: bool = False
    context_crop_sampling: bool = False
    enforce_checkpoint_architecture_strict: bool = False


def full_model_args(**overrides) -> ModelArgs:
    """ProCyon-Full settings (configs/llama3-full.yml) with a live ESM2 encoder unless overridden."""
    base = dict(
        protein_encoder_num_params="3b", protein_pooling_opt="mean", text_encoder_fname="llama-3-8b",
        max_text_len=2048, num_layers_token_projector=3, num_layers_shared_projector=3, num_layers_lm_projector=3,
        hidden_size_token_projector=2560, hidden_size_shared_projector=2560, hidden_size_lm_projector=2560,
        ret_token_access="last", roll_num=0, train_qa_full_lm=False, contrastive_global=True,
        filter_negatives_by_id_contrastive=True,
    )
    base.update(overrides)
    return ModelArgs(**base)


class DataArgs:
    """Placeholder so pickled `data_args.pt` from reference checkpoints can be loaded (attributes restored as-is)."""

    data_dir = None


def _current_data_dir():
    import os

    return os.environ.get("DATA_DIR")


def update_model_args_data_dir(model_args: ModelArgs, prev_data_dir: str):
    """procyon/training/training_args_IT.py:1787-1801: a checkpoint's `model_args.pt` stores every `*_path` field under
    the DATA_DIR of the machine it was trained on; swap that prefix for the current DATA_DIR.  Works on the instance
    `__dict__` (an unpickled reference ModelArgs carries ~90 fields this mirror class does not declare)."""
    if not isinstance(model_args, ModelArgs):
        raise ValueError(f"expected ModelArgs, got: {type(model_args)}")
    data_dir = _current_data_dir()
    if data_dir is None or prev_data_dir is None or data_dir == prev_data_dir:
        return
    import os

    for name, cur in list(vars(model_args).items()):
        if name.endswith("path") and isinstance(cur, str) and cur.startswith(prev_data_dir):
            suffix = cur.replace(prev_data_dir, "").lstrip("/")
            setattr(model_args, name, os.path.join(data_dir, suffix))


def update_data_args_data_dir(data_args):
    """procyon/training/training_args_IT.py:1803-1811."""
    data_dir = _current_data_dir()
    if data_dir is not None and getattr(data_args, "data_dir", None) != data_dir:
        data_args.data_dir = data_dir


class TrainArgs:
    """Placeholder so pickled `training_args.pt` from reference checkpoints can be loaded."""

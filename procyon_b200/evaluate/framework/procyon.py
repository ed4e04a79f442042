"""Evaluation plugins — the callers the reference's `evaluate/framework/core.py` instantiates by name.

Mirror of procyon/evaluate/framework/procyon.py ("Eval plugin"):

  ProcyonCaptionEval    reference :49-111   get_predictions(data_loader) -> DataFrame{seq_id, generated_caption}
  ProcyonQAEval         reference :114-211  get_predictions(data_loader, aaseq_type) -> {pred, y, seq_ids, text_ids}
  ProcyonRetrievalEval  reference :214-406  get_predictions(query_loader, target_loader, query_order, target_order)
                                            -> Tensor[num_queries, num_targets] float64 (CPU)

Same constructor `(model_config, eval_args, model_args, device)`, same `get_predictions` contracts, same cache file
(`<checkpoint_dir>/<aaseq_type>_target_embeddings.pkl` = `torch.save((Tensor[N,d], ids))`).  What differs is where
the work runs: embeddings stay on the GPU between the encode and the scoring, and the (Q, N) cosine matrix comes
from the streaming scoring kernel (`pcy_cosine_scores`) instead of two `F.normalize` passes and a CPU matmul
(reference :398-404).  The loaders are duck-typed — anything iterable that yields the collators' model-input dicts
and carries `.dataset.aaseq_type` / `.collate_fn` works — so the reference's DataLoaders plug in unchanged, and the
data package itself (datasets, collators, metrics) stays with the reference.

`model_config` may hold a ready `UnifiedProCyon` under "model" (serving / tests) instead of "checkpoint_dir".
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass, fields, is_dataclass
from typing import Callable, Dict, Iterable, List, Mapping, Optional, Tuple

import numpy as np
import torch

from ...data.inference_utils import cosine_scores
from ...model.model_unified import UnifiedProCyon
from ...training.train_utils import get_qa_scores


@dataclass
class EvalArgs:
    """The fields of procyon/evaluate/framework/args.py:EvalArgs that the three plugins read."""

    batch_size: int = 16
    seed: int = 42
    caption_max_len: int = 200
    qa_num_samples: Optional[int] = None
    retrieval_use_cached_target_embeddings: bool = False
    retrieval_eval_all_aaseqs: bool = False


def move_inputs_to_device(data, device):
    """Nested dict / list / tuple of tensors -> same structure on `device` (reference utils.py:46-61)."""
    if isinstance(data, torch.Tensor):
        return data.to(device=device)
    if isinstance(data, Mapping):
        return type(data)({k: move_inputs_to_device(v, device) for k, v in data.items()})
    if isinstance(data, (list, tuple)):
        return type(data)(move_inputs_to_device(v, device) for v in data)
    return data


def compare_and_warn_model_args(model_args_a, model_args_b, ignore=("n_model_pieces", "model_splitting")) -> List[tuple]:
    """Warns about evaluation-time ModelArgs that differ from the checkpoint's (reference utils.py:103-140); fields
    rewritten by `from_pretrained` and `*path` fields are skipped.  Returns the mismatches."""
    if model_args_a is None or model_args_b is None or not is_dataclass(model_args_b):
        return []
    diff = []
    for f in fields(model_args_b):
        if f.name in ignore or f.name.endswith("path"):
            continue
        a, b = getattr(model_args_a, f.name, None), getattr(model_args_b, f.name, None)
        if a != b:
            diff.append((f.name, a, b))
    if diff:
        warnings.warn("model args differ from the checkpoint's: " + ", ".join(f"{n}: {a!r} != {b!r}" for n, a, b in diff))
    return diff


class _PluginBase:
    strict_load = False

    def __init__(self, model_config: Dict, eval_args, model_args, device):
        if not torch.cuda.is_available() or torch.device(device).type != "cuda":
            raise RuntimeError("procyon_b200 evaluation plugins need a CUDA device (there is no CPU path)")
        model = model_config.get("model")
        checkpoint_dir = model_config.get("checkpoint_dir")
        if model is None:
            from ...model.model_unified import DEFAULT_PRETRAINED_WEIGHTS_DIR

            model, ckpt_args = UnifiedProCyon.from_pretrained(pretrained_weights_dir=DEFAULT_PRETRAINED_WEIGHTS_DIR,
                                                              checkpoint_dir=checkpoint_dir,
                                                              strict_load=self.strict_load)
            compare_and_warn_model_args(model_args, ckpt_args)
        model.eval()
        model.bfloat16()
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.model_args = model_args
        self.checkpoint_dir = checkpoint_dir


def _last(indices):
    return indices[-1]


class ProcyonCaptionEval(_PluginBase):
    """Phenotype / caption generation for every sample of a loader (reference :49-111)."""

    def __init__(self, model_config: Dict, eval_args, model_args, device):
        super().__init__(model_config, eval_args, model_args, device)
        self.max_len = eval_args.caption_max_len
        self.method = model_config.get("generation_method", "beam")
        self.num_captions = model_config.get("num_captions", 5)
        self.beam_group_size = model_config.get("beam_group_size", 2)
        self.beam_size = model_config.get("beam_size", self.num_captions * self.beam_group_size)

    @torch.no_grad()
    def get_predictions(self, data_loader: Iterable):
        import pandas as pd

        aaseq_type = getattr(getattr(data_loader, "dataset", None), "aaseq_type", "protein")
        seq_ids, captions = [], []
        for batch in data_loader:
            batch = move_inputs_to_device(batch, self.device)
            _, _, _, texts = self.model.generate(batch, max_len=self.max_len, aaseq_type=aaseq_type,
                                                 return_all_internals=False, method=self.method,
                                                 beam_size=self.beam_size, beam_group_size=self.beam_group_size,
                                                 truncate_on_eos=True)
            # one caption per beam GROUP: the best beam of group j sits at index j * beam_group_size (:104-106)
            for per_input, ref in zip(texts, batch["reference_indices"]["input"]["seq"]):
                for j in range(self.num_captions):
                    seq_ids.append(_last(ref))
                    captions.append(per_input[j * self.beam_group_size])
        return pd.DataFrame({"seq_id": seq_ids, "generated_caption": captions})


class ProcyonQAEval(_PluginBase):
    """Yes/no question answering: the token predicted right after [ANSWER] (reference :114-211).  The forward only
    evaluates the LM head on the answer rows (`get_qa_scores` -> `outputs.logits_at`)."""

    def __init__(self, model_config: Dict, eval_args, model_args, device):
        super().__init__(model_config, eval_args, model_args, device)
        self.num_samples = eval_args.qa_num_samples
        self.rng = np.random.default_rng(seed=eval_args.seed)
        self.yes_token = self.model.yes_token
        self.no_token = self.model.no_token

    @torch.no_grad()
    def get_predictions(self, data_loader: Iterable, aaseq_type: str = "protein") -> Dict:
        keep = None
        if self.num_samples is not None and self.num_samples < len(data_loader):
            keep = set(self.rng.choice(np.arange(len(data_loader)), size=self.num_samples, replace=False).tolist())
        # with context augmentation the query text is the second-to-last text of a sample (:163-167)
        collate = getattr(data_loader, "collate_fn", None)
        ctx = collate._get_input_contexts([], []) if hasattr(collate, "_get_input_contexts") else None
        text_pos = -1 if ctx is None else -2

        res = {"seq_ids": [], "text_ids": [], "pred": [], "y": []}
        for i, batch in enumerate(data_loader):
            if keep is not None and i not in keep:
                continue
            out = self.model(move_inputs_to_device(batch, self.device), return_mlm=False, retrieval=False,
                             get_full_labels=True, aaseq_type=aaseq_type, crop_off=True)
            pred, _ = get_qa_scores(out, answer_token=self.model.answer_idx)
            res["seq_ids"].extend(_last(x) for x in batch["reference_indices"]["input"]["seq"])
            res["text_ids"].extend(x[text_pos] for x in batch["reference_indices"]["input"]["text"])
            res["pred"].append(pred)
            # the instructions of the eval split carry no answer word: labels come from target.text (:187-191)
            res["y"].append(torch.tensor([self.yes_token if y == "yes" else self.no_token
                                          for y in batch["target"]["text"]], dtype=torch.int64))
        res["pred"] = torch.cat(res["pred"]) if res["pred"] else torch.empty(0, dtype=torch.int64)
        res["y"] = torch.cat(res["y"]) if res["y"] else torch.empty(0, dtype=torch.int64)
        return res


def _is_ppi(dataset) -> bool:
    """Protein queries (AASeqDataset) vs text queries (AASeqTextUnifiedDataset), reference :245-250 — decided by
    class name so that the reference's dataset classes need not be importable here."""
    flag = getattr(dataset, "is_ppi", None)
    if flag is not None:
        return bool(flag)
    names = {c.__name__ for c in type(dataset).__mro__}
    if "AASeqTextUnifiedDataset" in names:
        return False
    if "AASeqDataset" in names:
        return True
    raise ValueError(f"unexpected dataset type: {type(dataset)}")


class ProcyonRetrievalEval(_PluginBase):
    """Text -> protein (or protein -> protein) retrieval scores for a query set against a target set
    (reference :214-406)."""

    def __init__(self, model_config: Dict, eval_args, model_args, device):
        super().__init__(model_config, eval_args, model_args, device)
        self.batch_size = eval_args.batch_size
        self.use_cached_target_embeddings = eval_args.retrieval_use_cached_target_embeddings
        self.all_targets_loader: Optional[Callable[[str], Iterable]] = model_config.get("all_targets_loader")

    @torch.no_grad()
    def _get_query_embeddings(self, query_loader: Iterable, query_order: List) -> torch.Tensor:
        side = "seq" if _is_ppi(query_loader.dataset) else "text"
        chunks, ids = [], []
        for batch in query_loader:
            batch["target"]["seq"] = None  # no target encodes: they only feed the training loss (:258-260)
            batch = move_inputs_to_device(batch, self.device)
            ids.extend(_last(x) for x in batch["reference_indices"]["input"][side])
            out = self.model(batch, retrieval=True, aaseq_type=query_loader.dataset.aaseq_type)
            chunks.append(out["contrastive_out"]["positive"]["text"].detach().float())
        # a query that occurs in several (query, target) relations keeps its LAST embedding (:283-293)
        row_of = {q: i for i, q in enumerate(ids)}
        rows = torch.tensor([row_of[q] for q in query_order], dtype=torch.int64, device=self.device)
        return torch.cat(chunks, dim=0).index_select(0, rows)

    @torch.no_grad()
    def _calculate_target_embeddings(self, target_loader: Iterable, collate_fn, aaseq_type: str = "protein"
                                     ) -> Tuple[torch.Tensor, List]:
        chunks, ids = [], []
        for protein_ids in target_loader:
            ids.extend(protein_ids.tolist() if hasattr(protein_ids, "tolist") else list(protein_ids))
            if self.model.config.use_aaseq_embeddings:
                seqs = protein_ids  # rows of the pre-computed embedding table
            else:
                seqs = collate_fn._convert_batch("sequence", protein_ids)  # ids -> ESM tokens (:311-313)
            seqs = move_inputs_to_device(seqs, self.device)
            chunks.append(self.model.forward_sequences(seqs, aaseq_type=aaseq_type)["shared"].detach().float())
        return torch.cat(chunks, dim=0), ids

    def _get_cached_target_embeddings(self, collate_fn, aaseq_type: str) -> Tuple[torch.Tensor, List]:
        path = os.path.join(self.checkpoint_dir, f"{aaseq_type}_target_embeddings.pkl")
        if os.path.exists(path):
            emb, ids = torch.load(path, weights_only=False)
            return emb.to(self.device).float(), list(ids)
        if self.all_targets_loader is None:
            raise FileNotFoundError(
                f"{path} not found and no `all_targets_loader` was given in model_config to compute it (the reference "
                "enumerates every protein through its data package, evaluate/framework/retrieval.py:85-110)")
        emb, ids = self._calculate_target_embeddings(self.all_targets_loader(aaseq_type), collate_fn, aaseq_type)
        torch.save((emb.cpu(), ids), path)
        return emb, ids

    @torch.no_grad()
    def _get_target_embeddings(self, target_loader: Iterable, target_order: List, collate_fn, aaseq_type: str
                               ) -> torch.Tensor:
        if self.use_cached_target_embeddings:
            emb, ids = self._get_cached_target_embeddings(collate_fn, aaseq_type)
        else:
            emb, ids = self._calculate_target_embeddings(target_loader, collate_fn, aaseq_type=aaseq_type)
        row_of = {t: i for i, t in enumerate(ids)}
        rows = torch.tensor([row_of[t] for t in target_order], dtype=torch.int64, device=emb.device)
        return emb.index_select(0, rows)

    @torch.no_grad()
    def get_predictions(self, query_loader: Iterable, target_loader: Iterable, query_order: List,
                        target_order: List) -> torch.Tensor:
        q = self._get_query_embeddings(query_loader, query_order)
        t = self._get_target_embeddings(target_loader, target_order, getattr(query_loader, "collate_fn", None),
                                        query_loader.dataset.aaseq_type)
        return cosine_scores(q, t).detach().cpu().to(torch.float64)


# name -> class, the entries `evaluate/framework/core.py:68-110` registers for this model family
model_zoo = {
    "caption": {"ProCyon": ProcyonCaptionEval},
    "qa": {"ProCyon": ProcyonQAEval},
    "retrieval": {"ProCyon": ProcyonRetrievalEval},
}

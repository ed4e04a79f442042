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


# ---- retrieval query builders (procyon/data/inference_utils.py:663-925) ---------------------------------------------
# Host code: strings and index lists only.  Files are the reference's own: the task descriptions under
# `$HOME_DIR/procyon/data/instruct_tune/tasks/<task_id>.json` and the text tables under
# `$DATA_DIR/integrated_data/v1/<dataset>/<dataset>_info_filtered_composed.pkl`; both roots can also be passed in.

# text columns a retrieval query may be drawn from, per `DataArgs.retrieval_subset_version` (the column names are the
# contract with the reference's data files, procyon/data/constants.py:241-330); versions 2 and 5 differ from 1 only in
# the entries listed
_RETRIEVAL_COLUMNS_V1 = {
    "go": ["description_name_type_def"],
    "pfam": ["description_pfam", "description_interpro"],
    "disgenet": ["description_" + s for s in ("air aot chv csp fma go hl7v3.0 hpo lnc mcm medlineplus msh nci pdq spn uwda "
                                              "primekg_mondo primekg_orphanet").split()],
    "reactome": ["description_name_description"],
    "protein": [None],
    "omim": ["description_" + s for s in "omim mondo umls orphanet mayo".split()],
    "drugbank": ["moa", "indication"],
    "drugbank:moa": ["moa"],
    "drugbank:indication": ["indication"],
    "gtop": ["description_name_" + s for s in "overview comments introduction".split()],
    "ec": ["description_explorenz"],
    "uniprot": ["function"],
}
RETRIEVAL_SUBSETS = {
    1: _RETRIEVAL_COLUMNS_V1,
    2: {**_RETRIEVAL_COLUMNS_V1, "disgenet": ["description_all_collapse"]},
    5: {**_RETRIEVAL_COLUMNS_V1, "disgenet": ["description_all_collapse"], "go": ["go_def"], "reactome": ["description"],
        "omim": ["omim_" + s + "_curated" for s in "def clinical molecular title".split()],
        "gtop": ["target_family_overview", "target_family_comments"]},
}


def construct_task_id(aaseq_type: str, text_type: str, relation_type: str, task_type: str) -> str:
    """procyon/data/it_collator.py:886-897"""
    kind = aaseq_type.lower()
    if kind == "domain":
        return f"domain_{text_type}_{relation_type}_{task_type}"
    if kind in ("protein", "peptide"):
        return f"{text_type}_{relation_type}_{task_type}"
    raise NotImplementedError(f"No dataset found for aaseq_type = {aaseq_type}")


def _first_text(row) -> str:
    """first non-null description among the candidate columns of one table row"""
    vals = row[row.notna()].tolist()
    return vals[0]


def create_input_retrieval(input_description: str, data_args, instruction_source_dataset: str,
                           drug_input_idx=None, task_definition: Optional[str] = None,
                           instruction_source_relation: str = "all", aaseq_type: str = "protein",
                           icl_example_number: int = 1, *, home_dir: Optional[str] = None,
                           data_dir: Optional[str] = None, text_table=None, drug_mask=None) -> Dict:
    """Model-input dict of ONE free-text retrieval query: the task's instruction with `icl_example_number` in-context
    examples (their descriptions come from the dataset's text table, their proteins by index) followed by the query
    text.  Same arguments and output layout as the reference; `home_dir` / `data_dir` / `text_table` / `drug_mask`
    (keyword-only extensions) replace the environment roots and the files read from them."""
    import json
    import os

    from .instruct_tune.instruct_constructor import get_prompt, get_prompt_open_def

    dataset = instruction_source_dataset.lower()
    home = home_dir or os.environ.get("HOME_DIR")
    if home is None:
        raise RuntimeError("HOME_DIR is not set (root of the checkout that holds procyon/data/instruct_tune/tasks)")
    task_id = construct_task_id(aaseq_type, dataset, instruction_source_relation, "retrieval")
    with open(os.path.join(home, "procyon", "data", "instruct_tune", "tasks", f"{task_id}.json")) as fh:
        task = json.load(fh)
    kw = dict(task=task, num_examples=icl_example_number, is_special_definition=False, is_ppi=(dataset == "protein"),
              aaseq_type=aaseq_type)
    if task_definition is None:
        instruction, _, _, ex_text, ex_seq = get_prompt(**kw)
    else:
        instruction, _, _, _, ex_text, ex_seq = get_prompt_open_def(**kw)
        instruction = instruction.format(definition=task_definition)

    if text_table is None:
        import pandas as pd

        root = data_dir or os.environ.get("DATA_DIR")
        if root is None:
            raise RuntimeError("DATA_DIR is not set (root of integrated_data/v1/<dataset>/..._info_filtered_composed.pkl)")
        text_table = pd.read_pickle(os.path.join(root, "integrated_data", "v1", dataset,
                                                 f"{dataset}_info_filtered_composed.pkl"))
    cols = RETRIEVAL_SUBSETS[data_args.retrieval_subset_version][dataset]
    candidates = text_table[cols] if dataset != "protein" else None  # (get_text_sequences_compositions)
    descriptions = [_first_text(candidates.iloc[i, :]) for i in ex_text]

    # optional drug soft tokens: examples that have a structure embedding, then the query's own drug
    slot_of_drug, drug_rows = None, None
    if drug_input_idx is not None:
        if drug_mask is None:
            root = data_dir or os.environ.get("DATA_DIR")
            drug_mask = torch.load(os.path.join(root, "integrated_data/v1/drugbank/drugbank_mask.pt"))
        slot_of_drug, drug_rows = [], []
        if dataset == "drugbank":
            for i, tid in enumerate(ex_text):
                if drug_mask[tid]:
                    descriptions[i] += "\nDrug: <|drug|>"
                    slot_of_drug.append(i)
                    drug_rows.append(tid)
        if bool(drug_mask[drug_input_idx]):
            input_description = input_description + "\nDrug: <|drug|>"
            slot_of_drug.append(len(slot_of_drug))
            drug_rows.append(drug_input_idx)
            slot_of_drug = torch.LongTensor(slot_of_drug).unsqueeze(0).tolist()
            drug_rows = torch.LongTensor(drug_rows)
        else:
            print("WARNING: not inserting drug index because we don't have one in our database")
            slot_of_drug, drug_rows = None, None

    seq_ids = torch.LongTensor(ex_seq) if len(ex_seq) > 0 else None
    texts = descriptions + [input_description]
    if instruction.endswith("[EXT]"):
        instruction = instruction[:-5]
    instruction = instruction.replace("[CONTEXT]", "")
    return {
        "data": {"seq": seq_ids, "seq_idx": seq_ids, "text": texts, "drug": drug_rows},
        "input": {"seq": [list(range(len(ex_seq)))] if seq_ids is not None else None,
                  "text": [list(range(len(texts)))], "drug": slot_of_drug},
        "target": {"seq": None, "text": None, "drug": None},
        "instructions": [instruction],
    }


def create_batched_input_retrieval(input_descriptions, data_args, task_definitions=None,
                                   instruction_source_dataset: Optional[str] = None,
                                   instruction_source_relation: str = "all", aaseq_type: str = "protein",
                                   icl_example_number: int = 1, **roots) -> Dict:
    """One merged model-input dict for several queries (procyon/data/inference_utils.py:886-925)."""
    n = len(input_descriptions)
    if task_definitions is None:
        task_definitions = [None] * n
    assert len(task_definitions) == n
    dicts = [create_input_retrieval(input_description=input_descriptions[i], data_args=data_args,
                                    task_definition=task_definitions[i],
                                    instruction_source_dataset=instruction_source_dataset,
                                    instruction_source_relation=instruction_source_relation, aaseq_type=aaseq_type,
                                    icl_example_number=icl_example_number, **roots) for i in range(n)]
    return merge_model_input_dicts(dicts)


# ---- caption / QA query builders (procyon/data/inference_utils.py:67-420) -----------------------------------------
_V5 = {"disgenet": ["description_all_collapse"], "go": ["go_def"], "reactome": ["description"],
       "omim": ["omim_" + s + "_curated" for s in "def clinical molecular title".split()],
       "gtop": ["target_family_overview", "target_family_comments"]}
_DRUG_QA = {"drugbank": ["indication", "moa"]}
# (procyon/data/constants.py: QA_SUBSETS, CAPTION_SUBSETS - column names of the reference's data files)
QA_SUBSETS = {
    1: {**_RETRIEVAL_COLUMNS_V1, **_DRUG_QA},
    5: {**_RETRIEVAL_COLUMNS_V1, **_V5},
    "ProtLLM": {**_RETRIEVAL_COLUMNS_V1, **_DRUG_QA, "disgenet": ["description_all_collapse"]},
    "ProtLLM_name": {**_RETRIEVAL_COLUMNS_V1, **_DRUG_QA, "disgenet": ["description_all_collapse"], "go": ["go_name"],
                     "ec": ["explorenz_accepted_name"]},
}
_CAP_COMMON = {"omim": ["description_omim"], "gtop": ["description_name_overview", "description_name_comments"]}
_CAP_2 = {**_RETRIEVAL_COLUMNS_V1, **_CAP_COMMON, "go": ["go_def"], "reactome": ["description"], "disgenet": ["allDescriptions"]}
CAPTION_SUBSETS = {
    1: {**_RETRIEVAL_COLUMNS_V1, **_CAP_COMMON, "disgenet": ["allDescriptions"], "ec": []},
    2: _CAP_2,
    3: {**_CAP_2, "disgenet": ["description_all_collapse"]},
    4: {**_CAP_2, "disgenet": ["description_all_collapse"]},
    5: {**_RETRIEVAL_COLUMNS_V1, **_V5},
}


def _instruction_and_examples(dataset, relation, aaseq_type, task_type, icl_example_number, task_definition, home_dir):
    """(instruction, example text ids, example sequence ids) of `<task_id>.json` under the task directory"""
    import json
    import os

    from .instruct_tune.instruct_constructor import get_prompt, get_prompt_open_def

    home = home_dir or os.environ.get("HOME_DIR")
    if home is None:
        raise RuntimeError("HOME_DIR is not set (root of the checkout that holds procyon/data/instruct_tune/tasks)")
    task_id = construct_task_id(aaseq_type, dataset, relation, task_type)
    with open(os.path.join(home, "procyon", "data", "instruct_tune", "tasks", f"{task_id}.json")) as fh:
        task = json.load(fh)
    kw = dict(task=task, num_examples=icl_example_number, is_special_definition=False, is_ppi=(dataset == "protein"),
              aaseq_type=aaseq_type)
    if task_definition is None:
        instruction, _, _, ex_text, ex_seq = get_prompt(**kw)
    else:
        instruction, _, _, _, ex_text, ex_seq = get_prompt_open_def(**kw)
        instruction = instruction.format(definition=task_definition)
    return instruction, ex_text, ex_seq


def _text_candidates(dataset, columns, data_dir, text_table):
    import os

    if text_table is None:
        import pandas as pd

        root = data_dir or os.environ.get("DATA_DIR")
        if root is None:
            raise RuntimeError("DATA_DIR is not set (root of integrated_data/v1/<dataset>/..._info_filtered_composed.pkl)")
        text_table = pd.read_pickle(os.path.join(root, "integrated_data", "v1", dataset,
                                                 f"{dataset}_info_filtered_composed.pkl"))
    if dataset == "protein":
        return None
    if columns is None:
        raise NotImplementedError("default description columns (ENTITY_DESCRIPTION_NAMES) are not mirrored: "
                                  "set the subset version in DataArgs")
    return text_table[columns[dataset]]


def _sequence_query_dict(instruction, descriptions, seq_ids, device, qa, disease_context_augmentation,
                         functional_descriptions):
    if instruction.endswith("[EXT]"):
        instruction = instruction[:-5]
    seq = torch.LongTensor(seq_ids).to(device)
    if disease_context_augmentation:
        if functional_descriptions is None:
            raise RuntimeError("disease_context_augmentation needs the protein function table "
                               "(integrated_data/v1/protein/uniprot_functional_descriptions.pkl): pass functional_descriptions=")
        instruction = instruction.replace("[CONTEXT]", "[EXT]")
        contexts = functional_descriptions.iloc[seq.cpu().numpy()].tolist()
        mixed = []
        if qa:  # description, then its protein's context
            for c, d in zip(contexts, descriptions):
                mixed += [d, f"Context: {c}"]
        else:   # context first; the query protein's context closes the list
            for i, c in enumerate(contexts):
                mixed.append(f"Context: {c}")
                if i != len(contexts) - 1:
                    mixed.append(descriptions[i])
        descriptions = mixed
    else:
        instruction = instruction.replace("[CONTEXT]", "")
    if qa:
        instruction = instruction.format(answer="null")
    return {
        "data": {"seq": seq, "seq_idx": seq, "text": descriptions, "drug": None},
        "input": {"seq": [list(range(len(seq_ids)))], "text": [list(range(len(descriptions)))], "drug": None},
        "target": {"seq": None, "text": None, "drug": None},
        "instructions": [instruction],
    }


def create_caption_input_simple(input_aaseq_ids, data_args, input_description=None, drug_inputs=None,
                                task_definition=None, instruction_source_dataset=None,
                                instruction_source_relation="all", aaseq_type="protein", task_type="caption",
                                icl_example_number=1, device=None, disease_context_augmentation=False, *,
                                home_dir=None, data_dir=None, text_table=None, functional_descriptions=None) -> Dict:
    """Model-input dict that asks for a description of the sequences `input_aaseq_ids` (indices into the model's
    protein / domain table): the caption task's instruction with its in-context examples, then the query sequences.
    Same arguments and output as the reference (:67-244); keyword-only extensions replace the environment roots."""
    assert drug_inputs is None
    if instruction_source_dataset is None:
        raise NotImplementedError
    dataset = instruction_source_dataset.lower()
    instruction, ex_text, ex_seq = _instruction_and_examples(dataset, instruction_source_relation, aaseq_type, "caption",
                                                             icl_example_number, task_definition, home_dir)
    columns = None
    if task_type == "qa" and data_args.qa_subset_version is not None:
        columns = QA_SUBSETS[data_args.qa_subset_version]
    elif task_type == "caption" and data_args.caption_subset_version is not None:
        columns = CAPTION_SUBSETS[data_args.caption_subset_version]
    cand = _text_candidates(dataset, columns, data_dir, text_table)
    descriptions = cand.iloc[ex_text, 0].tolist()  # (first candidate column, whatever it holds)
    if input_description is not None:
        descriptions = descriptions + [input_description]
    return _sequence_query_dict(instruction, descriptions, list(ex_seq) + list(input_aaseq_ids), device, False,
                                disease_context_augmentation, functional_descriptions)


def create_qa_input_simple(input_aaseq_ids, data_args, input_description, drug_inputs=None, task_definition=None,
                           instruction_source_dataset=None, instruction_source_relation="all", aaseq_type="protein",
                           icl_example_number=1, device=None, disease_context_augmentation=False, *,
                           home_dir=None, data_dir=None, text_table=None, functional_descriptions=None) -> Dict:
    """Model-input dict of a yes/no question "is `input_description` true of the sequences `input_aaseq_ids`?" (:247-
    420): the QA task's instruction with positive and negative in-context examples, answer slot filled with "null"."""
    assert drug_inputs is None
    if instruction_source_dataset is None:
        raise NotImplementedError
    dataset = instruction_source_dataset.lower()
    instruction, ex_text, ex_seq = _instruction_and_examples(dataset, instruction_source_relation, aaseq_type, "qa",
                                                             icl_example_number, task_definition, home_dir)
    cand = _text_candidates(dataset, QA_SUBSETS[data_args.qa_subset_version], data_dir, text_table)
    descriptions = [_first_text(cand.iloc[i, :]) for i in ex_text]
    if input_description is not None:
        descriptions = descriptions + [input_description]
    return _sequence_query_dict(instruction, descriptions, list(ex_seq) + list(input_aaseq_ids), device, True,
                                disease_context_augmentation, functional_descriptions)

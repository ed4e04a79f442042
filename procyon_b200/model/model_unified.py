"""UnifiedProCyon on the B200 kernels — host-side mirror of procyon/model/model_unified.py.

Same public surface as the reference class (`forward`, `generate`, `forward_sequences`, `from_pretrained`,
`get_checkpoint_configs`, `.config .tokenizer .yes_token .no_token .answer_idx .text_encoder
.protein_seq_encoder`), same input-dict format and same `state_dict` keys, but every tensor op on the
path — ESM2 encode, pooling, projector MLPs, embedding + soft-token splice, Llama prefill / KV-cache decode, token
selection, LM loss, InfoNCE — is a call into libprocyon_b200.so. Host code only does what the reference does on
the host: string templating, tokenisation, index bookkeeping.

Deviations from the reference, all deliberate and listed in DESIGN.md:
  * `_generate_sampling` is called with its intended signature (the reference call site drops `attn_masks`,
    model_unified.py:998-1005, and cannot run);
  * left-pad positions stay masked in decode steps (reference: visible after step 0, model_unified.py:769);
  * the protein encoder's LM-head logits are not computed on the pooled path (discarded by the reference, :391);
  * the `dataset_id` "gather" bug (:657-660) is not reproduced: ids are really all-gathered.
"""
from __future__ import annotations

import os
from itertools import chain
from typing import List, Optional

import torch
from torch import nn

from .. import _lib
from .._lib import c_i64, c_int, check, ptr, stream_ptr
from ..training.train_utils import barrier, unwrap_model
from ..training.training_args_IT import ModelArgs
from .contrastive import InfoNCEInBatch
from .esm import ESM_PLM
from .generation import generate_beam_search, generate_greedy
from .model_utils import compute_conflict_matrix, create_mlp, left_pad_tensors
from .pmc_llama import SELECT_GREEDY, LlamaConfig, LlamaPostTokenization

from ..training.trainIT import DATASET_ID as _DATASET_ID  # noqa: E402

DATASET_ID_PROTEIN = _DATASET_ID["protein"]  # procyon/data/constants.py:666-671 (= 4)
# procyon/model/model_unified.py:33 (`f'{DATA_DIR}/model_weights/'`); None when DATA_DIR is not configured
DEFAULT_PRETRAINED_WEIGHTS_DIR = (os.path.join(os.environ["DATA_DIR"], "model_weights") + os.sep
                                  if os.environ.get("DATA_DIR") else None)


def mask_before(full_labels, answer_idx, before_last_answer=False):
    """procyon/model/model_unified.py:39-60 (vectorised: no per-row `.item()`)."""
    is_ans = full_labels == answer_idx
    if not before_last_answer:
        if torch.any(is_ans.sum(dim=1) > 1):
            raise ValueError("More than one {} token detected in an input".format(answer_idx))
    if not bool(is_ans.any(dim=1).all()):
        raise ValueError("every input needs an [ANSWER] token")  # reference: .max() of an empty tensor
    S = full_labels.shape[1]
    ar = torch.arange(S, device=full_labels.device)
    last = torch.where(is_ans, ar[None, :], torch.full_like(ar, -1)[None, :]).max(dim=1).values
    return last[:, None] >= ar[None, :]


def multi_replace_tokens(a, b, replace_token, eval=False):
    """procyon/model/model_unified.py:83-108: substitute successive `replace_token` ids in list a by lists b[i]."""
    occ = [i for i, t in enumerate(a) if t == replace_token]
    if len(occ) != len(b):
        raise ValueError("Number of occurrences of replace_token does not match the length of b")
    if len(occ) == 0:
        return a
    result = a[: occ[0]]
    for i, o in enumerate(occ):
        if not (i == len(occ) - 1 and eval):
            result = result + list(b[i])
        if i == len(occ) - 1:
            result = result + a[o + 1 :]
        else:
            result = result + a[o + 1 : occ[i + 1]]
    return result


class UnifiedProCyon(nn.Module):
    def __init__(self, config: ModelArgs, pretrained_weights_dir=None, for_pretraining=True, *, tokenizer=None,
                 llama_config: Optional[LlamaConfig] = None, esm_custom_config=None, device=None, dtype=None,
                 protein_embeddings=None, domain_embeddings=None, peptide_embeddings=None,
                 protein_struct_embeddings=None, drug_embeddings=None):
        super().__init__()
        self.config = config
        self.pretrained_weights_dir = pretrained_weights_dir
        self.causal_qa = getattr(config, "causal_qa", True)
        self.train_qa_full_lm = getattr(config, "train_qa_full_lm", True)
        assert self.causal_qa, "Non-causal QA currently not working, causal_qa must be set to true"
        if not config.text_encoder_fname.lower().startswith("llama"):
            raise ValueError(f"Unrecognized text encoder: {config.text_encoder_fname}")
        if getattr(config, "freeze_text_encoder", None) in ("lora", "qlora"):
            raise NotImplementedError("LoRA text encoders are out of scope")

        self.text_encoder = LlamaPostTokenization(
            model_path=os.path.join(pretrained_weights_dir or "", config.text_encoder_fname),
            model_splitting=False, attention_type=config.attention_type, for_pretraining=for_pretraining,
            config=llama_config, device=device, dtype=dtype)
        if getattr(config, "text_encoder_debug", False):
            self.text_encoder.model.model.layers = self.text_encoder.model.model.layers[:2]
            self.text_encoder.model.config.num_hidden_layers = 2
        self._init_tokenizer(tokenizer)
        self.text_encoder.model.resize_token_embeddings(len(self.tokenizer) - 1)
        self.text_embed_dim = self.text_encoder.model.config.hidden_size
        self.input_embeddings = self.text_encoder.model.get_input_embeddings()

        # ---- protein side ----
        if config.use_aaseq_embeddings:
            self.protein_seq_encoder = None
            if protein_embeddings is None and getattr(config, "protein_seq_embeddings_path", None) \
                    and os.environ.get("DATA_DIR"):
                # the reference's own files (model_unified.py:189-211): embeddings in id-map order -> table order
                from ..data.data_utils import load_aaseq_embeddings

                # the id maps are ALWAYS taken from the current DATA_DIR, whatever the checkpoint's config says
                # (reference :197-198 overwrites both fields unconditionally)
                node = os.path.join(os.environ["DATA_DIR"], "generated_data/node_embeddings")
                config.protein_embeddings_idmap_path = os.path.join(node, "protein/protein_esm2-3b_mean.pkl")
                config.domain_embeddings_idmap_path = os.path.join(node, "domain/domain_esm2-3b_mean.pkl")
                protein_embeddings = load_aaseq_embeddings(config.protein_seq_embeddings_path,
                                                           config.protein_embeddings_idmap_path, "protein")
                if domain_embeddings is None and getattr(config, "domain_embeddings_path", None):
                    domain_embeddings = load_aaseq_embeddings(config.domain_embeddings_path,
                                                              config.domain_embeddings_idmap_path, "domain")
                if peptide_embeddings is None and getattr(config, "peptide_embeddings_path", None):
                    assert config.peptide_embeddings_idmap_path is not None
                    peptide_embeddings = load_aaseq_embeddings(config.peptide_embeddings_path,
                                                               config.peptide_embeddings_idmap_path, "peptide")
            if protein_embeddings is None:
                raise ValueError("use_aaseq_embeddings=True needs the pre-computed embedding tables "
                                 "(protein_embeddings=..., domain_embeddings=..., or config.*_embeddings_path with "
                                 "DATA_DIR set)")
            self.protein_seq_embeddings = nn.Embedding.from_pretrained(protein_embeddings, freeze=True)
            if domain_embeddings is not None:
                self.domain_embeddings = nn.Embedding.from_pretrained(domain_embeddings, freeze=True)
            if peptide_embeddings is not None:
                self.peptide_embeddings = nn.Embedding.from_pretrained(peptide_embeddings, freeze=True)
            self.protein_embed_dim = protein_embeddings.shape[1]
        else:
            self.protein_seq_encoder = ESM_PLM(
                pretrained_weights_dir=pretrained_weights_dir,
                num_params=config.protein_encoder_num_params,
                pooling_method=config.protein_pooling_opt,
                padding_idx=1, eos_idx=2,  # fair-esm Alphabet "ESM-1b"
                long_protein_strategy=config.long_protein_strategy,
                max_protein_len=config.max_protein_len,
                max_batch_forward_pass=config.protein_enc_batch_limit,
                protein_pooling_correction_option=config.protein_pooling_correction_option,
                custom_config=esm_custom_config,
            )
            self.protein_embed_dim = self.protein_seq_encoder.embedding_size

        d_txt = self.input_embeddings.weight.shape[-1]
        self.token_projectors = nn.ModuleDict({
            "aaseq": create_mlp(config.num_layers_token_projector, self.protein_embed_dim, d_txt,
                                config.hidden_size_token_projector)})
        if config.use_protein_struct:
            if protein_struct_embeddings is None and getattr(config, "protein_struct_embeddings_path", None):
                from ..data.data_utils import load_protein_struct_embeddings  # reference :271

                protein_struct_embeddings = load_protein_struct_embeddings(config.protein_struct_embeddings_path)
            if protein_struct_embeddings is None:
                raise ValueError("use_protein_struct=True needs protein_struct_embeddings=... or "
                                 "config.protein_struct_embeddings_path")
            self.protein_struct_embeddings = nn.Embedding.from_pretrained(protein_struct_embeddings, freeze=True)
            self.protein_struct_embed_dim = protein_struct_embeddings.shape[1]
            self.token_projectors.update({"prot_structure": create_mlp(
                config.num_layers_token_projector, self.protein_struct_embed_dim, d_txt,
                config.hidden_size_token_projector)})
        else:
            self.protein_struct_embeddings = None
        if config.use_drug_embeddings:
            if drug_embeddings is None and getattr(config, "drug_struct_embeddings_path", None):
                from ..data.data_utils import load_drug_structure_embeddings  # reference :286

                drug_embeddings = load_drug_structure_embeddings(config.drug_struct_embeddings_path)
            if drug_embeddings is None:
                raise ValueError("use_drug_embeddings=True needs drug_embeddings=... or "
                                 "config.drug_struct_embeddings_path")
            self.drug_structure_embeddings = nn.Embedding.from_pretrained(drug_embeddings, freeze=True)
            self.drug_embed_dim = drug_embeddings.shape[1]
            self.token_projectors.update({"drug": create_mlp(
                config.num_layers_token_projector, self.drug_embed_dim, d_txt, config.hidden_size_token_projector)})
        else:
            self.drug_structure_embeddings = None

        self.aaseq_shared_projector = create_mlp(config.num_layers_shared_projector, self.protein_embed_dim,
                                                 self.protein_embed_dim, config.hidden_size_shared_projector)
        self.aaseq_lm_projector = create_mlp(config.num_layers_lm_projector, self.text_embed_dim,
                                             self.protein_embed_dim, config.hidden_size_lm_projector)
        assert config.negative_sampling_strategy_retrieval == "in_batch"
        if config.cl_method.lower() != "infonce":
            raise NotImplementedError("only cl_method='infonce' is on the hot path")
        self.contrastive_head = InfoNCEInBatch(self.protein_embed_dim, use_projection=config.use_projection_cl,
                                               all_gather_version=config.contrastive_global)
        if "llama-3" in config.text_encoder_fname.lower():
            self.yes_token = self.tokenizer.encode(" yes", add_special_tokens=False)[0]
            self.no_token = self.tokenizer.encode(" no", add_special_tokens=False)[0]
        else:
            self.yes_token = self.tokenizer.encode("yes", add_special_tokens=False)[0]
            self.no_token = self.tokenizer.encode("no", add_special_tokens=False)[0]
        self.context_crop_sampling = config.context_crop_sampling
        self.struct_dropout_prob = config.protein_struct_dropout
        if device is not None or dtype is not None:
            self.to(device=device, dtype=dtype)

    # --------------------------------------------------------------------------------------------------------
    def _init_tokenizer(self, tokenizer=None):
        """model_unified.py:1088-1133: same special tokens, same order ([EXT] last)."""
        if tokenizer is None:
            import transformers

            if "llama-3" in self.config.text_encoder_fname.lower():
                tokenizer = transformers.AutoTokenizer.from_pretrained(os.getenv("LLAMA3_PATH"))
                tokenizer.padding_side = "right"
            else:
                tokenizer = transformers.LlamaTokenizer.from_pretrained(
                    os.path.join(self.pretrained_weights_dir, self.config.text_encoder_fname))
        self.tokenizer = tk = tokenizer
        self.use_llama_tokenizer = True

        def first_id(s):
            return tk(s, add_special_tokens=False).input_ids[0]

        if tk.sep_token is None:
            tk.add_tokens("[CLS]")
            tk.sep_token = "[CLS]"
            tk.sep_token_id = first_id(tk.sep_token)
        if tk.pad_token is None:
            tk.add_tokens("[PAD]")
            tk.pad_token = "[PAD]"
            tk.pad_token_id = first_id(tk.pad_token)
        tk.add_tokens("<|protein|>")
        self.prot_replacement_idx = first_id("<|protein|>")
        tk.add_tokens("[PROT]")
        self.prot_retrieval_idx = first_id("[PROT]")
        tk.add_tokens("[ANSWER]")
        self.answer_idx = first_id("[ANSWER]")
        tk.add_tokens("<|struct|>")
        self.struct_idx = first_id("<|struct|>")
        tk.add_tokens("<|drug|>")
        self.drug_idx = first_id("<|drug|>")
        tk.add_tokens("[EXT]")  # must come last: its row is dropped from the embedding table (:166)
        self.ext_idx = first_id("[EXT]")

    # --------------------------------------------------------------------------------------------------------
    def _aaseq_table(self, aaseq_type):
        return {"protein": "protein_seq_embeddings", "domain": "domain_embeddings",
                "peptide": "peptide_embeddings"}[aaseq_type]

    def _preprocessing(self, inputs, aaseq_type="protein", exclude_protein_structure=False, crop_off=False,
                       no_pad=False, retrieval=False, left_pad=False):
        """model_unified.py:352-481."""
        dev = self.input_embeddings.weight.device
        aaseq_token_embeddings = aaseq_ret_embeddings = None
        if inputs["data"]["seq"] is not None:
            if self.config.use_aaseq_embeddings:
                emb = getattr(self, self._aaseq_table(aaseq_type))(inputs["data"]["seq"].to(dev))
            else:
                emb, _ = self.protein_seq_encoder(inputs["data"]["seq"].to(dev), aggregate=True)
            aaseq_token_embeddings = aaseq_ret_embeddings = emb

        protein_soft_tokens = None
        if inputs["input"]["seq"] is not None:
            full_index = list(chain.from_iterable(inputs["input"]["seq"]))
            pz_inputs = aaseq_token_embeddings[full_index]
            protein_soft_tokens = self.token_projectors["aaseq"](pz_inputs)

        drug_soft_tokens = None
        if self.config.use_drug_embeddings and (inputs["data"].get("drug") is not None):
            full_index = list(chain.from_iterable(inputs["input"]["drug"]))
            drug_z = self.drug_structure_embeddings(inputs["data"]["drug"].to(dev))[full_index]
            drug_soft_tokens = self.token_projectors["drug"](drug_z)

        text_inputs = [[inputs["data"]["text"][i] for i in inp_list] for inp_list in inputs["input"]["text"]]
        instruction_list = list(inputs["instructions"])
        # Structure soft tokens (reference :409-460): every instruction that survives the structure dropout gets one
        # <|struct|> placeholder after each <|protein|>, filled with the projected GearNet embedding of that protein.
        # One entry per instruction: a [k, d] tensor, or [] for an instruction whose structure was dropped.
        protein_struct_tokens = []
        if (not exclude_protein_structure) and self.config.use_protein_struct and inputs["input"]["seq"]:
            n_instr = len(instruction_list)
            keep = torch.bernoulli(torch.full((n_instr,), 1 - self.struct_dropout_prob)).bool().tolist()
            table_rows = []  # per kept instruction: rows of the structure table, one per input protein
            for i in range(n_instr):
                if keep[i]:
                    instruction_list[i] = instruction_list[i].replace("<|protein|>", "<|protein|> <|struct|>")
                    table_rows.append(torch.stack([torch.as_tensor(inputs["data"]["seq_idx"][j]).reshape(())
                                                   for j in inputs["input"]["seq"][i]]))
            if table_rows:
                table_rows = torch.stack(table_rows, dim=0)  # [kept, proteins per instruction]
                distinct, where = table_rows.unique(return_inverse=True)  # project every distinct protein once
                if aaseq_type == "protein":
                    z = self.protein_struct_embeddings(distinct.to(dev))
                else:  # domains / peptides have no structure table: a zero embedding goes through the projector
                    z = torch.zeros((distinct.shape[0], self.protein_struct_embed_dim), device=dev,
                                    dtype=self.protein_struct_embeddings.weight.dtype)
                per_instruction = iter(self.token_projectors["prot_structure"](z)[where.to(dev)])
                protein_struct_tokens = [next(per_instruction) if k else [] for k in keep]

        input_ids, attn_masks = self._prepare_text_inputs_and_tokenize(
            instruction_list, text_inputs, crop_off=crop_off, retrieval=retrieval, no_pad=no_pad, left_pad=left_pad)
        input_ids = input_ids.to(dev)
        attn_masks = attn_masks.to(dev)
        input_embeds, ret_output_indices = self._prepare_input_embeddings(
            input_ids, protein_soft_tokens=protein_soft_tokens, protein_struct_tokens=protein_struct_tokens,
            drug_soft_tokens=drug_soft_tokens)
        return input_embeds, input_ids, attn_masks, ret_output_indices, aaseq_token_embeddings, aaseq_ret_embeddings

    def _prepare_input_embeddings(self, input_ids, protein_soft_tokens=None, protein_struct_tokens=[],
                                  drug_soft_tokens=None):
        """model_unified.py:1135-1175: one fused gather + scatter launch (pcy_embed_splice)."""
        lib = _lib.load()
        dev = input_ids.device
        B, S = input_ids.shape
        table = self.input_embeddings.weight
        if table.dtype != torch.bfloat16 or not table.is_cuda:
            raise _lib.ProcyonB200Error("the model must be on CUDA in bfloat16 (`model.bfloat16().cuda()`)")
        d = table.shape[1]
        soft_parts, index = [], torch.full((B, S), -1, device=dev, dtype=torch.int64)
        offset = 0

        def place(mask, soft, what):
            nonlocal offset
            n = int(mask.sum())
            assert n == soft.shape[0], f"{what}: expected {n} soft tokens, got {soft.shape[0]}"
            order = torch.cumsum(mask.flatten().to(torch.int64), 0) - 1 + offset
            index.view(-1)[mask.flatten()] = order[mask.flatten()]
            soft_parts.append(soft.to(torch.bfloat16))
            offset += n

        if protein_soft_tokens is not None:
            place(input_ids == self.prot_replacement_idx, protein_soft_tokens, "protein")
        if len(protein_struct_tokens) > 0:
            smask = input_ids == self.struct_idx
            rows = [t for t in protein_struct_tokens if not isinstance(t, list)]
            for i in range(B):
                cnt = int(smask[i].sum())
                if cnt > 0:
                    assert cnt == protein_struct_tokens[i].shape[0], \
                        f"expected: {cnt} got: {protein_struct_tokens[i].shape[0]}"
            if rows:
                place(smask, torch.cat([r.reshape(-1, d) for r in rows], 0), "structure")
        if drug_soft_tokens is not None:
            place(input_ids == self.drug_idx, drug_soft_tokens, "drug")

        ids32 = input_ids.to(torch.int32).contiguous()
        out = torch.empty((B, S, d), device=dev, dtype=torch.bfloat16)
        soft = torch.cat(soft_parts, 0).contiguous() if soft_parts else None
        idx32 = index.to(torch.int32).contiguous() if soft_parts else None
        check(lib.pcy_embed_splice(ptr(ids32), ptr(table), ptr(soft), ptr(idx32), ptr(out), c_i64(B * S), c_int(d),
                                   stream_ptr(dev)), "pcy_embed_splice")
        ret = input_ids == self.prot_retrieval_idx
        if self.config.roll_num != 0:
            ret = ret.roll(self.config.roll_num, 1)
        return out, ret

    def _prepare_text_inputs_and_tokenize(self, instructions: List[str], text_input_list: List[List[str]],
                                          crop_off=False, retrieval=False, no_pad=False, left_pad=False):
        """model_unified.py:1177-1293 (host string/token work, unchanged semantics)."""
        import random

        tk = self.tokenizer
        assert all([t[-1] != tk.sep_token for t in instructions])
        instruction_tokens = tk(instructions, padding=False, truncation=True, add_special_tokens=True,
                                max_length=self.config.max_text_len)["input_ids"]
        max_len = max(len(l) for l in instruction_tokens)
        joint_tokens, attention_masks = [], []
        for i, text_input in enumerate(text_input_list):
            n_in = len(text_input)
            if n_in != 0:
                text_input = [t if isinstance(t, str) else "null" for t in text_input]
                toks = tk(text_input, padding=False, truncation=False, add_special_tokens=False)["input_ids"]
                max_len_for_sample = (self.config.max_text_len - max_len) // n_in
                for j in range(len(toks)):
                    drug_add = None
                    if self.drug_idx in toks[j]:
                        where_drug = toks[j].index(self.drug_idx) - 3
                        drug_add = toks[j][(where_drug - 3):]
                        toks[j] = toks[j][: (where_drug - 3)]
                    if self.training and self.context_crop_sampling and (not crop_off):
                        top_end = len(toks[j]) - max_len_for_sample
                        start_i = 0 if top_end <= 0 else random.randint(0, top_end)
                    else:
                        start_i = 0
                    end_i = start_i + max_len_for_sample
                    if drug_add is not None:
                        end_i -= len(drug_add)
                    toks[j] = toks[j][start_i:end_i]
                    if drug_add is not None:
                        toks[j] = toks[j] + drug_add
            else:
                toks = []
            L = multi_replace_tokens(instruction_tokens[i], toks, self.ext_idx, eval=False)
            if no_pad:
                L = torch.tensor(L)
            else:
                L = torch.tensor(L + [tk.eos_token_id] + [tk.pad_token_id] * max(self.config.max_text_len - len(L) - 1, 0))
            joint_tokens.append(L)
            attention_masks.append((L != tk.pad_token_id).int())
            assert not torch.any(L == self.ext_idx), "ERROR [EXT] found in input"
        if left_pad:
            return left_pad_tensors(joint_tokens, pad_value=tk.pad_token_id)
        return torch.stack(joint_tokens, dim=0), torch.stack(attention_masks, dim=0)

    # --------------------------------------------------------------------------------------------------------
    def forward(self, inputs, return_mlm=False, retrieval=False, get_full_labels=False, aaseq_type="protein",
                exclude_protein_structure=False, crop_off=False, output_attentions=False):
        """model_unified.py:483-699."""
        if return_mlm:  # masked-LM logits of the protein encoder only; the text encoder is not touched (:505-509)
            _, logits = self.protein_seq_encoder(inputs["data"]["seq"].to(self.input_embeddings.weight.device),
                                                 aggregate=False)
            return {"mlm": logits}
        ignore_struct = (inputs["target"]["seq"] is not None) and self.training and not exclude_protein_structure
        (input_embeds, input_ids, attn_masks, ret_output_indices, protein_token_embeddings,
         protein_ret_embeddings) = self._preprocessing(inputs, aaseq_type=aaseq_type, crop_off=crop_off,
                                                       retrieval=retrieval, exclude_protein_structure=ignore_struct)
        full_labels = None
        if not retrieval:
            full_labels = input_ids.clone()
            all_masks = ((full_labels == self.tokenizer.pad_token_id) | (full_labels == self.prot_replacement_idx)
                         | (full_labels == self.prot_retrieval_idx) | (full_labels == self.drug_idx)
                         | (full_labels == self.struct_idx))
            if self.use_llama_tokenizer:
                all_masks[:, -1] = True
            if not self.train_qa_full_lm:
                all_masks |= mask_before(full_labels, self.answer_idx, before_last_answer=True)
            full_labels = torch.where(all_masks, -100, full_labels)

        sum_all = retrieval and self.config.ret_token_access == "all"
        outputs = self.text_encoder(input_embeds=input_embeds, attn_masks=attn_masks, full_labels=full_labels,
                                    output_attentions=output_attentions,
                                    sum_hidden_rows=ret_output_indices if sum_all else None)
        out_dict = {"outputs": outputs, "text_toks": input_ids,
                    "full_labels": full_labels if get_full_labels else None,
                    "contrastive_out": None, "contrastive_loss": None}
        if retrieval:
            contrastive_out = {"positive": {}, "negative": {}}
            if self.config.ret_token_access == "last":
                extracted_ret = outputs.hidden_states[-1][ret_output_indices]
            elif self.config.ret_token_access == "all":
                # sum over all L+1 hidden states (FROMAGe-style, model_unified.py:560-563), accumulated in fp32 at the
                # [PROT] rows only while the layers run; rounded once like torch's bf16 sum
                extracted_ret = outputs.hidden_sum.to(outputs.hidden_states[-1].dtype)
            else:
                raise NotImplementedError("Invalid option {} for ret_token_access".format(self.config.ret_token_access))
            shared_lm_output = self.aaseq_lm_projector(extracted_ret)
            if inputs["target"]["text"] is None:
                contrastive_out["positive"]["text"] = shared_lm_output
            else:
                contrastive_out["positive"]["text"] = shared_lm_output[inputs["target"]["text"]["positive"]]
                if inputs["target"]["text"]["negative"] is not None:
                    raise NotImplementedError
            if inputs["target"]["seq"] is not None:
                shared_plm_output = self.aaseq_shared_projector(protein_ret_embeddings)
                contrastive_out["positive"]["sequence"] = shared_plm_output[inputs["target"]["seq"]["positive"]]
                if inputs["target"]["seq"]["negative"] is not None:
                    raise NotImplementedError
                conflict_mat = None
                if self.config.filter_negatives_by_id_contrastive and self.training:
                    conflict_mat = self._conflict_matrix(inputs, aaseq_type, shared_plm_output.device)
                if self.training:
                    out_dict["contrastive_loss"] = self.contrastive_head(contrastive_out, negatives_mask=conflict_mat)
                else:
                    out_dict["contrastive_loss"] = -999.0
            out_dict["contrastive_out"] = contrastive_out
        return out_dict

    def _conflict_matrix(self, inputs, aaseq_type, dev):
        """model_unified.py:596-684 (host-side id bookkeeping; ids are int64 vectors of the batch size)."""
        if inputs["target"]["text"] is not None:
            raise NotImplementedError
        if any(inputs["input"]["text"]):
            local = [row[-1] for row in inputs["input"]["text"]]
            text_ids = torch.LongTensor([inputs["data"]["text_idx"][i] for i in local]).to(dev)
        else:
            local = [row[-1] for row in inputs["input"]["seq"]]
            text_ids = torch.LongTensor([(-1 - int(inputs["data"]["seq_idx"][i])) for i in local]).to(dev)
        prot_ids = torch.LongTensor([int(inputs["data"]["seq_idx"][i]) for i in inputs["target"]["seq"]["positive"]]).to(dev)
        dset_ids = inputs["dataset_id"].to(dev) if "dataset_id" in inputs else None
        if self.config.contrastive_global and torch.distributed.is_available() and torch.distributed.is_initialized():
            W = torch.distributed.get_world_size()

            def gather(v):
                buf = [torch.empty_like(v) for _ in range(W)]
                torch.distributed.all_gather(buf, v)
                return torch.cat(buf, dim=0)

            barrier()
            text_ids, prot_ids = gather(text_ids), gather(prot_ids)
            if dset_ids is not None:
                dset_ids = gather(dset_ids)
        code = {"protein": 0, "domain": 1, "peptide": 2}[aaseq_type]
        ind = torch.full_like(prot_ids, code)
        aaseq_overlap = ind[None, :] == ind[:, None]
        text_conflict = compute_conflict_matrix(text_ids, prot_ids)
        prot_conflict = aaseq_overlap & compute_conflict_matrix(prot_ids, text_ids)
        if dset_ids is not None:
            dset_overlap = dset_ids[None, :] == dset_ids[:, None]
            ppi = dset_ids == DATASET_ID_PROTEIN
            ppi_matrix = ppi[None, :] == ppi[:, None]
            text_conflict = dset_overlap & text_conflict
            text_conflict[ppi_matrix] = False
        return ~(text_conflict | prot_conflict)

    # --------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _generate_beam_search(self, input_embeds, attn_mask, max_len=64, beam_size=5, beam_group_size=5,
                              diversity_penalty=0.8, return_logits=True):
        """model_unified.py:701-842 — runs on the device (procyon_b200.model.generation); batches of more than 16
        beam rows step through several lock-step sessions that share the reference's whole-batch stop condition."""
        any_pad = attn_mask is not None and bool((attn_mask == 0).any())
        return generate_beam_search(
            self.text_encoder, input_embeds, attn_mask if any_pad else None, max_len=max_len, beam_size=beam_size,
            beam_group_size=beam_group_size, diversity_penalty=diversity_penalty,
            eos_token_id=self.tokenizer.eos_token_id, return_logits=return_logits)

    def _get_nucleus_mask(self, probs, nucleus_prob):
        """model_unified.py:844-858."""
        remove_prob = 1 - nucleus_prob
        sorted_vals, indices = probs.sort(dim=-1, descending=False)
        keep = (sorted_vals.cumsum(dim=-1) >= remove_prob).nonzero(as_tuple=True)
        mask = torch.zeros_like(probs)
        mask[keep[0], indices[keep]] = 1
        return mask

    @torch.no_grad()
    def _generate_sampling(self, input_embeds, attn_masks, max_len=64, num_text_per_instance=1, temperature=1.0,
                           greedy=False, nucleus_prob=None, return_logits=True):
        """model_unified.py:860-921 with the intended signature. Greedy is fully on-device; temperature / nucleus
        sampling draw with torch.multinomial from the device logits (the model step is still one graph-free call
        per token)."""
        assert nucleus_prob is None or (0 < nucleus_prob < 1)
        n = input_embeds.shape[0]
        any_pad = attn_masks is not None and bool((attn_masks == 0).any())
        am = attn_masks if any_pad else None
        out_list, lp_list, logit_list = [], [], []
        for _ in range(num_text_per_instance):
            if greedy:
                outs, lps, lgs = [], [], []
                for i0 in range(0, n, 16):
                    sl = slice(i0, min(n, i0 + 16))
                    o, lp, lg = generate_greedy(self.text_encoder, input_embeds[sl], am[sl] if am is not None else None,
                                                max_len=max_len, return_logits=return_logits)
                    outs.append(o), lps.append(lp), lgs.append(lg)
                out = torch.cat(outs, 0)
                total = torch.cat(lps, 0)
                logits = torch.cat(lgs, 0) if return_logits else None
            else:
                out, total, logits = self._sample_loop(input_embeds, am, max_len, temperature, nucleus_prob,
                                                       return_logits)
            out_list.append(out)
            lp_list.append(total)
            logit_list.append(logits)
        out_tokens = torch.stack(out_list, dim=1).cpu()
        out_logits = torch.stack(logit_list, dim=1) if return_logits else None
        log_probs = torch.stack(lp_list).T
        return out_tokens, log_probs, out_logits

    def _sample_loop(self, input_embeds, attn_masks, max_len, temperature, nucleus_prob, return_logits):
        """Temperature / nucleus sampling (model_unified.py:885-911): logits stay on the device, torch.multinomial draws
        there; any batch size (rows beyond 16 go through further decode sessions)."""
        from .pmc_llama import SessionGroup

        te = self.text_encoder
        n, S, _ = input_embeds.shape
        dev = input_embeds.device
        sel = torch.arange(n, device=dev, dtype=torch.int32) * S + (S - 1)
        kv, _, logits, valid = te.prefill(input_embeds, attn_masks, want_cache=True, want_hidden=False, sel_rows=sel)
        sess = te.sessions_from_prefill(kv, valid if attn_masks is not None else None, logits, max_len)
        if not isinstance(sess, SessionGroup):
            sess = SessionGroup([sess])
        del kv
        total = torch.zeros(n, device=dev)
        all_logits, toks = [], []
        cur = sess.logits()
        for i in range(max_len):
            if return_logits:
                all_logits.append(cur)
            lp = torch.log_softmax(cur, dim=-1)
            if nucleus_prob is not None:
                probs = cur.softmax(dim=-1)
                probs = probs * self._get_nucleus_mask(probs, nucleus_prob)
            else:
                probs = (cur / temperature).softmax(dim=-1)
            nxt = torch.multinomial(probs, 1)
            total += lp[torch.arange(n, device=dev), nxt.squeeze(-1)]
            toks.append(nxt)
            if i + 1 < max_len:
                cur = sess.step(nxt.squeeze(-1))
        out = torch.cat(toks, dim=-1).cpu()
        return out, total.cpu(), (torch.stack(all_logits, 1).cpu() if return_logits else None)

    @torch.no_grad()
    def generate(self, inputs, max_len=64, aaseq_type="protein", method="sampling", temperature=1.0, greedy=False,
                 num_text_per_instance=1, return_all_internals=False, beam_size=5, beam_group_size=5,
                 diversity_penalty=0.8, exclude_protein_structure=False, nucleus_prob=0.9, truncate_on_eos=True,
                 return_logits=True):
        """model_unified.py:923-1027. `return_logits=False` skips the (n, beams, steps, V) logits history."""
        assert method in ["sampling", "temperature", "greedy", "beam", "nucleus"]
        if method == "beam":
            num_text_per_instance = beam_size
        elif method == "greedy":
            greedy = True
        elif method in ["sampling", "nucleus"]:
            temperature = 1
        if temperature < 1e-8:
            greedy = True
        self.text_encoder.eval()
        if self.protein_seq_encoder is not None:
            self.protein_seq_encoder.eval()
        (input_embeds, input_ids, attn_masks, ret_output_indices, _, _) = self._preprocessing(
            inputs, aaseq_type=aaseq_type, crop_off=True, no_pad=True,
            exclude_protein_structure=exclude_protein_structure, left_pad=True)
        whole_instructions = self.tokenizer.batch_decode(input_ids)
        gt_text = None
        if inputs["target"]["text"] is not None:
            gt_text = [inputs["data"]["text"][i] for i in inputs["target"]["text"]]
        batch_size = input_embeds.shape[0]
        if method == "beam":
            output_tokens, log_probs, output_logits = self._generate_beam_search(
                input_embeds, attn_masks, max_len=max_len, beam_size=beam_size, diversity_penalty=diversity_penalty,
                beam_group_size=beam_group_size, return_logits=return_logits)
        else:
            output_tokens, log_probs, output_logits = self._generate_sampling(
                input_embeds, attn_masks, max_len, num_text_per_instance, temperature, greedy,
                nucleus_prob if method == "nucleus" else None, return_logits=return_logits)
        flattened = torch.flatten(output_tokens, start_dim=0, end_dim=1)
        texts = self.tokenizer.batch_decode(flattened)
        if truncate_on_eos:
            texts = [x.split(self.tokenizer.eos_token)[0].strip() for x in texts]
        texts = [texts[i * num_text_per_instance:(i + 1) * num_text_per_instance] for i in range(batch_size)]
        if return_all_internals:
            return {
                "out_tokens": output_tokens, "out_logits": output_logits, "out_log_probs": log_probs, "text": texts,
                "input_instructions": whole_instructions, "ground_truth_text": gt_text,
                "text_references": inputs["reference_indices"]["target"]["text"],
                "seq_references": [inputs["reference_indices"]["input"]["seq"][j][-1]
                                   for j, _ in enumerate(inputs["input"]["seq"])],
            }
        return output_tokens, log_probs, output_logits, texts

    def forward_sequences(self, seq_input, get_soft_tokens=False, aaseq_type="protein"):
        """model_unified.py:1029-1086."""
        if isinstance(seq_input, dict):
            seq_input = seq_input["data"]
        dev = self.input_embeddings.weight.device
        if self.config.use_aaseq_embeddings:
            protein_embeddings = getattr(self, self._aaseq_table(aaseq_type))(seq_input.to(dev))
        else:
            protein_embeddings, _ = self.protein_seq_encoder(seq_input.to(dev), aggregate=True)
        output = {"original": protein_embeddings, "shared": self.aaseq_shared_projector(protein_embeddings),
                  "token": None}
        if get_soft_tokens:
            output["token"] = self.token_projectors["aaseq"](protein_embeddings)
        return output

    # --------------------------------------------------------------------------------------------------------
    @staticmethod
    def from_pretrained(*, pretrained_weights_dir=None, checkpoint_dir=None, model: nn.Module = None, config_only=False,
                        config=None, state_dict_relative_path: str = "txllm_model_ckpt.pt", strict_load=False,
                        load_plm_directly=False, protein_pooling_correction_option=False, **model_kwargs):
        """model_unified.py:1295-1394: same checkpoint layout (model_args.pt, data_args.pt, txllm_model_ckpt.pt)."""
        from .. import compat

        compat.install()  # lets pickles of procyon.training.training_args_IT.* resolve
        config_checkpoint = torch.load(os.path.join(checkpoint_dir, "model_args.pt"), weights_only=False)
        if config is None:
            config = config_checkpoint
        if config_only:
            return None, config
        config.n_model_pieces = 1
        config.model_splitting = False
        if load_plm_directly and config.use_aaseq_embeddings:
            name = os.path.basename(config.protein_seq_embeddings_path).split(".")[0].split("_")
            _, nparams_name, pooling_method = name
            assert pooling_method in ["max", "mean"]
            nparams = {"esm2-3b": "3b", "esm-650m": "650m"}.get(nparams_name)
            if nparams is None:
                raise NotImplementedError("Invalid number of parameters")
            config.use_aaseq_embeddings = False
            config.freeze_protein_encoder = "all"
            config.protein_encoder_num_params = nparams
            config.protein_pooling_opt = pooling_method
            config.long_protein_strategy = "split"
            config.max_protein_len = 1024
            config.protein_enc_batch_limit = None
            config.protein_pooling_correction_option = protein_pooling_correction_option
        path = os.path.join(checkpoint_dir, state_dict_relative_path)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found; consolidate DeepSpeed ZeRO shards with zero_to_fp32 first "
                                    "(deepspeed is not a dependency of this build)")
        state_dict = torch.load(path, map_location="cpu", weights_only=False)
        if model is None:
            # stale DATA_DIR prefixes of the training machine -> the current DATA_DIR (reference :1372)
            data_args_path = os.path.join(checkpoint_dir, "data_args.pt")
            if os.path.exists(data_args_path):
                from ..training.training_args_IT import update_model_args_data_dir

                data_args = torch.load(data_args_path, weights_only=False)
                update_model_args_data_dir(config, prev_data_dir=getattr(data_args, "data_dir", None))
            model = UnifiedProCyon(pretrained_weights_dir=pretrained_weights_dir, config=config,
                                   for_pretraining=False, **model_kwargs)
            model.load_state_dict(state_dict, strict=strict_load)
            return model, config
        model.load_state_dict(state_dict, strict=strict_load)
        return

    @staticmethod
    def get_checkpoint_configs(resume_from_checkpoint):
        from .. import compat

        compat.install()
        load = lambda n: torch.load(os.path.join(resume_from_checkpoint, n), weights_only=False)
        return load("data_args.pt"), load("model_args.pt"), load("training_args.pt")

    def save_pretrained(self, output_dir=None):
        state_dict = unwrap_model(self).state_dict()
        if output_dir is None:
            return state_dict, self.config
        torch.save(state_dict, os.path.join(output_dir, "txllm_model_ckpt.pt"))
        torch.save(self.config, os.path.join(output_dir, "model_args.pt"))

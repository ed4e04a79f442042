"""ESM2 protein encoder on the B200 kernels — host-side mirror of procyon/model/esm.py.

`ESM_PLM` keeps the reference's constructor arguments, attribute names, `forward(tokens, aggregate)` contract and
parameter names (`model.*` = fair-esm ESM2 names, so reference checkpoints load unchanged), but the forward
pass is one call into libprocyon_b200.so (`pcy_esm_encode` + `pcy_pool_segments`); there is no PyTorch
implementation of the math in this file and no fallback.

Reference call sites replaced: ESM_PLM.__init__ (esm.py:318-499), ESM_PLM.forward (esm.py:504-558),
ProteinPooler (esm.py:131-217).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from .. import _lib
from .._lib import c_i64, c_int, check, ptr, stream_ptr
from ..training.train_utils import batched_split_long_seq, reverse_batched_split

# (n_layers, d_model, n_heads) — procyon/model/esm.py:378-403 and fair-esm's published configs
ESM2_CONFIGS = {
    "8m": (6, 320, 20),
    "35m": (12, 480, 20),
    "150m": (30, 640, 20),
    "650m": (33, 1280, 20),
    "3b": (36, 2560, 40),
    "15b": (48, 5120, 40),
}
ESM_VOCAB, ESM_CLS, ESM_PAD, ESM_EOS, ESM_MASK = 33, 0, 1, 2, 32


class _EsmConfigC(ctypes.Structure):
    _fields_ = [("n_layers", c_int), ("d_model", c_int), ("n_heads", c_int), ("ffn_dim", c_int), ("vocab", c_int),
                ("pad_idx", c_int), ("mask_idx", c_int), ("token_dropout", c_int), ("ln_eps", ctypes.c_float)]


_KIND = dict(EMBED=0, LNF_G=1, LNF_B=2, LN1_G=3, LN1_B=4, WQKV=5, BQKV=6, WO=7, BO=8, LN2_G=9, LN2_B=10, W1=11,
             B1=12, W2=13, B2=14)


def rope_cos_sin_table(n_pos: int, head_dim: int, theta: float = 10000.0,
                       table_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """[n_pos, head_dim/2, 2] fp32 (cos, sin) as fair-esm RotaryEmbedding builds it.

    table_dtype=torch.bfloat16 reproduces the reference after `model.bfloat16()` (the inv_freq buffer and the
    position vector are cast to bf16 before the outer product); float32 is the exact table (default).
    """
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    inv_freq = inv_freq.to(table_dtype)
    t = torch.arange(n_pos).to(table_dtype)
    freqs = torch.outer(t, inv_freq)
    return torch.stack([freqs.cos().float(), freqs.sin().float()], dim=-1).contiguous()


class _Linear(nn.Module):
    """Parameter container with nn.Linear's state_dict layout (never called: the kernels read the packed copy)."""

    def __init__(self, in_f: int, out_f: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(out_f), requires_grad=False) if bias else None


class _LayerNormP(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(d), requires_grad=False)


class _SelfAttnP(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (_Linear(d, d) for _ in range(4))


class _EsmLayerP(nn.Module):
    def __init__(self, d: int, ffn: int):
        super().__init__()
        self.self_attn = _SelfAttnP(d)
        self.self_attn_layer_norm = _LayerNormP(d)
        self.fc1 = _Linear(d, ffn)
        self.fc2 = _Linear(ffn, d)
        self.final_layer_norm = _LayerNormP(d)


class _EsmLMHeadP(nn.Module):
    """fair-esm RobertaLMHead parameters: dense -> gelu -> layer_norm -> tied decoder (+ bias)."""

    def __init__(self, d: int, embed_weight: nn.Parameter):
        super().__init__()
        self.dense = _Linear(d, d)
        self.layer_norm = _LayerNormP(d)
        self.weight = embed_weight  # shared with embed_tokens.weight, as in fair-esm
        self.bias = nn.Parameter(torch.zeros(embed_weight.shape[0]), requires_grad=False)


class ESM2Params(nn.Module):
    """fair-esm `ESM2` parameter tree (names only): embed_tokens, layers.N.*, emb_layer_norm_after."""

    def __init__(self, n_layers: int, d: int, n_heads: int, ffn: Optional[int] = None):
        super().__init__()
        self.num_layers, self.embed_dim, self.attention_heads = n_layers, d, n_heads
        self.ffn_dim = ffn or 4 * d
        self.embed_tokens = nn.Embedding(ESM_VOCAB, d, padding_idx=ESM_PAD)
        self.embed_tokens.weight.requires_grad_(False)
        self.layers = nn.ModuleList([_EsmLayerP(d, self.ffn_dim) for _ in range(n_layers)])
        self.emb_layer_norm_after = _LayerNormP(d)
        self.lm_head = _EsmLMHeadP(d, self.embed_tokens.weight)
        self.token_dropout = True
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.normal_(p, std=0.02)
            elif p is not self.emb_layer_norm_after.weight and p is not self.lm_head.layer_norm.weight and not any(
                    p is l.self_attn_layer_norm.weight or p is l.final_layer_norm.weight for l in self.layers):
                nn.init.zeros_(p)


class ProteinPooler(nn.Module):
    """procyon/model/esm.py:131-217 — segmented mean/max over non-pad tokens, one kernel launch."""

    def __init__(self, pooling_method: str = "mean", protein_pooling_correction_option: bool = True):
        super().__init__()
        self.pooling_method = pooling_method.lower()
        self.protein_pooling_correction_option = protein_pooling_correction_option
        if self.pooling_method not in ("max", "mean", "cls_token"):
            raise NotImplementedError(f"Protein pooling method {self.pooling_method} is not implemented")

    def forward(self, protein_embeds: torch.Tensor, batch_keys: Optional[torch.Tensor] = None,
                padding_mask: Optional[torch.Tensor] = None, tokens: Optional[torch.Tensor] = None,
                padding_idx: int = ESM_PAD, out_fp32: bool = False) -> torch.Tensor:
        """protein_embeds bf16 [B', T, d] on CUDA. Either `tokens` (int32 [B',T]) or `padding_mask` is needed."""
        lib = _lib.load()
        _lib.require_cuda(protein_embeds)
        if protein_embeds.dtype != torch.bfloat16:
            raise TypeError("ProteinPooler expects bfloat16 residue states")
        Bp, T, d = protein_embeds.shape
        if self.pooling_method == "cls_token":
            return protein_embeds[:, 0, :]
        if batch_keys is None:
            batch_keys = torch.arange(Bp, dtype=torch.int64)
        if tokens is None:
            if padding_mask is None:
                raise ValueError("need tokens or padding_mask")
            tokens = torch.where(padding_mask, padding_idx, padding_idx + 1).to(torch.int32)
        tokens = tokens.to(device=protein_embeds.device, dtype=torch.int32).contiguous()
        seg_ptr, seg_rows, n_out = _segments(batch_keys)
        dev = protein_embeds.device
        seg_ptr_d = torch.from_numpy(seg_ptr).to(dev, non_blocking=True)
        seg_rows_d = torch.from_numpy(seg_rows).to(dev, non_blocking=True)
        out = torch.empty((n_out, d), device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
        x = protein_embeds.contiguous()
        check(lib.pcy_pool_segments(ptr(x), ptr(tokens), ptr(seg_ptr_d), ptr(seg_rows_d), ptr(out),
                                    c_int(1 if out_fp32 else 0), c_int(T), c_int(d), c_int(n_out), c_int(padding_idx),
                                    c_int(1 if self.pooling_method == "max" else 0),
                                    c_int(1 if (self.protein_pooling_correction_option and
                                                self.pooling_method == "mean") else 0),
                                    stream_ptr(dev)), "pcy_pool_segments")
        return out


def _segments(batch_keys: torch.Tensor):
    """CSR of chunk rows per protein id (ascending ids, gaps skipped; rows in ascending row order)."""
    keys = batch_keys.detach().cpu().numpy().astype(np.int64)
    order = np.argsort(keys, kind="stable").astype(np.int32)
    uniq, counts = np.unique(keys, return_counts=True)
    seg_ptr = np.zeros(len(uniq) + 1, dtype=np.int32)
    np.cumsum(counts, out=seg_ptr[1:])
    return seg_ptr, order, len(uniq)


class ESM_PLM(nn.Module):
    """Drop-in for procyon.model.esm.ESM_PLM (fair-esm sizes 8m/35m/150m/650m/3b/15b).

    Out of scope, as in SURVEY §2 row 2: LoRA / QLoRA / prefix / 'official' HF variants (ProCyon-Full sets none).
    """

    def __init__(
        self,
        pretrained_weights_dir=None,
        num_params="3b",
        pooling_method="max",
        padding_idx=1,
        eos_idx=2,
        max_protein_len=1024,
        long_protein_strategy="split",
        max_batch_forward_pass=None,
        use_lora=False,
        use_q_lora=False,
        use_task_spc_lora=False,
        lora_alpha=8,
        lora_r=8,
        use_adapter=False,
        adapter_rank=8,
        use_prefix=False,
        prefix_dropout=0.0,
        prefix_mid_dim=800,
        prefix_attn_bn=30,
        protein_attention_type="vanilla",
        lora_parameters="default",
        lora_num=2,
        protein_pooling_correction_option=False,
        rope_table_dtype: torch.dtype = torch.float32,
        custom_config=None,
    ):
        super().__init__()
        if use_lora or use_q_lora or use_task_spc_lora or use_adapter or use_prefix:
            raise NotImplementedError("LoRA / adapter / prefix variants of the protein encoder are out of scope")
        self.num_params = num_params.lower()
        if custom_config is not None:
            n_layers, d, n_heads = custom_config[:3]
            ffn = custom_config[3] if len(custom_config) > 3 else 4 * d
        elif self.num_params in ESM2_CONFIGS:
            n_layers, d, n_heads = ESM2_CONFIGS[self.num_params]
            ffn = 4 * d
        else:
            raise ValueError(f"ESM model with {self.num_params} parameters is not implemented")
        assert not ((pooling_method == "cls_token") and (long_protein_strategy == "split")), \
            "Cannot use CLS token with split strategy"
        self.pooling_method = pooling_method
        self.padding_idx, self.eos_idx = padding_idx, eos_idx
        self.protein_pooling_correction_option = protein_pooling_correction_option
        self.pooler = ProteinPooler(pooling_method=pooling_method,
                                    protein_pooling_correction_option=protein_pooling_correction_option)
        self.long_protein_strategy = long_protein_strategy
        self.max_protein_len = max_protein_len
        self.max_batch_forward_pass = max_batch_forward_pass
        self.repr_layer = n_layers
        self.embedding_size = d
        self.rope_table_dtype = rope_table_dtype
        self.model = ESM2Params(n_layers, d, n_heads, ffn)
        self._handle = None
        self._packed_version = None
        self._rope_pos = 0
        self._workspace = None
        # micro-batch budget for the layer stack (tokens per pass); 256 Ki tokens = 4.0 GB workspace at d=1280
        # (measured on B200, 256 proteins x 514 tokens: 64 Ki 1258, 128 Ki 1275, 256 Ki 1293 proteins/s).
        # Passes are balanced (ceil(rows / n_passes) rows each) instead of full passes plus a sliver that cannot
        # fill the GPU.
        self.max_tokens_per_pass = 256 * 1024

    # ---- weight packing ------------------------------------------------------------------------------------
    def _param_version(self):
        # (storage address, in-place version) of every parameter: any optimizer step, load_state_dict, .to() or dtype
        # cast shows up here and triggers a re-pack.  The Parameter OBJECTS are listed once (walking the module tree
        # costs ~2 ms per call at 290 parameters - 10 % of a retrieval query); code that swaps a Parameter object for
        # another must call `_forget_parameters()`.
        ps = self.__dict__.get("_param_list")
        if ps is None:
            ps = self.__dict__["_param_list"] = list(self.model.parameters())
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _forget_parameters(self):
        self.__dict__.pop("_param_list", None)

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / dtype casts may swap Parameter objects
        self._forget_parameters()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):  # (assign=True swaps them)
        self._forget_parameters()
        return super().load_state_dict(*args, **kwargs)

    def _ensure_packed(self, device: torch.device):
        lib = _lib.load()
        ver = (self._param_version(), str(device))
        if self._handle is not None and self._packed_version == ver:
            return
        self.release()
        m = self.model
        cfg = _EsmConfigC(m.num_layers, m.embed_dim, m.attention_heads, m.ffn_dim, ESM_VOCAB, self.padding_idx,
                          ESM_MASK, 1 if m.token_dropout else 0, 1e-5)
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(lib.pcy_esm_create(ctypes.byref(cfg), ctypes.byref(handle)), "pcy_esm_create")
            self._handle = handle

            def put(kind, layer, t, dtype):
                t = t.detach().to(device=device, dtype=dtype).contiguous()
                # the library copies with a blocking cudaMemcpy on the legacy stream, which does NOT order itself after
                # a non-default (non-blocking) torch stream: finish the conversions that produced `t` first
                torch.cuda.current_stream(device).synchronize()
                check(lib.pcy_esm_load_tensor(handle, c_int(_KIND[kind]), c_int(layer), ptr(t),
                                              c_i64(t.numel() * t.element_size())), f"pcy_esm_load_tensor({kind})")

            bf, f32 = torch.bfloat16, torch.float32
            put("EMBED", 0, m.embed_tokens.weight, bf)
            put("LNF_G", 0, m.emb_layer_norm_after.weight, bf)
            put("LNF_B", 0, m.emb_layer_norm_after.bias, bf)
            for l, y in enumerate(m.layers):
                a = y.self_attn
                put("LN1_G", l, y.self_attn_layer_norm.weight, bf)
                put("LN1_B", l, y.self_attn_layer_norm.bias, bf)
                put("WQKV", l, torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0), bf)
                # biases are stored in the module dtype by the reference; keep that rounding, then widen
                put("BQKV", l, torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).to(bf), f32)
                put("WO", l, a.out_proj.weight, bf)
                put("BO", l, a.out_proj.bias.to(bf), f32)
                put("LN2_G", l, y.final_layer_norm.weight, bf)
                put("LN2_B", l, y.final_layer_norm.bias, bf)
                put("W1", l, y.fc1.weight, bf)
                put("B1", l, y.fc1.bias.to(bf), f32)
                put("W2", l, y.fc2.weight, bf)
                put("B2", l, y.fc2.bias.to(bf), f32)
        self._packed_version = ver
        self._rope_pos = 0

    def _ensure_rope(self, T: int, device):
        if T <= self._rope_pos:
            return
        n_pos = max(T, self.max_protein_len + 2)
        hd = self.model.embed_dim // self.model.attention_heads
        tab = rope_cos_sin_table(n_pos, hd, 10000.0, self.rope_table_dtype)
        with torch.cuda.device(device):
            check(_lib.load().pcy_esm_set_rope_table(self._handle, ptr(tab), c_int(n_pos)), "pcy_esm_set_rope_table")
        self._rope_pos = n_pos

    def release(self):
        if self._handle is not None:
            _lib.load().pcy_esm_destroy(self._handle)
            self._handle = None
            self._workspace = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # ---- encode ----------------------------------------------------------------------------------------------
    def encode_tokens(self, tokens: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tokens int [B,T] on CUDA -> residue states bf16 [B,T,d] (representations[repr_layer])."""
        lib = _lib.load()
        _lib.require_cuda(tokens)
        dev = tokens.device
        self._ensure_packed(dev)
        B, T = tokens.shape
        self._ensure_rope(T, dev)
        tok32 = tokens.to(torch.int32).contiguous()
        d = self.embedding_size
        if out is None:
            out = torch.empty((B, T, d), device=dev, dtype=torch.bfloat16)
        need = lib.pcy_esm_workspace_bytes
        need.restype = ctypes.c_int64
        nbytes = need(self._handle, c_int(B), c_int(T))
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != dev:
            self._workspace = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        check(lib.pcy_esm_encode(self._handle, ptr(tok32), c_int(B), c_int(T), ptr(out), ptr(self._workspace),
                                 c_i64(self._workspace.numel()), stream_ptr(dev)), "pcy_esm_encode")
        return out

    def lm_head_logits(self, z: torch.Tensor) -> torch.Tensor:
        """fair-esm RobertaLMHead on residue states [B, T, d] -> logits [B, T, 33] in z's dtype:
        layer_norm(gelu(dense(z))) @ embed_tokens.weight^T + bias (procyon/model/esm.py:549-555 passes them on)."""
        from .. import ops

        B, T, d = z.shape
        hp = self.model.lm_head
        dev = z.device
        x = z.reshape(B * T, d).contiguous()
        h = ops.linear(x, hp.dense.weight.detach().to(dev, torch.bfloat16), hp.dense.bias.detach().to(dev, torch.float32),
                       act=ops.ACT_GELU)
        h = ops.layernorm(h, hp.layer_norm.weight.detach(), hp.layer_norm.bias.detach(), eps=1e-5)
        logits = ops.linear(h, hp.weight.detach().to(dev, torch.bfloat16), hp.bias.detach().to(dev, torch.float32),
                            out_fp32=True)
        return logits.view(B, T, -1).to(z.dtype)

    def forward(self, tokens: torch.Tensor, aggregate: bool = True):
        """Same contract as the reference: returns (z, logits). On the pooled path `logits` is None — every caller
        discards them there (procyon/model/model_unified.py:391), so they are not computed."""
        _lib.require_cuda(tokens)
        if self.long_protein_strategy == "split":
            batch_tokens, batch_keys, eos_loc = batched_split_long_seq(
                tokens, padding_idx=self.padding_idx, eos_idx=self.eos_idx,
                long_protein_strategy="split", max_protein_len=self.max_protein_len)
        else:
            batch_tokens, _, _ = batched_split_long_seq(
                tokens, padding_idx=self.padding_idx, eos_idx=self.eos_idx,
                long_protein_strategy=self.long_protein_strategy, max_protein_len=self.max_protein_len)
            # the reference pooler crashes on batch_keys=None (esm.py:158,217); one protein per row is the intent
            batch_keys, eos_loc = torch.arange(batch_tokens.shape[0], dtype=torch.int64), None
        Bp, T = batch_tokens.shape
        d = self.embedding_size
        dev = tokens.device
        if not aggregate:  # residue states + masked-LM logits (the `return_mlm` path, model_unified.py:505-509)
            z = self.encode_tokens(batch_tokens)
            logits = self.lm_head_logits(z)
            if eos_loc is not None and Bp != tokens.shape[0]:
                z = reverse_batched_split(z, batch_keys, eos_locs=eos_loc)
                logits = reverse_batched_split(logits, batch_keys, eos_locs=eos_loc)
            return z, logits

        rows_per_pass = max(1, self.max_tokens_per_pass // T)
        if self.max_batch_forward_pass is not None:
            rows_per_pass = min(rows_per_pass, int(self.max_batch_forward_pass))
        if Bp <= rows_per_pass:
            z = self.encode_tokens(batch_tokens)
            return self.pooler(z, batch_keys=batch_keys, tokens=batch_tokens, padding_idx=self.padding_idx), None
        rows_per_pass = -(-Bp // -(-Bp // rows_per_pass))  # same number of passes, equal sizes

        # micro-batched: keep all chunks of a protein in the same pass so pooling stays local to the pass
        keys_np = batch_keys.cpu().numpy()
        order = np.argsort(keys_np, kind="stable")
        sorted_keys = keys_np[order]
        uniq = np.unique(sorted_keys)
        out = torch.empty((len(uniq), d), device=dev, dtype=torch.bfloat16)
        order_t = torch.from_numpy(order).to(dev)
        toks_sorted = batch_tokens.index_select(0, order_t)
        start, n_done = 0, 0
        while start < Bp:
            end = min(Bp, start + rows_per_pass)
            if end < Bp:  # do not cut through a protein
                while end > start and sorted_keys[end] == sorted_keys[end - 1]:
                    end -= 1
                if end == start:  # one protein larger than a pass: take it whole
                    end = start + int((sorted_keys == sorted_keys[start]).sum())
            chunk = toks_sorted[start:end]
            z = self.encode_tokens(chunk)
            ck = torch.from_numpy(sorted_keys[start:end])
            pooled = self.pooler(z, batch_keys=ck, tokens=chunk, padding_idx=self.padding_idx)
            out[n_done : n_done + pooled.shape[0]] = pooled
            n_done += pooled.shape[0]
            start = end
        return out, None

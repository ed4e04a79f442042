"""Llama decoder on the B200 kernels — host-side mirror of procyon/model/pmc_llama.py.

`LlamaPostTokenization` keeps the reference's `forward(input_embeds | input_ids, attn_masks, full_labels,
past_key_values, use_cache, output_attentions)` contract and the HF parameter names under `.model`
(`model.embed_tokens`, `model.layers.N.{self_attn,mlp,...}`, `model.norm`, `lm_head`), so `state_dict` keys are
those of the reference checkpoint.  All math runs in libprocyon_b200.so; nothing here falls back to
PyTorch.

Out of scope: FlashLlamaAttention / pipeline wrappers / (Q)LoRA — not reached by ProCyon-Full.
"""
from __future__ import annotations

import ctypes
import json
import os
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from .. import _lib
from .._lib import c_float, c_i64, c_int, check, ptr, stream_ptr


@dataclass
class LlamaConfig:
    """The fields of HF `LlamaConfig` this path reads (names as in config.json)."""

    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    vocab_size: int = 128256
    rms_norm_eps: float = 1e-5
    max_position_embeddings: int = 8192
    # transformers 4.31 (the reference pin) ignores config.json's rope_theta and uses base 10000
    rope_theta: float = 10000.0

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @classmethod
    def from_json(cls, path: str, honour_rope_theta: bool = False) -> "LlamaConfig":
        with open(path) as f:
            raw = json.load(f)
        kw = {k: raw[k] for k in cls.__dataclass_fields__ if k in raw and k != "rope_theta"}
        cfg = cls(**kw)
        if honour_rope_theta and "rope_theta" in raw:
            cfg.rope_theta = float(raw["rope_theta"])
        return cfg


LLAMA3_8B = LlamaConfig()


class _LlamaConfigC(ctypes.Structure):
    _fields_ = [("n_layers", c_int), ("d_model", c_int), ("n_heads", c_int), ("n_kv_heads", c_int),
                ("head_dim", c_int), ("ffn_dim", c_int), ("vocab", c_int), ("rms_eps", ctypes.c_float)]


class DecodeBuffersC(ctypes.Structure):
    _fields_ = [("n_inputs", c_int), ("beams", c_int), ("S", c_int), ("max_gen", c_int),
                ("kv_prompt", ctypes.c_void_p), ("prompt_valid", ctypes.c_void_p), ("kv_gen", ctypes.c_void_p),
                ("tokens", ctypes.c_void_p), ("slots", ctypes.c_void_p), ("logprobs", ctypes.c_void_p),
                ("logits_cur", ctypes.c_void_p), ("logits_hist", ctypes.c_void_p), ("state", ctypes.c_void_p),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64)]


_KIND = dict(EMBED=0, LM_HEAD=1, NORM=2, LN1=3, LN2=4, WQKV=5, WO=6, WGATEUP=7, WDOWN=8)
SELECT_GREEDY, SELECT_BEAM = 0, 1


def llama_rope_table(n_pos: int, head_dim: int, theta: float, table_dtype: torch.dtype = torch.float32):
    """[n_pos, head_dim/2, 2] fp32 (cos, sin). HF LlamaRotaryEmbedding computes the tables in fp32 at init; after
    `.bfloat16()` the VALUES are rounded to bf16 (positions stay exact) — table_dtype=bfloat16 reproduces that."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2).float() / head_dim))
    freqs = torch.outer(torch.arange(n_pos, dtype=torch.float32), inv_freq)
    return torch.stack([freqs.cos().to(table_dtype).float(), freqs.sin().to(table_dtype).float()], -1).contiguous()


class _W(nn.Module):
    def __init__(self, out_f, in_f, device=None, dtype=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f, device=device, dtype=dtype), requires_grad=False)


class _Norm(nn.Module):
    def __init__(self, d, device=None, dtype=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d, device=device, dtype=dtype), requires_grad=False)


class _Attn(nn.Module):
    def __init__(self, c: LlamaConfig, device, dtype):
        super().__init__()
        d, hd = c.hidden_size, c.head_dim
        self.q_proj = _W(c.num_attention_heads * hd, d, device, dtype)
        self.k_proj = _W(c.num_key_value_heads * hd, d, device, dtype)
        self.v_proj = _W(c.num_key_value_heads * hd, d, device, dtype)
        self.o_proj = _W(d, c.num_attention_heads * hd, device, dtype)


class _MLP(nn.Module):
    def __init__(self, c: LlamaConfig, device, dtype):
        super().__init__()
        self.gate_proj = _W(c.intermediate_size, c.hidden_size, device, dtype)
        self.up_proj = _W(c.intermediate_size, c.hidden_size, device, dtype)
        self.down_proj = _W(c.hidden_size, c.intermediate_size, device, dtype)


class _Layer(nn.Module):
    def __init__(self, c, device, dtype):
        super().__init__()
        self.self_attn = _Attn(c, device, dtype)
        self.mlp = _MLP(c, device, dtype)
        self.input_layernorm = _Norm(c.hidden_size, device, dtype)
        self.post_attention_layernorm = _Norm(c.hidden_size, device, dtype)


class _LlamaModelP(nn.Module):
    def __init__(self, c, device, dtype):
        super().__init__()
        self.embed_tokens = nn.Embedding(c.vocab_size, c.hidden_size, device=device, dtype=dtype)
        self.embed_tokens.weight.requires_grad_(False)
        self.layers = nn.ModuleList([_Layer(c, device, dtype) for _ in range(c.num_hidden_layers)])
        self.norm = _Norm(c.hidden_size, device, dtype)


class LlamaForCausalLMParams(nn.Module):
    """Parameter tree with HF `LlamaForCausalLM` names + the few methods UnifiedProCyon calls on it."""

    def __init__(self, config: LlamaConfig, device=None, dtype=None, init_std: Optional[float] = 0.02):
        super().__init__()
        self.config = config
        self.vocab_size = config.vocab_size
        self.model = _LlamaModelP(config, device, dtype)
        self.lm_head = _W(config.vocab_size, config.hidden_size, device, dtype)
        if init_std is not None:
            for n, p in self.named_parameters():
                if p.dim() > 1:
                    nn.init.normal_(p, std=init_std)

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def resize_token_embeddings(self, new_num_tokens: int):
        """HF semantics: keep existing rows, new rows ~ N(0, 0.02) (model_unified.py:166)."""
        old = self.model.embed_tokens.weight
        if new_num_tokens == old.shape[0]:
            return self.model.embed_tokens
        for mod in (self.model.embed_tokens, self.lm_head):
            w = mod.weight
            new = torch.empty(new_num_tokens, w.shape[1], device=w.device, dtype=w.dtype)
            nn.init.normal_(new, std=0.02)
            n = min(new_num_tokens, w.shape[0])
            new[:n] = w.data[:n]
            mod.weight = nn.Parameter(new, requires_grad=False)
        self.model.embed_tokens.num_embeddings = new_num_tokens
        self.vocab_size = new_num_tokens
        self.config.vocab_size = new_num_tokens
        return self.model.embed_tokens


class _HiddenStates:
    """Stands in for HF's tuple of L+1 hidden states: only the last (post-final-norm) one is materialised."""

    def __init__(self, last: torch.Tensor, n: int):
        self._last, self._n = last, n

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if i == -1 or i == self._n - 1:
            return self._last
        raise NotImplementedError("only hidden_states[-1] is materialised (ret_token_access='last')")


class CausalLMOutput:
    """Duck-type of HF CausalLMOutputWithPast: .logits .loss .hidden_states .past_key_values"""

    def __init__(self, owner, hidden_last, n_layers, logits=None, loss=None, past=None):
        self._owner = owner
        self.hidden_states = _HiddenStates(hidden_last, n_layers + 1) if hidden_last is not None else None
        self._logits = logits
        self.loss = loss
        self.past_key_values = past

    def logits_at(self, positions: torch.Tensor) -> torch.Tensor:
        """fp32 logits [B, V] of one position per sequence: the LM head runs on B gathered rows only (QA scoring
        needs the yes/no decision at a single position; the reference softmaxes the whole (B, S, V) tensor on the
        CPU for it, procyon/training/train_utils.py:1048-1070)."""
        if self._logits is not None:
            return self._logits[torch.arange(self._logits.shape[0], device=self._logits.device), positions]
        h = self.hidden_states[-1]
        rows = h[torch.arange(h.shape[0], device=h.device), positions.to(h.device)]
        return self._owner.lm_head_logits(rows.contiguous())

    @property
    def logits(self):
        if self._logits is None:  # full-sequence logits are only computed when somebody asks for them
            h = self.hidden_states[-1]
            B, S, d = h.shape
            self._logits = self._owner.lm_head_logits(h.reshape(B * S, d)).view(B, S, -1)
        return self._logits


class DecodeSession:
    """Device-resident generation state (KV caches, token / ancestry tables, logits) for one generate() call."""

    def __init__(self, owner: "LlamaPostTokenization", n_inputs: int, beams: int, S: int, max_gen: int,
                 device, masked: bool, keep_logits: bool):
        lib = _lib.load()
        self.owner = owner
        c = owner.model.config
        dev = torch.device(device)
        rows = n_inputs * beams
        if rows > 16:
            raise _lib.ProcyonB200Error(f"n_inputs*beams = {rows} > 16 rows per decode session; split the batch")
        kvd = c.num_key_value_heads * c.head_dim
        V = owner.model.vocab_size
        self.n_inputs, self.beams, self.S, self.max_gen, self.rows, self.V = n_inputs, beams, S, max_gen, rows, V
        self.kv_prompt = torch.empty((c.num_hidden_layers, 2, n_inputs, S, kvd), device=dev, dtype=torch.bfloat16)
        self.prompt_valid = torch.ones((n_inputs, S), device=dev, dtype=torch.uint8) if masked else None
        kv_prompt, prompt_valid = self.kv_prompt, self.prompt_valid
        self.kv_gen = torch.empty((c.num_hidden_layers, 2, rows, max_gen, kvd), device=dev, dtype=torch.bfloat16)
        self.tokens = torch.zeros((rows, max_gen), device=dev, dtype=torch.int32)
        self.slots = torch.zeros((rows, max_gen), device=dev, dtype=torch.int32)
        self.logprobs = torch.zeros((rows,), device=dev, dtype=torch.float32)
        self.logits_cur = torch.empty((rows, V), device=dev, dtype=torch.float32)
        self.logits_hist = torch.empty((max_gen, rows, V), device=dev, dtype=torch.float32) if keep_logits else None
        self.state = torch.zeros((8,), device=dev, dtype=torch.int32)
        f = lib.pcy_llama_decode_workspace_bytes
        f.restype = ctypes.c_int64
        nbytes = f(owner._handle, c_int(rows), c_int(S), c_int(max_gen))
        self.workspace = torch.zeros(nbytes, device=dev, dtype=torch.uint8)
        self.c = DecodeBuffersC(n_inputs, beams, S, max_gen, kv_prompt.data_ptr(),
                                prompt_valid.data_ptr() if prompt_valid is not None else None,
                                self.kv_gen.data_ptr(), self.tokens.data_ptr(), self.slots.data_ptr(),
                                self.logprobs.data_ptr(), self.logits_cur.data_ptr(),
                                self.logits_hist.data_ptr() if keep_logits else None, self.state.data_ptr(),
                                self.workspace.data_ptr(), nbytes)
        self.device = dev
        self._graph = None
        self._graph_key = None
        self.graph_launches = 0

    def reset(self, prefill_logits: Optional[torch.Tensor]):
        check(_lib.load().pcy_decode_reset(self.owner._handle, ctypes.byref(self.c), ptr(prefill_logits),
                                           stream_ptr(self.device)), "pcy_decode_reset")

    def forward(self):
        check(_lib.load().pcy_llama_decode_forward(self.owner._handle, ctypes.byref(self.c), stream_ptr(self.device)),
              "pcy_llama_decode_forward")

    def select(self, mode: int, group: int = 1, penalty: float = 0.0, eos_id: int = -1, stop_on_all_eos: bool = False,
               group_state: Optional[torch.Tensor] = None, group_last: bool = True):
        """`group_state` (int32 [4] on the device, [3] = total inputs): shared by the sessions that split one batch and
        step in lock-step, so that the all-EOS stop is taken over the WHOLE batch like the reference's (:833)."""
        check(_lib.load().pcy_decode_select_group(self.owner._handle, ctypes.byref(self.c), c_int(mode), c_int(group),
                                                  c_float(penalty), c_int(eos_id), c_int(1 if stop_on_all_eos else 0),
                                                  ptr(group_state), c_int(1 if group_last else 0),
                                                  stream_ptr(self.device)), "pcy_decode_select_group")

    def step_graph(self, mode, group, penalty, eos_id, stop_on_all_eos):
        """CUDA graph of one (forward, select) step; the step index lives in device memory so it replays as is."""
        key = (mode, group, float(penalty), eos_id, bool(stop_on_all_eos))
        if self._graph is not None and self._graph_key == key:
            return self._graph
        lib = _lib.load()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        n0 = lib.pcy_launch_count()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self.forward()
                self.select(mode, group, penalty, eos_id, stop_on_all_eos)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.graph_launches = lib.pcy_launch_count() - n0  # kernels replayed by every g.replay()
        self._graph, self._graph_key = g, key
        return g


class SessionGroup:
    """More than 16 rows of token-by-token decoding (`forward(use_cache=True)` on a large batch, sampling loops): the
    rows are split over DecodeSessions of <= 16 rows that are stepped one after the other.  Presents the few members
    the stepping code uses (`rows`, `device`, `step(tokens)`, `logits()`)."""

    def __init__(self, sessions):
        self.sessions = sessions
        self.rows = sum(s.rows for s in sessions)
        self.device = sessions[0].device
        self.t = 0  # tokens fed so far (host-side copy of state[0]; the sessions advance together)

    def logits(self) -> torch.Tensor:
        return torch.cat([s.logits_cur for s in self.sessions], 0)

    def step(self, tokens: torch.Tensor) -> torch.Tensor:
        """Feeds one token per row (int tensor [rows]) and returns the next-token logits fp32 [rows, V]."""
        r0 = 0
        for s in self.sessions:
            s.tokens[:, self.t] = tokens[r0:r0 + s.rows].to(s.tokens.dtype)
            s.slots[:, self.t] = torch.arange(s.rows, device=s.device, dtype=torch.int32)
            s.state[0] = self.t + 1
            s.forward()
            r0 += s.rows
        self.t += 1
        return self.logits()


class LlamaPostTokenization(nn.Module):
    def __init__(
        self,
        model_path: str = "llama-3-8b",
        max_gen_len=None,
        model_splitting=False,
        n_model_pieces=2,
        attention_type="vanilla",
        use_lora=False,
        use_q_lora=False,
        lora_r=16,
        lora_alpha=8,
        use_task_spc_lora=False,
        lora_num=2,
        for_pretraining=True,
        config: Optional[LlamaConfig] = None,
        device=None,
        dtype=None,
        rope_table_dtype: torch.dtype = torch.float32,
    ):
        super().__init__()
        if use_lora or use_q_lora or use_task_spc_lora or model_splitting:
            raise NotImplementedError("LoRA / QLoRA / pipeline-split text encoders are out of scope")
        self.model_path = model_path
        self.attention_type = attention_type
        if config is None:
            cfg_json = None
            for root in (os.getenv("LLAMA3_PATH"), model_path):
                if root and os.path.exists(os.path.join(root, "config.json")):
                    cfg_json = os.path.join(root, "config.json")
                    break
            config = LlamaConfig.from_json(cfg_json) if cfg_json else LlamaConfig()
        # weights: the reference loads the base checkpoint when for_pretraining, else builds from config and
        # lets from_pretrained() fill it (pmc_llama.py:478-484); loading base weights is the caller's job here.
        self.model = LlamaForCausalLMParams(config, device=device, dtype=dtype)
        self.rope_table_dtype = rope_table_dtype
        self._handle = None
        self._packed_version = None
        self._rope_pos = 0
        self._ws = None
        self.kv_cache = None

    # ---- packing ---------------------------------------------------------------------------------------------
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

    def _ensure_packed(self, device):
        lib = _lib.load()
        ver = (self._param_version(), str(device))
        if self._handle is not None and self._packed_version == ver:
            return
        self.release()
        c = self.model.config
        cfg = _LlamaConfigC(c.num_hidden_layers, c.hidden_size, c.num_attention_heads, c.num_key_value_heads,
                            c.head_dim, c.intermediate_size, self.model.vocab_size, c.rms_norm_eps)
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(lib.pcy_llama_create(ctypes.byref(cfg), ctypes.byref(h)), "pcy_llama_create")
            self._handle = h
            bf = torch.bfloat16

            def put(kind, layer, t):
                t = t.detach().to(device=device, dtype=bf).contiguous()
                # the library copies with a blocking cudaMemcpy on the legacy stream, which does NOT order itself after
                # a non-default (non-blocking) torch stream: finish the cat / pack kernels that produced `t` first
                torch.cuda.current_stream(device).synchronize()
                check(lib.pcy_llama_load_tensor(h, c_int(_KIND[kind]), c_int(layer), ptr(t),
                                                c_i64(t.numel() * 2)), f"pcy_llama_load_tensor({kind})")

            m = self.model.model
            put("EMBED", 0, m.embed_tokens.weight)
            put("LM_HEAD", 0, self.model.lm_head.weight)
            put("NORM", 0, m.norm.weight)
            from .. import ops

            for l, y in enumerate(m.layers):
                a, f = y.self_attn, y.mlp
                put("LN1", l, y.input_layernorm.weight)
                put("LN2", l, y.post_attention_layernorm.weight)
                put("WQKV", l, torch.cat([a.q_proj.weight.to(device=device, dtype=bf),
                                          a.k_proj.weight.to(device=device, dtype=bf),
                                          a.v_proj.weight.to(device=device, dtype=bf)], 0))
                put("WO", l, a.o_proj.weight)
                put("WGATEUP", l, ops.pack_gate_up(f.gate_proj.weight.detach().to(device=device, dtype=bf),
                                                   f.up_proj.weight.detach().to(device=device, dtype=bf)))
                put("WDOWN", l, f.down_proj.weight)
            torch.cuda.synchronize(device)
        self._packed_version = ver
        self._rope_pos = 0

    def _ensure_rope(self, n_pos, device):
        if n_pos <= self._rope_pos:
            return
        c = self.model.config
        n_pos = max(n_pos, min(c.max_position_embeddings, 4096))
        tab = llama_rope_table(n_pos, c.head_dim, c.rope_theta, self.rope_table_dtype)
        with torch.cuda.device(device):
            check(_lib.load().pcy_llama_set_rope_table(self._handle, ptr(tab), c_int(n_pos)), "pcy_llama_set_rope_table")
        self._rope_pos = n_pos

    def release(self):
        if self._handle is not None:
            _lib.load().pcy_llama_destroy(self._handle)
            self._handle = None
            self._ws = None
            self.__dict__.pop("_sessions", None)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, device=device, dtype=torch.uint8)
        return self._ws

    # ---- kernels-level entry points ---------------------------------------------------------------------------
    def prefill(self, input_embeds: torch.Tensor, attn_masks: Optional[torch.Tensor], *, want_cache: bool,
                want_hidden: bool, sel_rows: Optional[torch.Tensor] = None, kv_out: Optional[torch.Tensor] = None,
                acc_rows: Optional[torch.Tensor] = None):
        """input_embeds bf16 [B,S,d] on CUDA. Returns (kv_prompt | None, hidden [B,S,d] | None, sel_logits | None,
        valid).  With `acc_rows` (flat token rows) the fp32 sum of all L+1 hidden states of those rows is left in
        `self.last_hidden_sum` [n, d] (ret_token_access='all')."""
        lib = _lib.load()
        _lib.require_cuda(input_embeds)
        dev = input_embeds.device
        self._ensure_packed(dev)
        B, S_full, d = input_embeds.shape
        c = self.model.config
        valid = None
        if attn_masks is not None:
            valid = (attn_masks.to(dev) != 0).to(torch.uint8).contiguous()
        # Trailing padding.  The reference pads every forward() batch to max_text_len (2048 for ProCyon-Full,
        # model_unified.py:1283) and HF then runs all of those positions through the 32 layers.  Positions after the
        # last valid token of the longest row are masked as keys and never read as outputs (labels -100, no [PROT] /
        # answer token there), and with causal attention nothing before them depends on them: they are not computed
        # here; their hidden states come back as zeros (reference: values derived from pad embeddings).  Only for
        # forward passes without a KV cache (generation prompts are un-padded or left-padded).
        S = S_full
        if valid is not None and not want_cache and self.trim_trailing_pads and S_full > 1:
            pos = torch.arange(1, S_full + 1, device=dev, dtype=torch.int32)
            S = max(1, int((valid.to(torch.int32) * pos).amax().item()))
        self._ensure_rope(S_full, dev)
        if S < S_full:
            def remap(rows):  # flat token rows b * S_full + p -> b * S + p
                rows = rows.to(device=dev, dtype=torch.int64)
                b, p_ = rows // S_full, rows % S_full
                assert bool((p_ < S).all()), "a selected row lies in the trailing padding"
                return (b * S + p_).to(torch.int32)

            sel_rows = remap(sel_rows) if sel_rows is not None and sel_rows.numel() > 0 else sel_rows
            acc_rows = remap(acc_rows) if acc_rows is not None and acc_rows.numel() > 0 else acc_rows
            input_embeds = input_embeds[:, :S]
            valid_full, valid = valid, valid[:, :S].contiguous()
        x = input_embeds.to(torch.bfloat16).contiguous()
        kvd = c.num_key_value_heads * c.head_dim
        kv = None
        if want_cache:
            kv = kv_out if kv_out is not None else torch.empty((c.num_hidden_layers, 2, B, S, kvd), device=dev,
                                                               dtype=torch.bfloat16)
        hidden = torch.empty((B, S, d), device=dev, dtype=torch.bfloat16) if want_hidden else None
        n_sel, logits = 0, None
        if sel_rows is not None and sel_rows.numel() > 0:
            sel_rows = sel_rows.to(device=dev, dtype=torch.int32).contiguous()
            n_sel = sel_rows.numel()
            logits = torch.empty((n_sel, self.model.vocab_size), device=dev, dtype=torch.float32)
        f = lib.pcy_llama_prefill_workspace_bytes
        f.restype = ctypes.c_int64
        ws = self._workspace(f(self._handle, c_int(B), c_int(S)), dev)
        n_acc, acc = 0, None
        if acc_rows is not None and acc_rows.numel() > 0:
            assert want_hidden, "acc_rows needs want_hidden=True"
            acc_rows = acc_rows.to(device=dev, dtype=torch.int32).contiguous()
            n_acc = acc_rows.numel()
            acc = torch.empty((n_acc, d), device=dev, dtype=torch.float32)
        check(lib.pcy_llama_prefill_ex(self._handle, ptr(x), ptr(valid), c_int(B), c_int(S), ptr(kv), ptr(hidden),
                                       ptr(sel_rows) if n_sel else None, c_int(n_sel), ptr(logits),
                                       ptr(acc_rows) if n_acc else None, c_int(n_acc), ptr(acc), ptr(ws),
                                       c_i64(ws.numel()), stream_ptr(dev)), "pcy_llama_prefill_ex")
        self.last_hidden_sum = acc
        if S < S_full:
            if hidden is not None:
                full = torch.zeros((B, S_full, d), device=dev, dtype=hidden.dtype)
                full[:, :S] = hidden
                hidden = full
            valid = valid_full
        return kv, hidden, logits, valid

    def lm_head_logits(self, hidden_rows: torch.Tensor) -> torch.Tensor:
        """fp32 logits of already-normalised hidden rows [n,d] (LM head only)."""
        from .. import ops

        w = self._lm_head_bf16(hidden_rows.device)
        return ops.linear(hidden_rows.to(torch.bfloat16).contiguous(), w, out_fp32=True)

    def _lm_head_bf16(self, device):
        w = self.model.lm_head.weight
        if w.dtype == torch.bfloat16 and w.device == device:
            return w.detach()
        key = (w.data_ptr(), w._version)
        if getattr(self, "_lm_cache_key", None) != key:
            self._lm_cache = w.detach().to(device=device, dtype=torch.bfloat16)
            self._lm_cache_key = key
        return self._lm_cache

    # ---- per-SM slice calibration of the persistent decode kernel ----------------------------------------------
    _calibrated_devices = {}

    def calibrate_decode_shares(self, device, rounds: int = 3, iters: int = 3):
        """Measures, with the kernel's own per-CTA stamps, how fast each SM streams its slice of the long weight
        phases (gate/up of every layer, LM head) and hands the rates to the library (pcy_set_decode_sm_shares),
        which then sizes the slices accordingly.  With all SMs pulling from HBM at once some get up to ~20 % less
        bandwidth than others (systematic by SM id, different from GPU to GPU), and with equal slices every grid
        barrier waits for the slowest.  Results are unaffected.  Returns the shares (or None if SM ids are not
        0..n-1, e.g. under MIG: equal slices stay)."""
        lib = _lib.load()
        dev = torch.device(device)
        G = torch.cuda.get_device_properties(dev).multi_processor_count
        L = self.model.config.num_hidden_layers
        n_ph = 4 * L + 1
        self._ensure_packed(dev)
        self._ensure_rope(64, dev)
        sess = DecodeSession(self, 1, 1, 16, 4, dev, False, False)
        sess.kv_prompt.zero_()
        sess.kv_gen.zero_()
        sess.state[0] = 1
        buf = torch.zeros(4096 + n_ph * G * 2 + 64, device=dev, dtype=torch.int64)
        buf[4095] = 0x534B4557  # "SKEW": asks the kernel for the per-CTA phase-end stamps after the first 4096 words
        shares = torch.ones(G, dtype=torch.float32)
        set_shares = lib.pcy_set_decode_sm_shares
        long_phases = [4 * l + 2 for l in range(1, L)] + [4 * L]  # gate/up of every layer + the LM head
        ok = True
        try:
            for _ in range(rounds):
                check(set_shares(shares.numpy().ctypes.data_as(ctypes.c_void_p), c_int(G)), "pcy_set_decode_sm_shares")
                for _ in range(2):
                    sess.forward()
                lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(buf.data_ptr()))
                rate = torch.zeros(G, dtype=torch.float64)
                for _ in range(iters):
                    sess.forward()
                    torch.cuda.synchronize(dev)
                    t = buf[4096:4096 + n_ph * G * 2].cpu().view(n_ph, G, 2)
                    smid = t[long_phases[0], :, 1]
                    if sorted(smid.tolist()) != list(range(G)):
                        ok = False
                        break
                    for ph in long_phases:
                        start = t[ph - 1, :, 0].max()
                        dur = (t[ph, :, 0] - start).double().clamp_min(1.0)
                        sm = t[ph, :, 1]
                        rate[sm] += float(dur.mean()) * shares[sm].double() / dur  # longer phases weigh more
                lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(0))
                if not ok:
                    break
                new = (rate / rate.mean()).float()
                shares = (0.5 * shares + 0.5 * new).clamp(0.86, 1.13)  # damped: rates shift when the slices do
                shares = shares / shares.mean()
        finally:
            lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(0))
        if ok:
            check(set_shares(shares.numpy().ctypes.data_as(ctypes.c_void_p), c_int(G)), "pcy_set_decode_sm_shares")
        else:
            check(set_shares(None, c_int(0)), "pcy_set_decode_sm_shares")
        torch.cuda.synchronize(dev)
        return shares if ok else None

    def new_session(self, n_inputs, beams, S, max_gen, device, masked, keep_logits) -> DecodeSession:
        """An uncached session (batches split over several lock-step sessions need distinct ones of the same shape)."""
        self._ensure_packed(torch.device(device))
        self._ensure_rope(S + max_gen, device)
        return DecodeSession(self, n_inputs, beams, S, max_gen, device, masked, keep_logits)

    def get_session(self, n_inputs, beams, S, max_gen, device, masked, keep_logits) -> DecodeSession:
        """Decode sessions (KV caches, tables, captured step graph) are cached per shape and reused across calls."""
        self._ensure_packed(torch.device(device))
        self._ensure_rope(S + max_gen, device)
        c = self.model.config
        dkey = str(torch.device(device))
        if (n_inputs * beams <= 4 and c.hidden_size >= 2048 and c.num_hidden_layers >= 8 and c.head_dim == 128
                and dkey not in LlamaPostTokenization._calibrated_devices
                and os.environ.get("PCY_DECODE_CALIBRATE", "0") == "1"):
            LlamaPostTokenization._calibrated_devices[dkey] = True  # (set first: calibration builds a session itself)
            LlamaPostTokenization._calibrated_devices[dkey] = self.calibrate_decode_shares(device)
        key = (n_inputs, beams, S, max_gen, str(device), bool(masked), bool(keep_logits), id(self._handle))
        cache = self.__dict__.setdefault("_sessions", {})
        if key not in cache:
            if len(cache) >= 4:  # bound the memory held by cached sessions
                cache.pop(next(iter(cache)))
            cache[key] = DecodeSession(self, n_inputs, beams, S, max_gen, device, masked, keep_logits)
        return cache[key]

    def sessions_from_prefill(self, kv, valid, sel_logits, max_gen):
        """Token-by-token decode state after a prefill of B prompts (kv [L, 2, B, S, kvd], last-position logits
        [B, V]): one DecodeSession for B <= 16, else a SessionGroup of <= 16-row sessions."""
        L, _, B, S, _ = kv.shape
        dev = kv.device
        self._ensure_rope(S + max_gen, dev)
        out = []
        for r0 in range(0, B, 16):
            r1 = min(B, r0 + 16)
            sess = DecodeSession(self, r1 - r0, 1, S, max_gen, dev, valid is not None, keep_logits=False)
            sess.kv_prompt.copy_(kv[:, :, r0:r1])
            if valid is not None:
                sess.prompt_valid.copy_(valid[r0:r1])
            sess.reset(sel_logits[r0:r1].contiguous())
            out.append(sess)
        return out[0] if len(out) == 1 else SessionGroup(out)

    # compute only up to the last valid position of a padded forward() batch (see prefill); False = every position
    trim_trailing_pads = True

    # ---- reference-facing forward -------------------------------------------------------------------------------
    def forward(self, input_embeds=None, input_ids=None, attn_masks=None, full_labels=None, past_key_values=None,
                use_cache=False, output_attentions=None, sum_hidden_rows=None):
        """`sum_hidden_rows` (bool mask [B, S], extension): rows whose L+1 hidden states are summed on the fly
        (-> `out.hidden_sum` [n, d] fp32); the reference stacks all 33 full tensors for that (ret_token_access='all')."""
        assert (input_embeds is not None) != (input_ids is not None), \
            "Only one of input_embeds or input_ids can be provided"
        if output_attentions:
            raise NotImplementedError("attention maps are never materialised by the fused attention kernels")
        n_layers = self.model.config.num_hidden_layers
        if past_key_values is not None:
            # single decode step on an existing session (greedy-style stepping: one row per input)
            assert input_ids is not None and input_ids.shape[1] == 1
            if isinstance(past_key_values, SessionGroup):
                logits = past_key_values.step(input_ids[:, 0].to(past_key_values.device))
                return CausalLMOutput(self, None, n_layers, logits=logits.unsqueeze(1), past=past_key_values)
            sess: DecodeSession = past_key_values
            t = int(sess.state[0].item())
            sess.tokens[:, t] = input_ids[:, 0].to(sess.tokens.dtype)
            sess.slots[:, t] = torch.arange(sess.rows, device=sess.device, dtype=torch.int32)
            sess.state[0] = t + 1
            sess.forward()
            return CausalLMOutput(self, None, n_layers, logits=sess.logits_cur.clone().unsqueeze(1), past=sess)
        if input_ids is not None:
            emb = self.model.model.embed_tokens.weight
            input_embeds = emb[input_ids.to(emb.device)]
        B, S, _ = input_embeds.shape
        dev = input_embeds.device
        sel = None
        if use_cache:
            sel = torch.arange(B, device=dev, dtype=torch.int32) * S + (S - 1)
        acc_rows = None
        if sum_hidden_rows is not None:
            acc_rows = sum_hidden_rows.reshape(-1).nonzero().squeeze(-1)
        kv, hidden, sel_logits, valid = self.prefill(input_embeds, attn_masks, want_cache=use_cache, want_hidden=True,
                                                     sel_rows=sel, acc_rows=acc_rows)
        loss = None
        if full_labels is not None:
            from .model_utils import lm_loss

            loss = lm_loss(self, hidden, full_labels)
        past = None
        if use_cache:
            past = self.sessions_from_prefill(kv, valid if attn_masks is not None else None, sel_logits, 256)
        out = CausalLMOutput(self, hidden, n_layers, loss=loss, past=past)
        out.hidden_sum = self.last_hidden_sum if sum_hidden_rows is not None else None
        return out

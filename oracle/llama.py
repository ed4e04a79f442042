"""Oracle: Llama decoder forward with KV cache (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates `LlamaPostTokenization.forward` -> HF `LlamaForCausalLM.forward` as pinned by the reference
(transformers==4.31.0, pyproject.toml), following the in-tree copy of `LlamaModel.forward`
(procyon/model/pmc_llama.py:287-406: positions = arange(past, past+S), `_prepare_decoder_attention_mask`,
final norm appended to hidden_states) and the eager attention math shown at procyon/model/pmc_llama.py:221-247
(scores / sqrt(d) + additive mask, clamp at finfo.min, fp32 softmax, cast back, @ V; GQA via repeat_kv).
transformers itself is not vendored in the reference tree; the restatement is cross-checked against the installed
HF `LlamaForCausalLM` (eager attention) in tests/golden/make_golden.py.

RoPE base: transformers 4.31 predates `rope_theta`, so the reference ran Llama-3 with base 10000;
`rope_theta` is a parameter here, default 10000.0.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class LlamaCfg:
    d_model: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 8
    ffn_dim: int = 14336
    vocab: int = 128263  # 128256 + 8 added tokens - 1 ([EXT] row dropped, model_unified.py:166)
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_pos: int = 8192

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads


def _rounder(act_round: str):
    if act_round == "bf16":
        return lambda t: t.to(torch.bfloat16).to(torch.float32)
    if act_round == "none":
        return lambda t: t
    raise ValueError(act_round)


def rope_cos_sin(n_pos: int, head_dim: int, theta: float, table_dtype: torch.dtype = torch.float32):
    """HF LlamaRotaryEmbedding tables [n_pos, head_dim/2]: computed in fp32 at init, then cast with the module
    (`.bfloat16()` rounds the VALUES, positions stay exact) — table_dtype=bfloat16 reproduces that."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(n_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    return freqs.cos().to(table_dtype).float(), freqs.sin().to(table_dtype).float()


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, rnd) -> torch.Tensor:
    var = x.float().pow(2).mean(-1, keepdim=True)
    return rnd(w.float() * rnd(x.float() * torch.rsqrt(var + eps)))


def _rope(x, cos, sin, pos0):
    # x [B, H, S, hd]
    S = x.shape[-2]
    half = x.shape[-1] // 2
    c, s = cos[pos0 : pos0 + S], sin[pos0 : pos0 + S]
    x1, x2 = x[..., :half], x[..., half:]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


def llama_forward(
    sd: Dict[str, torch.Tensor],
    cfg: LlamaCfg,
    *,
    inputs_embeds: Optional[torch.Tensor] = None,
    input_ids: Optional[torch.Tensor] = None,
    attention_mask: Optional[torch.Tensor] = None,
    past: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None,
    labels: Optional[torch.Tensor] = None,
    act_round: str = "none",
    rope_table_dtype: torch.dtype = torch.float32,
    fused_swiglu_round: bool = True,
):
    """Returns dict(logits [B,S,V] fp32, hidden_states list of L+1 (last is post-norm), past, loss).

    `sd` uses HF names: model.embed_tokens.weight, model.layers.N.{self_attn.{q,k,v,o}_proj, mlp.{gate,up,down}_proj,
    input_layernorm, post_attention_layernorm}.weight, model.norm.weight, lm_head.weight.
    attention_mask: [B, past+S] (1 = attend) or None (= all ones, which is what the reference's decode steps pass).
    fused_swiglu_round: round silu(gate)*up once (what the CUDA epilogue stores) instead of after each op.
    """
    rnd = _rounder(act_round)
    g = lambda k: sd[k].float()
    if inputs_embeds is None:
        inputs_embeds = g("model.embed_tokens.weight")[input_ids]
    x = rnd(inputs_embeds.float())
    B, S, d = x.shape
    H, KVH, hd = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
    past_len = 0 if past is None else past[0][0].shape[2]
    ctx = past_len + S
    if attention_mask is None:
        attention_mask = torch.ones(B, ctx)
    neg = torch.finfo(torch.float32).min
    # _prepare_decoder_attention_mask: causal (S x ctx, offset by past_len) + padding, additive
    i = torch.arange(S)[:, None] + past_len
    j = torch.arange(ctx)[None, :]
    add_mask = torch.zeros(B, 1, S, ctx)
    if S > 1:
        add_mask = add_mask.masked_fill((j > i)[None, None], neg)
    add_mask = add_mask + (1.0 - attention_mask.float())[:, None, None, :] * neg
    add_mask = add_mask.clamp(min=neg)  # min + min overflows to -inf in HF too; clamp == their torch.max(...)
    cos, sin = rope_cos_sin(max(cfg.max_pos, ctx), hd, cfg.rope_theta, rope_table_dtype)
    hidden_states = []
    new_past = []
    for l in range(cfg.n_layers):
        p = f"model.layers.{l}."
        hidden_states.append(x)
        h = rmsnorm(x, sd[p + "input_layernorm.weight"], cfg.rms_eps, rnd)
        q = rnd(h @ g(p + "self_attn.q_proj.weight").t()).view(B, S, H, hd).transpose(1, 2)
        k = rnd(h @ g(p + "self_attn.k_proj.weight").t()).view(B, S, KVH, hd).transpose(1, 2)
        v = rnd(h @ g(p + "self_attn.v_proj.weight").t()).view(B, S, KVH, hd).transpose(1, 2)
        q, k = rnd(_rope(q, cos, sin, past_len)), rnd(_rope(k, cos, sin, past_len))
        if past is not None:
            k = torch.cat([past[l][0], k], dim=2)
            v = torch.cat([past[l][1], v], dim=2)
        new_past.append((k, v))
        kk = k.repeat_interleave(H // KVH, dim=1)
        vv = v.repeat_interleave(H // KVH, dim=1)
        s = q @ kk.transpose(-1, -2) / math.sqrt(hd) + add_mask
        s = torch.max(s, torch.tensor(neg))
        pr = rnd(torch.softmax(s, dim=-1, dtype=torch.float32))
        a = rnd((pr @ vv).transpose(1, 2).reshape(B, S, H * hd))
        x = rnd(x + a @ g(p + "self_attn.o_proj.weight").t())
        h = rmsnorm(x, sd[p + "post_attention_layernorm.weight"], cfg.rms_eps, rnd)
        gate = h @ g(p + "mlp.gate_proj.weight").t()
        up = h @ g(p + "mlp.up_proj.weight").t()
        if fused_swiglu_round:
            act = rnd(F.silu(gate) * up)
        else:
            act = rnd(rnd(F.silu(rnd(gate))) * rnd(up))
        x = rnd(x + act @ g(p + "mlp.down_proj.weight").t())
    x = rmsnorm(x, sd["model.norm.weight"], cfg.rms_eps, rnd)
    hidden_states.append(x)
    logits = x @ g("lm_head.weight").t()
    loss = None
    if labels is not None:
        shift_logits = logits[..., :-1, :].reshape(-1, logits.shape[-1])
        shift_labels = labels[..., 1:].reshape(-1)
        loss = F.cross_entropy(shift_logits, shift_labels, ignore_index=-100)
    return {"logits": logits, "hidden_states": hidden_states, "past": new_past, "loss": loss}


def random_llama_state_dict(cfg: LlamaCfg, seed: int = 0, std: float = 0.02, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    d, f, hd = cfg.d_model, cfg.ffn_dim, cfg.head_dim

    def w(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(dtype)

    sd = {"model.embed_tokens.weight": w(cfg.vocab, d, s=0.5)}
    for l in range(cfg.n_layers):
        p = f"model.layers.{l}."
        sd[p + "self_attn.q_proj.weight"] = w(cfg.n_heads * hd, d, s=1 / math.sqrt(d))
        sd[p + "self_attn.k_proj.weight"] = w(cfg.n_kv_heads * hd, d, s=1 / math.sqrt(d))
        sd[p + "self_attn.v_proj.weight"] = w(cfg.n_kv_heads * hd, d, s=1 / math.sqrt(d))
        sd[p + "self_attn.o_proj.weight"] = w(d, cfg.n_heads * hd, s=1 / math.sqrt(d))
        sd[p + "mlp.gate_proj.weight"] = w(f, d, s=1 / math.sqrt(d))
        sd[p + "mlp.up_proj.weight"] = w(f, d, s=1 / math.sqrt(d))
        sd[p + "mlp.down_proj.weight"] = w(d, f, s=1 / math.sqrt(f))
        sd[p + "input_layernorm.weight"] = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
        sd[p + "post_attention_layernorm.weight"] = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    sd["model.norm.weight"] = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    sd["lm_head.weight"] = w(cfg.vocab, d, s=1 / math.sqrt(d))
    return sd

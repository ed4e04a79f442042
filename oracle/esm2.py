"""Oracle: ESM2 encoder, ProteinPooler, long-sequence chunking (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates:
  * fair-esm 2.0.0 `ESM2.forward(tokens, repr_layers=[L])` as called at procyon/model/esm.py:526,536.
    fair-esm is an un-vendored dependency (pyproject.toml: fair-esm==2.0.0); the algorithm restated here is
    the published ESM2 architecture and is cross-checked against HF `EsmForMaskedLM` (modeling_esm.py) in
    tests/golden/make_golden.py.
  * `ProteinPooler.forward`            procyon/model/esm.py:154-217
  * `batched_split_long_seq`           procyon/training/train_utils.py:1497-1596
  * `reverse_batched_split`            procyon/training/train_utils.py:1599-1649
  * `ESM_PLM.forward`                  procyon/model/esm.py:504-558

`act_round` emulates the storage precision of the CUDA path: "bf16" rounds every tensor the kernels store
to bf16 (weights are assumed already bf16-representable), "none" is pure fp32.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

ESM_SIZES = {
    # name: (n_layers, d_model, n_heads)      procyon/model/esm.py:378-403
    "8m": (6, 320, 20),
    "35m": (12, 480, 20),
    "150m": (30, 640, 20),
    "650m": (33, 1280, 20),
    "3b": (36, 2560, 40),
    "15b": (48, 5120, 40),
}
# fair-esm Alphabet "ESM-1b": <cls>=0 <pad>=1 <eos>=2 <unk>=3, 20 aa + X B U Z O . - , <mask>=32
CLS_IDX, PAD_IDX, EOS_IDX, UNK_IDX, MASK_IDX, VOCAB = 0, 1, 2, 3, 32, 33


def _rounder(act_round: str):
    if act_round == "bf16":
        return lambda t: t.to(torch.bfloat16).to(torch.float32)
    if act_round == "none":
        return lambda t: t
    raise ValueError(act_round)


def rope_cos_sin(n_pos: int, head_dim: int, theta: float = 10000.0, table_dtype: torch.dtype = torch.float32):
    """cos/sin tables [n_pos, head_dim/2] as fair-esm RotaryEmbedding._update_cos_sin_tables builds them.

    table_dtype=torch.bfloat16 reproduces what the reference gets after `model.bfloat16()` (inv_freq buffer and
    the position vector are cast to bf16 before the outer product); float32 is the exact table.
    """
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    inv_freq = inv_freq.to(table_dtype)
    t = torch.arange(n_pos).to(table_dtype)
    freqs = torch.outer(t, inv_freq)
    return freqs.cos().float(), freqs.sin().float()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [..., T, head_dim]; rotate-half convention: x*cos + rotate_half(x)*sin."""
    half = x.shape[-1] // 2
    x1, x2 = x[..., :half], x[..., half:]
    c, s = cos[: x.shape[-2]], sin[: x.shape[-2]]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


def esm2_embed(tokens: torch.Tensor, table: torch.Tensor, token_dropout: bool = True, act_round: str = "none"):
    """fair-esm ESM2.forward embedding stage (mask-dropout rescale + pad zeroing)."""
    rnd = _rounder(act_round)
    padding_mask = tokens.eq(PAD_IDX)
    x = table[tokens]
    if token_dropout:
        x = x.masked_fill((tokens == MASK_IDX).unsqueeze(-1), 0.0)
        mask_ratio_train = 0.15 * 0.8
        src_lengths = (~padding_mask).sum(-1)
        # in the module dtype: n_mask.to(x.dtype) / src_lengths, then x * 0.88 (rounded), then / (1 - ratio)
        mask_ratio_observed = rnd(rnd((tokens == MASK_IDX).sum(-1).float()) / src_lengths)
        x = rnd(x * (1 - mask_ratio_train))
        x = rnd(x / rnd(1 - mask_ratio_observed)[:, None, None])
    x = x * (1 - padding_mask.unsqueeze(-1).to(x.dtype))
    return x


def esm2_forward(
    sd: Dict[str, torch.Tensor],
    tokens: torch.Tensor,
    n_layers: int,
    n_heads: int,
    *,
    token_dropout: bool = True,
    act_round: str = "none",
    rope_table_dtype: torch.dtype = torch.float32,
    ln_eps: float = 1e-5,
) -> torch.Tensor:
    """Returns representations[n_layers] (after emb_layer_norm_after), fp32 [B, T, d].

    `sd` uses fair-esm parameter names (the reference checkpoint layout):
    embed_tokens.weight, layers.N.self_attn.{q,k,v,out}_proj.{weight,bias}, layers.N.self_attn_layer_norm.*,
    layers.N.fc1.*, layers.N.fc2.*, layers.N.final_layer_norm.*, emb_layer_norm_after.*
    """
    rnd = _rounder(act_round)
    g = lambda k: sd[k].float()
    B, T = tokens.shape
    d = sd["embed_tokens.weight"].shape[1]
    hd = d // n_heads
    padding_mask = tokens.eq(PAD_IDX)
    x = rnd(esm2_embed(tokens, g("embed_tokens.weight"), token_dropout, act_round))
    cos, sin = rope_cos_sin(T, hd, 10000.0, rope_table_dtype)
    scaling = hd ** -0.5
    for l in range(n_layers):
        p = f"layers.{l}."
        h = rnd(F.layer_norm(x, (d,), g(p + "self_attn_layer_norm.weight"), g(p + "self_attn_layer_norm.bias"), ln_eps))
        q = rnd((h @ g(p + "self_attn.q_proj.weight").t() + g(p + "self_attn.q_proj.bias")) * scaling)
        k = rnd(h @ g(p + "self_attn.k_proj.weight").t() + g(p + "self_attn.k_proj.bias"))
        v = rnd(h @ g(p + "self_attn.v_proj.weight").t() + g(p + "self_attn.v_proj.bias"))
        q = q.view(B, T, n_heads, hd).transpose(1, 2)
        k = k.view(B, T, n_heads, hd).transpose(1, 2)
        v = v.view(B, T, n_heads, hd).transpose(1, 2)
        q, k = rnd(apply_rope(q, cos, sin)), rnd(apply_rope(k, cos, sin))
        s = q @ k.transpose(-1, -2)
        s = s.masked_fill(padding_mask[:, None, None, :], float("-inf"))
        pr = rnd(torch.softmax(s, dim=-1))  # fp32 softmax, probabilities stored in the module dtype
        a = rnd((pr @ v).transpose(1, 2).reshape(B, T, d))
        x = rnd(x + a @ g(p + "self_attn.out_proj.weight").t() + g(p + "self_attn.out_proj.bias"))
        h = rnd(F.layer_norm(x, (d,), g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias"), ln_eps))
        h = rnd(F.gelu(h @ g(p + "fc1.weight").t() + g(p + "fc1.bias")))
        x = rnd(x + h @ g(p + "fc2.weight").t() + g(p + "fc2.bias"))
    x = rnd(F.layer_norm(x, (d,), g("emb_layer_norm_after.weight"), g("emb_layer_norm_after.bias"), ln_eps))
    return x


def esm2_lm_head(sd: Dict[str, torch.Tensor], z: torch.Tensor, *, act_round: str = "none",
                 ln_eps: float = 1e-5) -> torch.Tensor:
    """fair-esm 2.0.0 `RobertaLMHead` (esm/modules.py; absent from /root/reference, pinned through HF EsmLMHead):
    logits = layer_norm(gelu(dense(z))) @ weight^T + bias with `weight` tied to embed_tokens.weight.  The reference
    returns them from ESM_PLM.forward (procyon/model/esm.py:549-555) for the `return_mlm` path
    (procyon/model/model_unified.py:505-509)."""
    rnd = _rounder(act_round)
    g = lambda k: sd[k].float()
    d = z.shape[-1]
    w = g("lm_head.weight") if "lm_head.weight" in sd else g("embed_tokens.weight")
    h = rnd(F.gelu(z @ g("lm_head.dense.weight").t() + g("lm_head.dense.bias")))
    h = rnd(F.layer_norm(h, (d,), g("lm_head.layer_norm.weight"), g("lm_head.layer_norm.bias"), ln_eps))
    return h @ w.t() + g("lm_head.bias")


def batched_split_long_seq(toks: torch.Tensor, padding_idx: int = PAD_IDX, eos_idx: int = EOS_IDX,
                           max_protein_len: int = 1024):
    """'split' strategy of procyon/training/train_utils.py:1497-1596 (does NOT mutate its input).

    Returns (new_toks [B', <= max_len+2], batch_keys [B'] int64, eos_loc list[int]).
    """
    toks = toks.clone()
    B, W = toks.shape
    cls_idx = int(toks[0, 0])
    eos_loc = []
    for i in range(B):
        nz = (toks[i] == eos_idx).nonzero(as_tuple=True)[0]
        if nz.numel() != 1:
            raise ValueError("each row must contain exactly one EOS token")  # reference: ambiguous bool() error
        eos_loc.append(int(nz[0]))
    batch_keys = list(range(B))
    extra = []
    for i in range(B):
        if eos_loc[i] <= max_protein_len + 1:
            continue
        overage = eos_loc[i]
        num_add = overage // (max_protein_len + 1)
        for j in range(num_add):
            bot = (j + 1) * max_protein_len + 1
            new = torch.full((W,), padding_idx, dtype=toks.dtype)
            tail = toks[i, bot:]
            # reference: new_empty[0, 1:(n+1)] = new_tmp — n = W - bot always fits because bot >= 1
            new[1 : tail.shape[0] + 1] = tail
            new[0] = cls_idx
            if j < num_add - 1:
                new[max_protein_len + 1] = eos_idx
                new[max_protein_len + 2 :] = 1  # literal 1 in the reference (train_utils.py:1557)
            extra.append(new)
            batch_keys.append(i)
        toks[i, max_protein_len + 2 :] = padding_idx
        toks[i, max_protein_len + 1] = eos_idx
    new_toks = torch.cat([toks] + [e.unsqueeze(0) for e in extra], dim=0) if extra else toks
    new_toks = new_toks[:, : max_protein_len + 2]
    return new_toks, torch.tensor(batch_keys, dtype=torch.int64), eos_loc


def protein_pooler(z: torch.Tensor, batch_keys: torch.Tensor, padding_mask: torch.Tensor, method: str = "mean",
                   correction: bool = False) -> torch.Tensor:
    """ProteinPooler.forward, procyon/model/esm.py:154-217 (z [B', T, d] fp32)."""
    z = z.clone()
    if method == "max":
        z[padding_mask] = -float("inf")
    out = []
    for i in range(int(batch_keys.max()) + 1):
        sel = batch_keys == i
        if sel.sum() == 0:
            continue
        rows = z[sel].reshape(-1, z.shape[-1])
        if method == "mean":
            rows = rows[~padding_mask[sel].reshape(-1)]
            if correction:
                rows = rows[1:-1]
            out.append(rows.nanmean(dim=-2))
        elif method == "max":
            out.append(rows.max(dim=-2)[0])
        else:
            raise NotImplementedError(method)
    return torch.stack(out)


def reverse_batched_split(z: torch.Tensor, batch_keys: torch.Tensor, eos_locs: List[int]) -> torch.Tensor:
    """procyon/training/train_utils.py:1599-1649: stitch chunk rows back into per-protein token sequences."""
    out = []
    for i in range(int(batch_keys.max()) + 1):
        idx = (batch_keys == i).nonzero(as_tuple=True)[0].sort()[0]
        if idx.numel() == 0:
            continue
        cp = z[idx]
        keep_eos = torch.ones(cp.shape[0], cp.shape[1], dtype=torch.bool)
        keep_eos[:-1, -1] = False
        keep_cls = torch.ones(cp.shape[0], cp.shape[1], dtype=torch.bool)
        keep_cls[1:, 0] = False
        cp = cp.reshape(-1, z.shape[-1])[(keep_eos & keep_cls).flatten()]
        out.append(cp)
    max_size = max(eos_locs) + 1
    for i in range(len(out)):
        diff = max_size - out[i].shape[0]
        if diff > 0:
            out[i] = torch.cat([out[i], torch.zeros(diff, z.shape[-1])], dim=0)
        elif diff < 0:
            out[i] = out[i][:max_size]
    return torch.stack(out)


def esm_plm_forward(sd, tokens, n_layers, n_heads, *, pooling: str = "mean", correction: bool = False,
                    max_protein_len: int = 1024, aggregate: bool = True, **kw) -> torch.Tensor:
    """ESM_PLM.forward (procyon/model/esm.py:504-558) without the LM head: chunk -> encode -> pool."""
    new_toks, keys, eos_loc = batched_split_long_seq(tokens, max_protein_len=max_protein_len)
    z = esm2_forward(sd, new_toks, n_layers, n_heads, **kw)
    if aggregate:
        return protein_pooler(z, keys, new_toks == PAD_IDX, pooling, correction)
    return reverse_batched_split(z, keys, eos_loc)


def random_esm_state_dict(n_layers: int, d: int, ffn: Optional[int] = None, seed: int = 0, std: float = 0.02,
                          dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights in the fair-esm layout (N(0, 0.02), LN weight 1, biases small)."""
    g = torch.Generator().manual_seed(seed)
    ffn = ffn or 4 * d
    sd = {}

    def w(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(dtype)

    sd["embed_tokens.weight"] = w(VOCAB, d, s=0.5)
    sd["embed_tokens.weight"][PAD_IDX] = 0
    for l in range(n_layers):
        p = f"layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{nm}.weight"] = w(d, d, s=1.0 / math.sqrt(d))
            sd[p + f"self_attn.{nm}.bias"] = w(d, s=0.02)
        sd[p + "self_attn_layer_norm.weight"] = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
        sd[p + "self_attn_layer_norm.bias"] = w(d, s=0.05)
        sd[p + "final_layer_norm.weight"] = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
        sd[p + "final_layer_norm.bias"] = w(d, s=0.05)
        sd[p + "fc1.weight"] = w(ffn, d, s=1.0 / math.sqrt(d))
        sd[p + "fc1.bias"] = w(ffn, s=0.02)
        sd[p + "fc2.weight"] = w(d, ffn, s=1.0 / math.sqrt(ffn))
        sd[p + "fc2.bias"] = w(d, s=0.02)
    sd["emb_layer_norm_after.weight"] = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    sd["emb_layer_norm_after.bias"] = w(d, s=0.05)
    # masked-LM head (drawn after everything else, so the encoder weights of a given seed are unchanged)
    sd["lm_head.dense.weight"] = w(d, d, s=1.0 / math.sqrt(d))
    sd["lm_head.dense.bias"] = w(d, s=0.02)
    sd["lm_head.layer_norm.weight"] = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    sd["lm_head.layer_norm.bias"] = w(d, s=0.05)
    sd["lm_head.weight"] = sd["embed_tokens.weight"]  # tied
    sd["lm_head.bias"] = w(VOCAB, s=0.1)
    return sd


def random_protein_tokens(n: int, length: int, seed: int = 1234, lengths: Optional[List[int]] = None) -> torch.Tensor:
    """Synthetic proteins: uniform over the 20 standard residues (ids 4..23) + CLS/EOS, right-padded."""
    g = torch.Generator().manual_seed(seed)
    lengths = lengths or [length] * n
    W = max(lengths) + 2
    toks = torch.full((n, W), PAD_IDX, dtype=torch.int64)
    for i, L in enumerate(lengths):
        toks[i, 0] = CLS_IDX
        toks[i, 1 : L + 1] = torch.randint(4, 24, (L,), generator=g)
        toks[i, L + 1] = EOS_IDX
    return toks

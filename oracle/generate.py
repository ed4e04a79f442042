"""Oracle: text generation loops of UnifiedProCyon (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates, on top of oracle.llama.llama_forward:
  * `_generate_beam_search`   procyon/model/model_unified.py:701-842 (diverse beam search, Hamming penalty,
                              in-place penalty on the log-prob view, beam reorder of out / log-probs / logits / KV)
  * `_generate_sampling`      procyon/model/model_unified.py:860-921 (greedy branch; intended signature — the
                              reference call site at :998-1005 drops `attn_masks`)
Both keep the reference quirk that decode steps pass NO attention mask (all positions visible), which only
matters for left-padded batches; `mask_pads_in_decode=True` switches to the corrected behaviour the CUDA path
implements (pads stay masked).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .llama import LlamaCfg, llama_forward


def _step(sd, cfg, i, embeds, attn_mask, out, past, mask_pads_in_decode, **kw):
    if i == 0:
        r = llama_forward(sd, cfg, inputs_embeds=embeds, attention_mask=attn_mask, **kw)
    else:
        am = None
        if mask_pads_in_decode and attn_mask is not None:
            am = torch.cat([attn_mask.float(), torch.ones(attn_mask.shape[0], i)], dim=1)
        r = llama_forward(sd, cfg, input_ids=out[:, i - 1].unsqueeze(-1), past=past, attention_mask=am, **kw)
    return r["logits"][:, -1, :], r["past"]


def beam_select_step(log_probs, out, cur, i, n, beam_size, beam_group_size, diversity_penalty, margins=None):
    """One step of the selection of `_generate_beam_search` (procyon/model/model_unified.py:783-828) on
    `log_probs` [n*beams, V] (= LogSoftmax(logits) + running scores; penalised IN PLACE like the reference's view).
    Updates `out` (token histories, reordered) and `cur` (scores) in place and returns `parents` [n*beams]: the flat
    row every new beam extends, i.e. the gather index the reference applies group by group to the logits history
    and to every layer's K / V (:827-832).  Group-by-group in-place gathers compose to one gather here because a
    group only ever reads rows of its own group (`orig_candidate_idxs` in [group_start, check_end))."""
    V = log_probs.shape[1]
    groups = beam_size // beam_group_size
    parents = torch.arange(n * beam_size)
    for inp in range(n):
        b0 = inp * beam_size
        for g in range(groups):
            inc = 1 if i == 0 else beam_group_size
            gs = b0 + g * beam_group_size
            ge = gs + beam_group_size
            lp = log_probs[gs : gs + inc]
            if g != 0:
                prev = out[b0:gs, i]
                lp -= diversity_penalty * torch.bincount(prev, minlength=V)  # in place on the view
            vals, idx = lp.ravel().topk(beam_group_size)
            if margins is not None:
                top = lp.ravel().topk(min(beam_group_size + 1, lp.numel())).values
                margins.extend((top[:-1] - top[1:]).tolist())
            toks = idx % V
            src = (idx // V) + gs
            out[gs:ge] = out[src]
            out[torch.arange(gs, ge), i] = toks
            cur[gs:ge] = vals
            parents[gs:ge] = src
    return parents


@torch.no_grad()
def generate_beam_search(sd, cfg: LlamaCfg, input_embeds, attn_mask, *, max_len=64, beam_size=5, beam_group_size=5,
                         diversity_penalty=0.8, eos_id: int = 2, mask_pads_in_decode: bool = False,
                         margins: Optional[List[float]] = None, trace: Optional[list] = None, **kw):
    """Returns (out [n, beams, max_len] int64, log_probs [n, beams], logits [n, beams, steps, V]).

    `trace` (test aid): if a list is passed, every step appends dict(logits [n*beams, V] as the model produced them
    for the beam rows BEFORE the selection, parents [n*beams], tokens [n*beams], scores [n*beams], margin = the
    smallest decision gap of the step).

    `margins` (test aid, not in the reference): if a list is passed, every group selection appends the gaps between
    consecutive entries of its top-(k+1) candidate scores — the k-1 gaps that fix the ORDER of the selected beams and
    the gap to the first rejected candidate. Their minimum is the numerical head-room of the whole search: an
    implementation whose scores deviate by less than half of it must reproduce every beam exactly."""
    n = input_embeds.shape[0]
    bb = n * beam_size
    V = cfg.vocab
    if beam_size % beam_group_size != 0:
        raise ValueError(f"beam_group_size must evenly divide beam_size, got: {beam_size} % {beam_group_size} != 0")
    groups = beam_size // beam_group_size
    embeds = torch.repeat_interleave(input_embeds, beam_size, dim=0)
    mask = torch.repeat_interleave(attn_mask, beam_size, dim=0) if attn_mask is not None else None
    cur = torch.zeros(bb)
    out = torch.zeros(bb, max_len, dtype=torch.int64)
    past = None
    output_logits = None
    for i in range(max_len):
        logits, past = _step(sd, cfg, i, embeds, mask, out, past, mask_pads_in_decode, **kw)
        past = [[k.clone(), v.clone()] for k, v in past]
        it = logits.clone().unsqueeze(1)
        output_logits = it if output_logits is None else torch.cat([output_logits, it], dim=1)
        log_probs = torch.log_softmax(logits, dim=-1) + cur[:, None]
        step_margins = [] if (trace is not None or margins is not None) else None
        parents = beam_select_step(log_probs, out, cur, i, n, beam_size, beam_group_size, diversity_penalty,
                                   step_margins)
        if margins is not None:
            margins.extend(step_margins)
        if trace is not None:
            trace.append(dict(logits=logits.clone(), parents=parents.clone(), tokens=out[:, i].clone(),
                              scores=cur.clone(), margin=min(step_margins)))
        output_logits = output_logits[parents]
        for l in range(len(past)):
            past[l][0] = past[l][0][parents]
            past[l][1] = past[l][1][parents]
        if torch.all((out == eos_id).any(dim=1)).item():
            break
    return (out.unflatten(0, (n, beam_size)), cur.unflatten(0, (n, beam_size)),
            output_logits.unflatten(0, (n, beam_size)))


@torch.no_grad()
def generate_greedy(sd, cfg: LlamaCfg, input_embeds, attn_mask, *, max_len=64, mask_pads_in_decode: bool = False, **kw):
    """Greedy branch of _generate_sampling. Returns (out [n, max_len], total_log_prob [n], logits [n, steps, V])."""
    n = input_embeds.shape[0]
    out = None
    past = None
    logits_all = []
    total = torch.zeros(n)
    for i in range(max_len):
        logits, past = _step(sd, cfg, i, input_embeds, attn_mask, out, past, mask_pads_in_decode, **kw)
        logits_all.append(logits.clone())
        lp = torch.log_softmax(logits, dim=-1)
        nxt = torch.argmax(logits, dim=-1, keepdim=True)
        total += lp[torch.arange(n), nxt.squeeze(-1)]
        out = nxt if out is None else torch.cat([out, nxt], dim=-1)
    return out, total, torch.stack(logits_all, 1)


def structured_lm_head(embed: torch.Tensor, seed: int, J: int = 12) -> torch.Tensor:
    """Test aid (not in the reference): a margin-controlled LM head for beam-search parity tests.  Random head rows
    give ~N(0,1) logits whose top-k gaps (~0.3) are only ~10x the bf16 forward noise, so some of the ~100 decisions of
    a beam search always fall inside the noise.  Here token v gets J graded successor tokens,
    lm_head[succ_j(v)] += c_j(v) * embed[v] / |embed[v]| with c_j in [0.55, 1]: with embeddings that dominate the
    residual stream the candidates that matter are ~10 logits above the rest and ~0.5 apart, while the forward noise
    stays ~3e-2 — decisions become numerically unambiguous, yet every kernel still runs on dense random data."""
    g = torch.Generator().manual_seed(seed)
    vocab, d = embed.shape
    E = embed.float()
    E = E / E.norm(dim=1, keepdim=True).clamp_min(1e-6)
    head = torch.zeros(vocab, d)
    for j in range(J):
        succ = torch.randperm(vocab, generator=g)
        c = 0.55 + 0.45 * torch.rand(vocab, generator=g)
        head.index_add_(0, succ, c[:, None] * E)
    return head.to(torch.bfloat16)

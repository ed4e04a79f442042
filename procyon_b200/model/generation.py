"""Device-side generation loops: prefill once per input, then CUDA-graph replays of (decode forward, select).

Mirrors UnifiedProCyon._generate_beam_search (procyon/model/model_unified.py:701-842) and the greedy branch of
_generate_sampling (:860-921). Differences from the reference loop, none of which change the returned values:
  * the prompt is prefilled once per input instead of beam_size identical times (:751-752);
  * the KV cache is never reordered (:830-832): beams index it through an ancestry table;
  * logits stay on the GPU (no per-step `.cpu()`, :773); the per-beam logits history the reference builds by
    re-indexing a growing CPU tensor every step (:827) is gathered once at the end from the same table;
  * left-pad positions stay masked during decode steps (the reference passes no mask after step 0, :769 — a
    defect for padded batches; identical for un-padded prompts).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .pmc_llama import SELECT_BEAM, SELECT_GREEDY, DecodeSession, LlamaPostTokenization

# kernels launched through CUDA-graph replays (the library's own counter only sees direct launches)
GRAPH_REPLAY_LAUNCHES = 0


def _run(text_encoder: LlamaPostTokenization, input_embeds, attn_mask, max_len, beams, mode, group, penalty, eos_id,
         stop_on_all_eos, return_logits, use_graph=True):
    n, S, _ = input_embeds.shape
    dev = input_embeds.device
    sel = torch.arange(n, device=dev, dtype=torch.int32) * S + (S - 1)
    sess = text_encoder.get_session(n, beams, S, max_len, dev, attn_mask is not None, return_logits)
    _, _, prefill_logits, valid = text_encoder.prefill(input_embeds, attn_mask, want_cache=True, want_hidden=False,
                                                        sel_rows=sel, kv_out=sess.kv_prompt)
    if attn_mask is not None:
        sess.prompt_valid.copy_(valid)
    sess.reset(prefill_logits)
    sess.select(mode, group, penalty, eos_id, stop_on_all_eos)  # step 0: tokens from the prefill logits
    steps_left = max_len - 1
    if steps_left > 0:
        if use_graph:
            # first (forward, select) runs eagerly: it loads every kernel before stream capture
            sess.forward()
            sess.select(mode, group, penalty, eos_id, stop_on_all_eos)
            steps_left -= 1
            if steps_left > 0:
                g = sess.step_graph(mode, group, penalty, eos_id, stop_on_all_eos)
                done = 0
                while done < steps_left:
                    burst = min(16, steps_left - done) if stop_on_all_eos else steps_left - done
                    for _ in range(burst):
                        g.replay()
                    global GRAPH_REPLAY_LAUNCHES
                    GRAPH_REPLAY_LAUNCHES += burst * sess.graph_launches
                    done += burst
                    if stop_on_all_eos and done < steps_left and int(sess.state[2].item()) != 0:
                        break
        else:
            for _ in range(steps_left):
                sess.forward()
                sess.select(mode, group, penalty, eos_id, stop_on_all_eos)
    return sess


def _collect(sess: DecodeSession, max_len: int, return_logits: bool):
    state = sess.state.cpu()
    finished, finish_step, t = int(state[2]), int(state[3]), int(state[0])
    steps = finish_step + 1 if finished else t
    tokens = sess.tokens.to(torch.int64)
    out = torch.zeros((sess.rows, max_len), dtype=torch.int64, device=tokens.device)
    out[:, :steps] = tokens[:, :steps]
    logits = None
    if return_logits:
        # logits of step s for beam row b were produced by physical row: s == 0 -> any row of the input (all equal),
        # s >= 1 -> slots[b][s-1] (the row that held this beam's history when step s ran)
        rows = torch.arange(sess.rows, device=tokens.device)
        src = torch.empty((sess.rows, steps), dtype=torch.int64, device=tokens.device)
        src[:, 0] = (rows // sess.beams) * sess.beams
        if steps > 1:
            src[:, 1:] = sess.slots[:, : steps - 1].to(torch.int64)
        step_idx = torch.arange(steps, device=tokens.device)[None, :].expand(sess.rows, steps)
        logits = sess.logits_hist[step_idx, src]  # [rows, steps, V]
    return out, sess.logprobs.clone(), logits, steps


@torch.no_grad()
def generate_beam_search(text_encoder, input_embeds, attn_mask, max_len=64, beam_size=5, beam_group_size=5,
                         diversity_penalty=0.8, eos_token_id: int = -1, return_logits: bool = True,
                         use_graph: bool = True):
    """Returns (out [n, beams, max_len] int64 cpu, log_probs [n, beams] cpu, logits [n, beams, steps, V] | None)."""
    if beam_size % beam_group_size != 0:
        raise ValueError("beam_group_size must evenly divide beam_size, got: "
                         f"{beam_size} % {beam_group_size} != 0")
    n = input_embeds.shape[0]
    sess = _run(text_encoder, input_embeds, attn_mask, max_len, beam_size, SELECT_BEAM, beam_group_size,
                diversity_penalty, eos_token_id, True, return_logits, use_graph)
    out, lp, logits, steps = _collect(sess, max_len, return_logits)
    out = out.cpu().unflatten(0, (n, beam_size))
    lp = lp.cpu().unflatten(0, (n, beam_size))
    if logits is not None:
        logits = logits.unflatten(0, (n, beam_size))
    return out, lp, logits


@torch.no_grad()
def generate_greedy(text_encoder, input_embeds, attn_mask, max_len=64, return_logits: bool = True,
                    use_graph: bool = True):
    """Returns (out [n, max_len] int64 cpu, total_log_prob [n] cpu, logits [n, steps, V] | None)."""
    sess = _run(text_encoder, input_embeds, attn_mask, max_len, 1, SELECT_GREEDY, 1, 0.0, -1, False, return_logits,
                use_graph)
    out, lp, logits, _ = _collect(sess, max_len, return_logits)
    return out.cpu(), lp.cpu(), logits

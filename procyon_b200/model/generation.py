"""Device-side generation loops: prefill once per input, then CUDA-graph replays of (decode forward, select).

Mirrors UnifiedProCyon._generate_beam_search (procyon/model/model_unified.py:701-842) and the greedy branch of
_generate_sampling (:860-921). Differences from the reference loop, none of which change the returned values
(batches of more than 16 beam rows are split over several sessions that step in lock-step and share the all-EOS stop,
`_run_group`, so the stop step is the whole batch's like the reference's):
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


def _run_group(text_encoder: LlamaPostTokenization, input_embeds, attn_mask, max_len, beams, group, penalty, eos_id,
               return_logits, use_graph=True):
    """Beam search over a batch of more than 16 beam rows: the inputs are split over sessions of <= 16 rows that run
    step by step in lock-step on one stream; a device-side group state accumulates, per step, the number of inputs whose
    beams all hold an EOS, and the last session of the step stops ALL of them when that is every input — the reference
    stops (only) then, and inputs that finished earlier keep extending and re-ranking their beams until then
    (procyon/model/model_unified.py:833).  Returns (sessions, group_state)."""
    n, S, _ = input_embeds.shape
    dev = input_embeds.device
    per = max(1, 16 // beams)
    gstate = torch.zeros(4, device=dev, dtype=torch.int32)
    gstate[3] = n
    sessions = []
    for i0 in range(0, n, per):
        sl = slice(i0, min(n, i0 + per))
        k = sl.stop - sl.start
        sess = text_encoder.new_session(k, beams, S, max_len, dev, attn_mask is not None, return_logits)
        sel = torch.arange(k, device=dev, dtype=torch.int32) * S + (S - 1)
        _, _, logits, valid = text_encoder.prefill(input_embeds[sl], attn_mask[sl] if attn_mask is not None else None,
                                                   want_cache=True, want_hidden=False, sel_rows=sel,
                                                   kv_out=sess.kv_prompt)
        if attn_mask is not None:
            sess.prompt_valid.copy_(valid)
        sess.reset(logits)
        sessions.append(sess)
    last = len(sessions) - 1

    def select_all():
        for k, s in enumerate(sessions):
            s.select(SELECT_BEAM, group, penalty, eos_id, True, gstate, k == last)

    def step_all():
        for k, s in enumerate(sessions):
            s.forward()
            s.select(SELECT_BEAM, group, penalty, eos_id, True, gstate, k == last)

    select_all()  # step 0: tokens from the prefill logits
    steps_left = max_len - 1
    if steps_left > 0:
        step_all()  # eager once: loads every kernel before stream capture
        steps_left -= 1
    if steps_left > 0 and use_graph:
        from .. import _lib

        lib = _lib.load()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        n0 = lib.pcy_launch_count()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                step_all()
        torch.cuda.current_stream(dev).wait_stream(side)
        per_replay = lib.pcy_launch_count() - n0
    done = 0
    while done < steps_left:
        burst = min(16, steps_left - done)
        for _ in range(burst):
            if use_graph:
                g.replay()
            else:
                step_all()
        if use_graph:
            global GRAPH_REPLAY_LAUNCHES
            GRAPH_REPLAY_LAUNCHES += burst * per_replay
        done += burst
        if done < steps_left and int(gstate[1].item()) != 0:
            break
    return sessions, gstate


def _collect(sess: DecodeSession, max_len: int, return_logits: bool, group_state=None):
    state = sess.state.cpu()
    finished, finish_step, t = int(state[2]), int(state[3]), int(state[0])
    if group_state is not None:
        gs = group_state.cpu()
        finished, finish_step = int(gs[1]), int(gs[2])
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
        # the reference returns the logits history on the HOST (it is built there, model_unified.py:773-781): one gather
        # on the device, one copy into page-locked memory (657 MB at 10 beams x 128 steps x V = 128263: ~15 ms over PCIe
        # instead of ~0.4 s into pageable memory; torch's caching host allocator reuses the block across calls)
        dev_logits = sess.logits_hist[step_idx, src]  # [rows, steps, V]
        logits = torch.empty(dev_logits.shape, dtype=dev_logits.dtype, pin_memory=True)
        logits.copy_(dev_logits, non_blocking=True)
        torch.cuda.current_stream(dev_logits.device).synchronize()
    return out, sess.logprobs.clone(), logits, steps


@torch.no_grad()
def generate_beam_search(text_encoder, input_embeds, attn_mask, max_len=64, beam_size=5, beam_group_size=5,
                         diversity_penalty=0.8, eos_token_id: int = -1, return_logits: bool = True,
                         use_graph: bool = True):
    """Returns (out [n, beams, max_len] int64 cpu, log_probs [n, beams] cpu, logits [n, beams, steps, V] fp32 cpu |
    None).  Any batch size: more than 16 beam rows are split over lock-step sessions sharing the stop condition."""
    if beam_size % beam_group_size != 0:
        raise ValueError("beam_group_size must evenly divide beam_size, got: "
                         f"{beam_size} % {beam_group_size} != 0")
    n = input_embeds.shape[0]
    if n * beam_size > 16 and n > 1:
        sessions, gstate = _run_group(text_encoder, input_embeds, attn_mask, max_len, beam_size, beam_group_size,
                                      diversity_penalty, eos_token_id, return_logits, use_graph)
        parts = [_collect(s, max_len, return_logits, gstate) for s in sessions]
        out = torch.cat([p[0] for p in parts], 0).cpu().unflatten(0, (n, beam_size))
        lp = torch.cat([p[1] for p in parts], 0).cpu().unflatten(0, (n, beam_size))
        logits = torch.cat([p[2] for p in parts], 0).unflatten(0, (n, beam_size)) if return_logits else None
        return out, lp, logits
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

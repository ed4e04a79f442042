"""Diverse beam search on the GPU against the oracle, with NO slack on token ids.

north_star: "exactly for argmax token ids".  Beam search is a chain of top-k decisions on log-probabilities that the
bf16 forward pass reproduces to ~3e-2 (measured on B200: max |log-prob error| 0.037 over 1003-way rows of the tiny
models, scripts/probe_logit_noise.py), so exactness is proven in three parts that together leave no room:

  (A) the SELECTION (LogSoftmax, Hamming penalty per group, ravel().topk, reorder of token histories / scores / KV
      ancestry, all-EOS stop) is bit-exact: `pcy_decode_select` is driven with the same fp32 logits as the oracle's
      `beam_select_step` for many steps — every token id, parent and physical KV row must be identical, at V = 1003
      and at the real V = 128263 with the evaluation default 10 beams in groups of 2;
  (B) the FORWARD under beam ancestry is right at every step: the oracle's own decisions (tokens + parents) are forced
      into the device state step by step and the logits of every beam row are compared with the oracle's — KV slots
      are never reordered on the device, so this is the test of the ancestry indirection and of the attention kernel
      that serves all beams of an input from one prompt copy;
  (C) whole generations through the public loop (CUDA-graph replays and eager): 100 % of the beams equal the oracle's,
      token for token, for prompts whose smallest decision margin in the oracle (over the WHOLE generation) exceeds
      MARGIN = twice the worst measured forward deviation; the seeds come from the committed search
      scripts/find_beam_seeds.py and the test re-asserts their margin, so a stale seed fails instead of passing;
  (D) EVERY decision of a device-run search — also with 10 and 16 beams, where ~150 decisions make a margin-free
      prompt impossible to find — is verified against the oracle evaluated on the device's own beam histories: the
      device's selection must be the oracle's, except between candidates the oracle itself scores closer than MARGIN
      (then it must still be a top-k within MARGIN, in an order consistent within MARGIN, without duplicates).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.generate import structured_lm_head  # noqa: E402

EMBED_GAIN = 4.0
MARGIN = 0.12  # > 2 x the measured worst log-prob deviation of the bf16 forward on the top candidates (0.056)


def _tiny_state(kind="gq4", max_pos=1024, vocab=1003, seed=3, structured_head=True):
    """Seeded tiny Llama (2 layers, head_dim 128, GQA).  `structured_head`: margin-controlled LM head — instead of
    random rows (whose 1003 logits are ~N(0,1) with top-k gaps of ~0.3, i.e. only ~10x the bf16 forward noise, so that
    some of the ~100 decisions of a beam search always land inside the noise), every token v gets J = 12 graded
    successor tokens: lm_head[succ_j(v)] += c_j(v) * embed[v] / |embed[v]|, c_j in [0.55, 1].  The residual stream
    keeps a large component along the last token's embedding, so the candidates that matter have logits ~10 apart by
    ~1 while the forward noise stays ~1e-2: decisions are numerically unambiguous, yet every kernel still runs on
    generic random data (attention, MLP and the head's cross-talk terms are all dense random)."""
    from oracle.llama import LlamaCfg, random_llama_state_dict

    H, d, f = {"gq2": (4, 512, 1024), "gq4": (8, 1024, 1024), "sel": (2, 256, 256)}[kind]
    oc = LlamaCfg(d_model=d, n_layers=2 if kind != "sel" else 1, n_heads=H, n_kv_heads=2 if kind != "sel" else 1,
                  ffn_dim=f, vocab=vocab, max_pos=max_pos)
    sd = random_llama_state_dict(oc, seed=seed)
    if structured_head and kind != "sel":
        # embeddings 4x larger than the sub-layer outputs: the last token's embedding dominates the residual stream
        sd["model.embed_tokens.weight"] = (sd["model.embed_tokens.weight"].float() * EMBED_GAIN).to(torch.bfloat16)
        sd["lm_head.weight"] = structured_lm_head(sd["model.embed_tokens.weight"], seed + 1000)
    return oc, sd


def _tiny(kind="gq4", max_pos=1024, vocab=1003, seed=3, structured_head=True):
    from procyon_b200.model.pmc_llama import LlamaConfig, LlamaPostTokenization

    oc, sd = _tiny_state(kind, max_pos, vocab, seed, structured_head)
    pc = LlamaConfig(hidden_size=oc.d_model, intermediate_size=oc.ffn_dim, num_hidden_layers=oc.n_layers,
                     num_attention_heads=oc.n_heads, num_key_value_heads=oc.n_kv_heads, vocab_size=vocab,
                     max_position_embeddings=max_pos)
    m = LlamaPostTokenization(config=pc, dtype=torch.bfloat16)
    m.model.load_state_dict(sd, strict=True)
    return oc, sd, m.cuda()


def _inputs(oc, sd, B, S, seed, pad_left=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, oc.vocab, (B, S), generator=g)
    emb = sd["model.embed_tokens.weight"][ids].clone()
    emb[:, S // 2] = (torch.randn(B, oc.d_model, generator=g) * 0.5).bfloat16()
    mask = torch.ones(B, S)
    if pad_left and B > 1:
        mask[1, :pad_left] = 0
    return ids, emb, mask


# ------------------------------------------------------------------------------------------------ (A) selection
@pytest.mark.parametrize("V,n,beams,group,steps,eos_boost", [
    (1003, 2, 4, 2, 12, 0.0), (1003, 1, 6, 1, 10, 0.0), (1003, 3, 5, 5, 10, 0.0), (1003, 1, 16, 4, 8, 0.0),
    (1003, 2, 8, 8, 9, 0.0), (128263, 1, 10, 2, 16, 0.0), (128263, 1, 10, 5, 6, 0.0), (1003, 2, 4, 2, 12, 9.0),
    (1003, 1, 2, 1, 12, 7.0),
])
def test_selection_is_bit_exact(cuda_device, V, n, beams, group, steps, eos_boost):
    """`pcy_decode_select` vs the oracle's `beam_select_step` on identical fp32 logits, step after step: tokens,
    parents (through the physical-KV-row table) and the all-EOS stop must be IDENTICAL, scores equal to fp32 rounding.
    eos_boost > 0 raises the EOS logit so that beams finish at different steps and the early stop triggers."""
    from oracle.generate import beam_select_step
    from procyon_b200.model.pmc_llama import SELECT_BEAM

    oc, sd, m = _tiny("sel", vocab=V, seed=1)
    dev = torch.device("cuda")
    eos = 17
    sess = m.get_session(n, beams, 8, steps, dev, False, False)
    sess.reset(None)
    bb = n * beams
    g = torch.Generator().manual_seed(V + 31 * beams + group)
    out = torch.zeros(bb, steps, dtype=torch.int64)
    cur = torch.zeros(bb)
    phys = torch.zeros(bb, steps, dtype=torch.int64)
    stopped_at = None
    for i in range(steps):
        logits = torch.randn(bb, V, generator=g) * 3.0
        if eos_boost and i >= 2:  # more and more rows are pushed towards EOS: beams finish at different steps
            logits[:, eos] += 3.0 * eos_boost * (torch.rand(bb, generator=g) < 0.25 * (i - 1)).float()
        if i == 0:  # step 0: every beam row of an input holds the same prefill logits
            logits = logits.view(n, beams, V)[:, :1].expand(n, beams, V).reshape(bb, V).contiguous()
        sess.logits_cur.copy_(logits)
        sess.select(SELECT_BEAM, group, 0.8, eos, True)
        lp = torch.log_softmax(logits, dim=-1) + cur[:, None]
        margins = []
        parents = beam_select_step(lp, out, cur, i, n, beams, group, 0.8, margins)
        assert min(margins) > 1e-4, "test logits produced a numerical tie; pick another seed"
        new_phys = phys[parents].clone()
        if i >= 1:
            new_phys[:, i - 1] = parents
        phys = new_phys
        torch.cuda.synchronize()
        assert torch.equal(sess.tokens[:, : i + 1].cpu().long(), out[:, : i + 1]), f"token histories differ at step {i}"
        if i >= 1:
            assert torch.equal(sess.slots[:, :i].cpu().long(), phys[:, :i]), f"KV ancestry differs at step {i}"
        torch.testing.assert_close(sess.logprobs.cpu(), cur, rtol=1e-5, atol=2e-4)
        if torch.all((out == eos).any(dim=1)).item():
            stopped_at = i
            break
    st = sess.state.cpu()
    if stopped_at is None:
        assert int(st[2]) == 0 and int(st[0]) == steps
    else:
        assert int(st[2]) == 1 and int(st[3]) == stopped_at, (st.tolist(), stopped_at)
    if eos_boost:
        assert stopped_at is not None, "the EOS case never stopped: raise eos_boost"


@pytest.mark.parametrize("n,beams,group,steps,per", [(5, 4, 2, 14, 2), (3, 10, 2, 12, 1), (7, 2, 1, 14, 3)])
def test_lockstep_sessions_share_the_whole_batch_stop(cuda_device, n, beams, group, steps, per):
    """A batch of more than 16 beam rows is split over several decode sessions.  The reference stops the WHOLE batch
    at the first step where every beam of every input holds an EOS, and inputs that finish early keep extending and
    re-ranking until then (model_unified.py:833): the sessions step in lock-step and share a device-side group state.
    Same logits into `pcy_decode_select_group` and the oracle's step — identical tokens / scores at every step for
    every input, and the same stop step."""
    from oracle.generate import beam_select_step
    from procyon_b200.model.pmc_llama import SELECT_BEAM

    V, eos = 1003, 17
    oc, sd, m = _tiny("sel", vocab=V, seed=1)
    dev = torch.device("cuda")
    chunks = [(i0, min(n, i0 + per)) for i0 in range(0, n, per)]
    sessions = [m.new_session(b - a, beams, 8, steps, dev, False, False) for a, b in chunks]
    for s_ in sessions:
        s_.reset(None)
    gstate = torch.zeros(4, device=dev, dtype=torch.int32)
    gstate[3] = n
    bb = n * beams
    g = torch.Generator().manual_seed(n * 100 + beams)
    out = torch.zeros(bb, steps, dtype=torch.int64)
    cur = torch.zeros(bb)
    stopped_at = None
    finish_of_input = {}
    for i in range(steps):
        logits = torch.randn(bb, V, generator=g) * 3.0
        if i >= 2:  # inputs are pushed towards EOS one after the other: they finish at different steps
            hot = (torch.arange(bb) // beams) < (i - 1) * max(1, n // 4)
            logits[hot, eos] += 30.0
        if i == 0:
            logits = logits.view(n, beams, V)[:, :1].expand(n, beams, V).reshape(bb, V).contiguous()
        for k, ((a, b), s_) in enumerate(zip(chunks, sessions)):
            s_.logits_cur.copy_(logits[a * beams:b * beams])
            s_.select(SELECT_BEAM, group, 0.8, eos, True, gstate, k == len(sessions) - 1)
        lp = torch.log_softmax(logits, dim=-1) + cur[:, None]
        mg = []
        beam_select_step(lp, out, cur, i, n, beams, group, 0.8, mg)
        assert min(mg) > 1e-4
        torch.cuda.synchronize()
        got = torch.cat([s_.tokens[:, : i + 1].cpu().long() for s_ in sessions], 0)
        assert torch.equal(got, out[:, : i + 1]), f"token histories differ at step {i}"
        torch.testing.assert_close(torch.cat([s_.logprobs.cpu() for s_ in sessions]), cur, rtol=1e-5, atol=2e-4)
        has = (out == eos).any(dim=1).view(n, beams).all(dim=1)
        for j in range(n):
            if bool(has[j]) and j not in finish_of_input:
                finish_of_input[j] = i
        if bool(has.all()):
            stopped_at = i
            break
        assert int(gstate[1].item()) == 0, f"stopped at step {i} although inputs {(~has).nonzero().flatten().tolist()} go on"
    assert stopped_at is not None and len(set(finish_of_input.values())) > 1, "inputs should finish at different steps"
    gs = gstate.cpu()
    assert int(gs[1]) == 1 and int(gs[2]) == stopped_at, (gs.tolist(), stopped_at)
    # one more step after the stop is a no-op for every session
    before = [s_.tokens.clone() for s_ in sessions]
    for k, s_ in enumerate(sessions):
        s_.select(SELECT_BEAM, group, 0.8, eos, True, gstate, k == len(sessions) - 1)
    torch.cuda.synchronize()
    assert all(torch.equal(a, s_.tokens) for a, s_ in zip(before, sessions))
    assert all(int(s_.state[0].item()) == stopped_at + 1 for s_ in sessions)


# ------------------------------------------------------------------------------------------------ (B) forward
@pytest.mark.parametrize("kind,n,beams,group,S,pad,steps", [
    ("gq4", 1, 10, 2, 300, 0, 8), ("gq4", 2, 6, 3, 260, 9, 6), ("gq2", 2, 4, 2, 24, 5, 8), ("gq4", 1, 16, 4, 129, 0, 5),
    ("gq4", 1, 2, 1, 200, 0, 8), ("gq4", 1, 4, 2, 140, 0, 8),
])
def test_forced_beam_ancestry_logits_match_oracle_every_step(cuda_device, kind, n, beams, group, S, pad, steps):
    """The oracle's beam decisions are forced into the device session (token tables + physical-KV-row tables, exactly
    what `select` would write) and the decode forward runs under them: the logits of EVERY beam row at EVERY step
    against the oracle's (which physically reorders its KV cache like the reference, model_unified.py:830-832).
    Row counts 2 and 4 go through the greedy persistent kernel, 6..16 through the persistent beam kernel (attention over the
    union of the beams' keys with per-key beam masks)."""
    from oracle.generate import generate_beam_search as oracle_beam
    from procyon_b200 import _lib

    oc, sd, m = _tiny(kind)
    ids, emb, mask = _inputs(oc, sd, n, S, seed=7 * beams + S, pad_left=pad)
    trace = []
    oracle_beam(sd, oc, emb.float(), mask, max_len=steps, beam_size=beams, beam_group_size=group,
                diversity_penalty=0.8, eos_id=-5, act_round="bf16", mask_pads_in_decode=True, trace=trace)
    dev = torch.device("cuda")
    lib = _lib.load()
    lib.pcy_set_decode_megakernel(4)
    try:
        sess = m.get_session(n, beams, S, steps, dev, pad > 0, False)
        sel = torch.arange(n, device=dev, dtype=torch.int32) * S + (S - 1)
        am = mask.cuda() if pad else None
        _, _, logits, valid = m.prefill(emb.cuda(), am, want_cache=True, want_hidden=False, sel_rows=sel,
                                        kv_out=sess.kv_prompt)
        if pad:
            sess.prompt_valid.copy_(valid)
        sess.reset(logits)
        bb = n * beams
        toks = torch.zeros(bb, steps, dtype=torch.int32)
        slots = torch.zeros(bb, steps, dtype=torch.int32)
        worst = 0.0
        for i, tr in enumerate(trace):
            got = sess.logits_cur.cpu()
            assert torch.isfinite(got).all()
            worst = max(worst, (got - tr["logits"]).abs().max().item())
            torch.testing.assert_close(got, tr["logits"], rtol=3e-2, atol=5e-2, msg=lambda t: f"step {i}: {t}")
            if i + 1 == len(trace):
                break
            par = tr["parents"]
            toks = toks[par].clone()
            toks[:, i] = tr["tokens"].to(torch.int32)
            slots = slots[par].clone()
            if i >= 1:
                slots[:, i - 1] = par.to(torch.int32)
            sess.tokens.copy_(toks)
            sess.slots.copy_(slots)
            sess.state[0] = i + 1
            sess.forward()
    finally:
        lib.pcy_set_decode_megakernel(1)
    print(f"forced-ancestry decode: worst |logit error| {worst:.4f} (logits are O(20) with the structured head)")


# ------------------------------------------------------------------------------------------------ (D) every decision
@pytest.mark.parametrize("kind,n,beams,group,S,pad,steps", [
    ("gq4", 1, 10, 2, 300, 0, 12), ("gq4", 1, 10, 5, 40, 0, 24), ("gq4", 2, 6, 3, 260, 9, 10), ("gq4", 1, 16, 4, 129, 0, 8),
    ("gq2", 2, 4, 2, 24, 5, 12), ("gq2", 2, 6, 1, 24, 0, 10), ("gq4", 3, 5, 5, 30, 4, 10), ("gq4", 1, 2, 1, 200, 0, 12),
])
def test_every_device_beam_decision_is_the_oracles(cuda_device, kind, n, beams, group, S, pad, steps):
    from oracle.generate import _step, beam_select_step
    from procyon_b200.model.pmc_llama import SELECT_BEAM

    oc, sd, m = _tiny(kind)
    V = oc.vocab
    ids, emb, mask = _inputs(oc, sd, n, S, seed=11 * beams + S, pad_left=pad)
    dev = torch.device("cuda")
    sess = m.get_session(n, beams, S, steps, dev, pad > 0, False)
    sel = torch.arange(n, device=dev, dtype=torch.int32) * S + (S - 1)
    _, _, logits0, valid = m.prefill(emb.cuda(), mask.cuda() if pad else None, want_cache=True, want_hidden=False,
                                     sel_rows=sel, kv_out=sess.kv_prompt)
    if pad:
        sess.prompt_valid.copy_(valid)
    sess.reset(logits0)
    bb = n * beams
    embeds_rep = torch.repeat_interleave(emb.float(), beams, dim=0)
    mask_rep = torch.repeat_interleave(mask, beams, dim=0)
    hist = torch.zeros(bb, steps, dtype=torch.int64)  # the DEVICE's beam histories, followed by the oracle
    score = torch.zeros(bb)                            # the oracle's score of those histories
    past = None
    n_identical = n_close_calls = 0
    for i in range(steps):
        logits_o, past = _step(sd, oc, i, embeds_rep, mask_rep, hist, past, True, act_round="bf16")
        past = [[k.clone(), v.clone()] for k, v in past]
        got = sess.logits_cur.cpu()
        torch.testing.assert_close(got, logits_o, rtol=3e-2, atol=8e-2, msg=lambda t: f"logits at step {i}: {t}")
        sess.select(SELECT_BEAM, group, 0.8, -5, True)
        torch.cuda.synchronize()
        tok_d = sess.tokens[:, : i + 1].cpu().long()
        par_d = sess.slots[:, i - 1].cpu().long() if i >= 1 else torch.arange(bb) // beams * beams
        # histories must be the parents' histories + one token (pure bookkeeping: exact)
        if i >= 1:
            assert torch.equal(tok_d[:, :i], hist[par_d][:, :i]), f"step {i}: reordered histories are not the parents'"
        base = torch.log_softmax(logits_o, dim=-1) + score[:, None]
        # what the oracle decides from the same state
        o_hist, o_score, mg = hist.clone(), score.clone(), []
        par_o = beam_select_step(base.clone(), o_hist, o_score, i, n, beams, group, 0.8, mg)
        identical = torch.equal(o_hist[:, : i + 1], tok_d) and (i == 0 or torch.equal(par_o, par_d))
        new_score = torch.empty(bb)
        for inp in range(n):
            b0 = inp * beams
            for g in range(beams // group):
                gs, ge = b0 + g * group, b0 + (g + 1) * group
                inc = 1 if i == 0 else group
                lp = base[gs:gs + inc].clone()
                if g:
                    lp -= 0.8 * torch.bincount(tok_d[b0:gs, i], minlength=V)  # penalty from the DEVICE's earlier groups
                flat = (par_d[gs:ge] - gs).clamp(0, inc - 1) * V + tok_d[gs:ge, i]
                assert ((par_d[gs:ge] >= gs) & (par_d[gs:ge] < gs + max(inc, 1))).all() or i == 0, \
                    f"step {i}: a beam extends a row outside its own group"
                assert flat.unique().numel() == group, f"step {i} group {g}: the same candidate was selected twice"
                mine = lp.ravel()[flat]
                best = lp.ravel().topk(group).values
                # a top-k within MARGIN, listed in an order consistent within MARGIN
                assert (mine >= best[-1] - MARGIN).all(), \
                    f"step {i} group {g}: selected scores {mine.tolist()} vs oracle top-k {best.tolist()}"
                assert (mine[:-1] >= mine[1:] - MARGIN).all(), f"step {i} group {g}: beams out of order {mine.tolist()}"
                new_score[gs:ge] = mine
        if identical:
            n_identical += 1
        else:
            assert min(mg) < MARGIN, (f"step {i}: the device's selection differs although every oracle margin is "
                                      f">= {min(mg):.3f}")
            n_close_calls += 1
        # the device's running scores vs the oracle's score of the same histories (per-step deviations add up)
        torch.testing.assert_close(sess.logprobs.cpu(), new_score, rtol=0, atol=0.06 * (i + 1) + 0.02)
        hist = torch.zeros_like(hist)
        hist[:, : i + 1] = tok_d
        score = new_score
        for l in range(len(past)):
            past[l][0] = past[l][0][par_d]
            past[l][1] = past[l][1][par_d]
        if i + 1 < steps:
            sess.forward()
    print(f"beam decisions: {n_identical} steps identical to the oracle, {n_close_calls} with a sub-margin deviation")
    assert n_identical >= 1


# ------------------------------------------------------------------------------------------------ (C) end to end
# (kind, n, beams, group, pad, S, max_len, seed): seeds from scripts/find_beam_seeds.py — the oracle's smallest decision
# margin over the whole generation is >= MARGIN for each of them (re-checked below, so a stale seed fails loudly)
E2E_CASES = [
    ("gq2", 1, 4, 4, 0, 24, 8, 2061),    # oracle margin 0.141
    ("gq2", 1, 6, 1, 0, 24, 8, 2315),    # oracle margin 0.188
    ("gq2", 2, 4, 2, 5, 24, 6, 2032),    # oracle margin 0.130 (left-padded second prompt)
    ("gq4", 1, 4, 2, 0, 24, 10, 2078),   # oracle margin 0.238 (4 rows: persistent decode kernel)
    ("gq4", 2, 2, 1, 3, 24, 10, 2075),   # oracle margin 0.298
    ("gq4", 1, 10, 2, 0, 300, 5, 2593),  # oracle margin 0.164 (evaluation default: 10 beams in groups of 2)
    ("gq4", 2, 6, 3, 9, 260, 4, 2104),   # oracle margin 0.125
    # __MORE_E2E_CASES__
]


@pytest.mark.parametrize("kind,n,beams,group,pad,S,max_len,seed", E2E_CASES)
def test_beam_search_every_beam_equals_oracle(cuda_device, kind, n, beams, group, pad, S, max_len, seed):
    from oracle.generate import generate_beam_search as oracle_beam
    from procyon_b200.model.generation import generate_beam_search

    oc, sd, m = _tiny(kind)
    ids, emb, mask = _inputs(oc, sd, n, S, seed=seed, pad_left=pad)
    trace = []
    ro, rlp, rlogits = oracle_beam(sd, oc, emb.float(), mask, max_len=max_len, beam_size=beams, beam_group_size=group,
                                   diversity_penalty=0.8, eos_id=-5, act_round="bf16", mask_pads_in_decode=True,
                                   trace=trace)
    margin = min(t["margin"] for t in trace)
    assert margin >= MARGIN, f"seed {seed} no longer has clear margins ({margin:.3f}): rerun scripts/find_beam_seeds.py"
    for use_graph in (True, False):
        out, lp, logits = generate_beam_search(m, emb.cuda(), mask.cuda() if pad else None, max_len=max_len,
                                               beam_size=beams, beam_group_size=group, diversity_penalty=0.8,
                                               eos_token_id=-5, use_graph=use_graph)
        assert torch.equal(out, ro), f"beams differ from the oracle (smallest oracle margin {margin:.3f})"
        torch.testing.assert_close(lp, rlp, rtol=1e-2, atol=8e-2)
        torch.testing.assert_close(logits.cpu(), rlogits, rtol=3e-2, atol=6e-2)

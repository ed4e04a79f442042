"""Llama prefill / decode / generation on the GPU vs the CPU oracle (tiny Llama-shaped config, head_dim 128)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfgs(kind="gq2", max_pos=512):
    from oracle.llama import LlamaCfg
    from procyon_b200.model.pmc_llama import LlamaConfig

    # gq2: 4 query / 2 kv heads -> one-launch-per-op decode path; gq4: 8 / 2 heads (Llama-3's ratio) -> rows <= 4 take
    # the persistent decode kernel
    # gq4wide: ffn 8448 = 2 full 4096-element weight chunks + a 256-element tail per w_down row (multi-chunk rows of
    # the persistent kernel's weight ring)
    H, d, f = {"gq2": (4, 512, 1024), "gq4": (8, 1024, 1024), "gq4wide": (8, 1024, 8448)}[kind]
    oc = LlamaCfg(d_model=d, n_layers=2, n_heads=H, n_kv_heads=2, ffn_dim=f, vocab=1003, max_pos=max_pos)
    pc = LlamaConfig(hidden_size=d, intermediate_size=f, num_hidden_layers=2, num_attention_heads=H,
                     num_key_value_heads=2, vocab_size=1003, max_position_embeddings=max_pos)
    return oc, pc


def _build(sd, pc):
    from procyon_b200.model.pmc_llama import LlamaPostTokenization

    m = LlamaPostTokenization(config=pc, dtype=torch.bfloat16)
    m.model.load_state_dict(sd, strict=True)
    return m.cuda()


@pytest.fixture(scope="module", params=["gq2", "gq4"])
def llama(cuda_device, request):
    from oracle.llama import random_llama_state_dict

    oc, pc = _cfgs(request.param)
    sd = random_llama_state_dict(oc, seed=3)
    return oc, sd, _build(sd, pc)


def _inputs(oc, sd, B, S, seed, pad_left=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, oc.vocab, (B, S), generator=g)
    emb = sd["model.embed_tokens.weight"][ids].clone()
    emb[:, S // 2] = (torch.randn(B, oc.d_model, generator=g) * 0.5).bfloat16()  # a spliced soft token
    mask = torch.ones(B, S)
    if pad_left and B > 1:
        mask[1, :pad_left] = 0
    return ids, emb, mask


def test_prefill_hidden_logits_loss(llama):
    from oracle.llama import llama_forward

    oc, sd, m = llama
    ids, emb, mask = _inputs(oc, sd, 2, 70, seed=1, pad_left=9)
    labels = ids.clone()
    labels[:, :20] = -100
    labels[1, :9] = -100
    ref = llama_forward(sd, oc, inputs_embeds=emb.float(), attention_mask=mask, labels=labels, act_round="bf16")
    out = m(input_embeds=emb.cuda(), attn_masks=mask.cuda(), full_labels=labels.cuda())
    keep = mask.bool()
    h = out.hidden_states[-1].float().cpu()
    assert len(out.hidden_states) == oc.n_layers + 1
    torch.testing.assert_close(h[keep], ref["hidden_states"][-1][keep], rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(out.logits.cpu()[keep], ref["logits"][keep], rtol=3e-2, atol=4e-2)
    assert abs(out.loss.item() - ref["loss"].item()) < 2e-2
    assert out.past_key_values is None


@pytest.mark.parametrize("S,lens", [(160, (70, 41, 133)), (300, (150, 150)), (64, (64, 10))])
def test_prefill_skips_trailing_padding(llama, S, lens):
    """forward() batches arrive right-padded to max_text_len (reference: model_unified.py:1283, 2048 for ProCyon-Full).
    Positions after the last valid token of the longest row are not computed (`trim_trailing_pads`): hidden states,
    logits and the LM loss at every VALID position must equal the full computation and the oracle, the [PROT]-style
    hidden-state sums too, and the skipped positions come back as zeros."""
    from oracle.llama import llama_forward
    from procyon_b200.model.pmc_llama import LlamaPostTokenization

    oc, sd, m = llama
    B = len(lens)
    ids, emb, _ = _inputs(oc, sd, B, S, seed=S)
    mask = torch.zeros(B, S)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    labels = ids.clone()
    labels[:, :8] = -100
    labels[mask == 0] = -100
    pick = torch.zeros(B, S, dtype=torch.bool)
    for b, n in enumerate(lens):
        pick[b, n - 1] = True  # the last valid token of every row, like a [PROT] position
    ref = llama_forward(sd, oc, inputs_embeds=emb.float(), attention_mask=mask, labels=labels, act_round="bf16")

    def run(trim):
        LlamaPostTokenization.trim_trailing_pads = trim
        try:
            return m(input_embeds=emb.cuda(), attn_masks=mask.cuda(), full_labels=labels.cuda(),
                     sum_hidden_rows=pick.cuda())
        finally:
            LlamaPostTokenization.trim_trailing_pads = True

    a, b_ = run(True), run(False)
    keep = mask.bool()
    ha, hb = a.hidden_states[-1].float().cpu(), b_.hidden_states[-1].float().cpu()
    # (300 k elements of O(1) magnitude after two bf16 layers: one or two land a bf16 ulp of an intermediate past 3e-2)
    torch.testing.assert_close(ha[keep], ref["hidden_states"][-1][keep], rtol=3e-2, atol=5e-2)
    torch.testing.assert_close(ha[keep], hb[keep], rtol=3e-2, atol=5e-2)
    torch.testing.assert_close(a.logits.cpu()[keep], ref["logits"][keep], rtol=3e-2, atol=4e-2)
    assert abs(a.loss.item() - ref["loss"].item()) < 2e-2 and abs(a.loss.item() - b_.loss.item()) < 2e-2
    torch.testing.assert_close(a.hidden_sum.cpu(), b_.hidden_sum.cpu(), rtol=3e-2, atol=6e-2)
    n_max = max(lens)
    if n_max < S:
        assert float(ha[:, n_max:].abs().max()) == 0.0          # skipped positions: zeros
        assert float(hb[:, n_max:].abs().max()) > 0.0           # (the full computation fills them)


def test_decode_step_matches_oracle_and_prefill(llama):
    from oracle.llama import llama_forward

    oc, sd, m = llama
    ids, emb, mask = _inputs(oc, sd, 2, 33, seed=2)
    r0 = llama_forward(sd, oc, inputs_embeds=emb.float(), act_round="bf16")
    out0 = m(input_embeds=emb.cuda(), attn_masks=None, use_cache=True)
    nxt = torch.tensor([[5], [77]])
    r1 = llama_forward(sd, oc, input_ids=nxt, past=r0["past"], act_round="bf16")
    out1 = m(input_ids=nxt.cuda(), past_key_values=out0.past_key_values, use_cache=True)
    torch.testing.assert_close(out1.logits.cpu()[:, 0], r1["logits"][:, 0], rtol=3e-2, atol=4e-2)
    nxt2 = torch.tensor([[9], [1]])
    r2 = llama_forward(sd, oc, input_ids=nxt2, past=r1["past"], act_round="bf16")
    out2 = m(input_ids=nxt2.cuda(), past_key_values=out1.past_key_values, use_cache=True)
    torch.testing.assert_close(out2.logits.cpu()[:, 0], r2["logits"][:, 0], rtol=3e-2, atol=4e-2)
    # property: decoding token x after a prefill of S positions == prefill of S+1 positions (same kernels' math)
    emb2 = torch.cat([emb, sd["model.embed_tokens.weight"][nxt]], dim=1)
    full = m(input_embeds=emb2.cuda(), attn_masks=None)
    torch.testing.assert_close(out1.logits[:, 0], full.logits[:, -1], rtol=2e-2, atol=3e-2)


def _margin_ok(ref_logits, tol=0.05):
    top2 = ref_logits.topk(2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) > tol


def test_greedy_generation(llama):
    from oracle.generate import generate_greedy as oracle_greedy
    from procyon_b200.model.generation import generate_greedy

    oc, sd, m = llama
    ids, emb, mask = _inputs(oc, sd, 2, 40, seed=4)
    ro, rlp, rlogits = oracle_greedy(sd, oc, emb.float(), None, max_len=12, act_round="bf16")
    for use_graph in (False, True):
        out, lp, logits = generate_greedy(m, emb.cuda(), None, max_len=12, use_graph=use_graph)
        assert out.shape == (2, 12) and logits.shape == (2, 12, oc.vocab)
        # exact token ids as long as the oracle's own top-2 margin is not inside the bf16 noise floor
        for b in range(2):
            for s in range(12):
                if not torch.equal(out[b, : s + 1], ro[b, : s + 1]):
                    assert not _margin_ok(rlogits[b, s]), f"token mismatch at row {b} step {s} with a clear margin"
                    break
            else:
                torch.testing.assert_close(lp[b], rlp[b], rtol=1e-2, atol=5e-2)
                torch.testing.assert_close(logits[b].cpu(), rlogits[b], rtol=3e-2, atol=5e-2)


BEAM_MARGIN = 0.12  # twice the worst log-prob deviation of the bf16 forward measured on B200 (test_gpu_beam_strict.py)


def _clear_steps(trace, margin=BEAM_MARGIN):
    """Number of leading steps whose every beam decision has a margin >= `margin` in the oracle: up to there a
    forward pass that is right to half the margin MUST reproduce every beam exactly."""
    n = 0
    for t in trace:
        if t["margin"] < margin:
            break
        n += 1
    return n


@pytest.mark.parametrize("n,beams,group,pad", [(2, 4, 2, 0), (2, 4, 4, 0), (2, 6, 1, 0), (2, 4, 2, 5), (1, 4, 2, 0),
                                                (2, 2, 1, 3)])
def test_beam_search(llama, n, beams, group, pad):
    """Random-weight model (flat logits, top-k gaps of ~0.3): EVERY beam must equal the oracle's on every step before
    the first decision the oracle itself takes with a margin inside the forward noise; nothing is asserted after it
    here — tests/test_gpu_beam_strict.py verifies every later decision one by one."""
    from oracle.generate import generate_beam_search as oracle_beam
    from procyon_b200.model.generation import generate_beam_search

    oc, sd, m = llama
    ids, emb, mask = _inputs(oc, sd, n, 24, seed=beams * 10 + group, pad_left=pad)
    am = mask if pad else None
    trace = []
    ro, rlp, rlogits = oracle_beam(sd, oc, emb.float(), mask if pad else torch.ones_like(mask), max_len=8,
                                   beam_size=beams, beam_group_size=group, diversity_penalty=0.8, eos_id=-5,
                                   act_round="bf16", mask_pads_in_decode=True, trace=trace)
    out, lp, logits = generate_beam_search(m, emb.cuda(), am.cuda() if am is not None else None, max_len=8,
                                           beam_size=beams, beam_group_size=group, diversity_penalty=0.8,
                                           eos_token_id=-5)
    assert out.shape == ro.shape == (n, beams, 8)
    clear = _clear_steps(trace)
    assert torch.equal(out[..., :clear], ro[..., :clear]), f"beams differ within the {clear} clear-margin steps"
    same = (out == ro).all(dim=-1)
    # beams whose whole token history agrees must agree on score and on the gathered per-step logits
    torch.testing.assert_close(lp[same], rlp[same], rtol=1e-2, atol=6e-2)
    torch.testing.assert_close(logits.cpu()[same], rlogits[same], rtol=3e-2, atol=6e-2)


def test_beam_search_stops_on_eos(llama):
    """Early exit only when EVERY beam holds an EOS (model_unified.py:833); the tail of `out` stays zero."""
    from oracle.generate import generate_beam_search as oracle_beam
    from procyon_b200.model.generation import generate_beam_search

    oc, sd, m = llama
    ids, emb, mask = _inputs(oc, sd, 1, 16, seed=99)
    out, lp, logits = generate_beam_search(m, emb.cuda(), None, max_len=10, beam_size=2, beam_group_size=2,
                                           eos_token_id=-5)
    eos = int(out[0, 0, 1])  # a token the best beam emits at step 1
    trace = []
    ro, rlp, rlogits = oracle_beam(sd, oc, emb.float(), torch.ones(1, 16), max_len=10, beam_size=2, beam_group_size=2,
                                   diversity_penalty=0.8, eos_id=eos, act_round="bf16", trace=trace)
    out2, lp2, logits2 = generate_beam_search(m, emb.cuda(), None, max_len=10, beam_size=2, beam_group_size=2,
                                              eos_token_id=eos)
    steps_ref = rlogits.shape[2]
    clear = _clear_steps(trace)
    assert torch.equal(out2[..., :clear], ro[..., :clear])
    if clear == steps_ref:  # every decision up to the oracle's stop was clear: the device must stop at the same step
        assert logits2.shape[2] == steps_ref
        assert int(out2[..., steps_ref:].abs().sum()) == 0
        torch.testing.assert_close(lp2, rlp, rtol=1e-2, atol=6e-2)
    # whatever the margins: once stopped, nothing is written after the stop step
    assert int(out2[..., logits2.shape[2]:].abs().sum()) == 0


def _forced_decode(m, emb, mask, forced, max_rows):
    """Prefill + teacher-forced decode steps; returns the logits of every step [rows, steps + 1, V] (cpu)."""
    from procyon_b200 import _lib

    lib = _lib.load()
    lib.pcy_set_decode_megakernel(max_rows)
    try:
        out = m(input_embeds=emb.cuda(), attn_masks=mask.cuda() if mask is not None else None, use_cache=True)
        sess = out.past_key_values
        logs = [sess.logits_cur.clone().cpu()]
        for i in range(forced.shape[1]):
            o = m(input_ids=forced[:, i:i + 1].cuda(), past_key_values=sess)
            logs.append(o.logits[:, 0].cpu())
    finally:
        lib.pcy_set_decode_megakernel(1)
    return torch.stack(logs, 1)


def _forced_oracle(sd, oc, emb, mask, forced):
    from oracle.llama import llama_forward

    r = llama_forward(sd, oc, inputs_embeds=emb.float(), attention_mask=mask, act_round="bf16")
    logs = [r["logits"][:, -1]]
    am = mask
    for i in range(forced.shape[1]):
        if am is not None:
            am = torch.cat([am, torch.ones(am.shape[0], 1)], dim=1)
        r = llama_forward(sd, oc, input_ids=forced[:, i:i + 1], past=r["past"], attention_mask=am, act_round="bf16")
        logs.append(r["logits"][:, -1])
    return torch.stack(logs, 1)


def _assert_argmax_where_clear(ours, ref, tol=0.05):
    top2 = ref.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > tol
    assert clear.any()
    assert torch.equal(ours.argmax(-1)[clear], ref.argmax(-1)[clear]), "arg-max token ids differ at a clear margin"


@pytest.mark.parametrize("kind,rows", [("gq4", 3), ("gq4wide", 1), ("gq4wide", 2), ("gq4wide", 4)])
def test_persistent_and_per_op_decode_match_oracle(cuda_device, kind, rows):
    """The single-launch decode step and the one-launch-per-op path, teacher-forced with the same tokens: the logits
    of every step of BOTH against the oracle, and exact arg-max ids wherever the oracle's margin is clear."""
    from oracle.llama import random_llama_state_dict

    oc, pc = _cfgs(kind)
    sd = random_llama_state_dict(oc, seed=3)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, rows, 50, seed=7, pad_left=4)
    forced = torch.randint(0, oc.vocab, (rows, 9), generator=torch.Generator().manual_seed(rows))
    ref = _forced_oracle(sd, oc, emb, mask, forced)
    for max_rows in (4, 0):
        got = _forced_decode(m, emb, mask, forced, max_rows)
        assert torch.isfinite(got).all()
        torch.testing.assert_close(got, ref, rtol=3e-2, atol=4e-2, msg=lambda t: f"megakernel max_rows={max_rows}: {t}")
        _assert_argmax_where_clear(got, ref)


@pytest.mark.parametrize("kind,rows", [("gq4", 1), ("gq4wide", 2), ("gq4wide", 4)])
def test_persistent_decode_without_producer_warp_is_bit_identical(cuda_device, kind, rows):
    """`pcy_set_decode_self_refill(1)`: the greedy persistent kernel with every consumer warp refilling its own ring
    slots (no producer warp, no empty barriers).  The arithmetic and its order are unchanged, so every step's logits
    must be bit-identical to the producer-warp form (multi-chunk rows of the wide config included)."""
    from oracle.llama import random_llama_state_dict
    from procyon_b200 import _lib

    oc, pc = _cfgs(kind)
    sd = random_llama_state_dict(oc, seed=3)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, rows, 150, seed=7, pad_left=4)
    forced = torch.randint(0, oc.vocab, (rows, 7), generator=torch.Generator().manual_seed(rows))
    lib = _lib.load()
    try:
        lib.pcy_set_decode_self_refill(0)
        a = _forced_decode(m, emb, mask, forced, 4)
        lib.pcy_set_decode_self_refill(1)
        b = _forced_decode(m, emb, mask, forced, 4)
    finally:
        lib.pcy_set_decode_self_refill(0)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)


@pytest.mark.parametrize("rows,S", [(1, 1300), (4, 1800)])
def test_persistent_decode_long_context(cuda_device, rows, S):
    """Long prompts in the single-launch decode step: more than 12 KV splits per head (two-pass merge of the split
    partials) and, with 4 rows x 2 kv heads x 19 splits, more attention work items than CTAs (several per CTA).
    Teacher-forced, every step against the oracle (and the per-op path against it too)."""
    from oracle.llama import random_llama_state_dict

    oc, pc = _cfgs("gq4", max_pos=2048)
    sd = random_llama_state_dict(oc, seed=5)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, rows, S, seed=11, pad_left=7)
    forced = torch.randint(0, oc.vocab, (rows, 5), generator=torch.Generator().manual_seed(S))
    ref = _forced_oracle(sd, oc, emb, mask, forced)
    for max_rows in (4, 0):
        got = _forced_decode(m, emb, mask, forced, max_rows)
        assert torch.isfinite(got).all()
        torch.testing.assert_close(got, ref, rtol=3e-2, atol=4e-2, msg=lambda t: f"megakernel max_rows={max_rows}: {t}")
        _assert_argmax_where_clear(got, ref)


@pytest.mark.parametrize("n,beams,S,pad", [(1, 10, 300, 0), (2, 6, 260, 9), (1, 16, 129, 0), (1, 5, 256, 0)])
def test_beam_search_shared_prompt_attention(cuda_device, n, beams, S, pad):
    """Beam search with prompts longer than a key split: the splits inside the prompt go through the tensor-core kernel
    that serves all beams of an input at once (decode_attn_shared_prompt_kernel), the generated tail through the
    per-row kernel.  One decode step on identical state must give the same logits for every beam row as the per-row
    kernel alone (+ the scalar GEMV path), and whole generations must mostly agree with it and with the oracle."""
    from oracle.generate import generate_beam_search as oracle_beam
    from oracle.llama import random_llama_state_dict
    from procyon_b200 import _lib
    from procyon_b200.model.generation import generate_beam_search
    from procyon_b200.model.pmc_llama import SELECT_BEAM

    oc, pc = _cfgs("gq4", max_pos=1024)
    sd = random_llama_state_dict(oc, seed=13)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, n, S, seed=beams + S, pad_left=pad)
    am = mask.cuda() if pad else None
    lib = _lib.load()
    # ---- one step, same device state, both paths ----
    sess = m.get_session(n, beams, S, 8, torch.device("cuda"), pad > 0, False)
    sel = torch.tensor([(i + 1) * S - 1 for i in range(n)], device="cuda", dtype=torch.int32)
    _, _, logits, valid = m.prefill(emb.cuda(), am, want_cache=True, want_hidden=False, sel_rows=sel,
                                    kv_out=sess.kv_prompt)
    if pad:
        sess.prompt_valid.copy_(valid)
    sess.reset(logits)
    group = max(1, beams // 2) if beams % 2 == 0 else beams
    sess.select(SELECT_BEAM, group, 0.8, -5, False)
    try:
        lib.pcy_set_decode_rows_megakernel(0)  # this part compares the two one-launch-per-op variants
        lib.pcy_set_skinny_mma(1)
        sess.forward()
        la = sess.logits_cur.clone()
        lib.pcy_set_skinny_mma(0)
        sess.forward()
        lb = sess.logits_cur.clone()
        lib.pcy_set_decode_rows_megakernel(1)  # ... and the persistent beam kernel on the same state against both
        sess.forward()
        lc = sess.logits_cur.clone()
    finally:
        lib.pcy_set_skinny_mma(1)
        lib.pcy_set_decode_rows_megakernel(1)
    assert torch.isfinite(la).all() and torch.isfinite(lc).all()
    torch.testing.assert_close(la, lb, rtol=3e-2, atol=4e-2)
    torch.testing.assert_close(lc, la, rtol=3e-2, atol=4e-2)
    # ---- whole generations ----
    kw = dict(max_len=5, beam_size=beams, beam_group_size=group, diversity_penalty=0.8, eos_token_id=-5)
    o1, lp1, lg1 = generate_beam_search(m, emb.cuda(), am, **kw)
    ro, rlp, rlogits = oracle_beam(sd, oc, emb.float(), mask if pad else torch.ones_like(mask), max_len=5,
                                   beam_size=beams, beam_group_size=group, diversity_penalty=0.8, eos_id=-5,
                                   act_round="bf16", mask_pads_in_decode=True)
    # every beam identical to the oracle's on all steps before the first sub-noise decision margin
    tr = []
    oracle_beam(sd, oc, emb.float(), mask if pad else torch.ones_like(mask), max_len=5, beam_size=beams,
                beam_group_size=group, diversity_penalty=0.8, eos_id=-5, act_round="bf16", mask_pads_in_decode=True,
                trace=tr)
    clear = _clear_steps(tr)
    assert torch.equal(o1[..., :clear], ro[..., :clear]), f"beams differ within the {clear} clear-margin steps"
    for i in range(n):  # sequences both searches found must carry the same score
        ours = {tuple(o1[i, b].tolist()): float(lp1[i, b]) for b in range(beams)}
        ref = {tuple(ro[i, b].tolist()): float(rlp[i, b]) for b in range(beams)}
        for k in set(ours) & set(ref):
            assert abs(ours[k] - ref[k]) < 6e-2 + 1e-2 * abs(ref[k])


@pytest.mark.parametrize("rows,S,steps", [(1, 90, 14), (2, 185, 14), (4, 93, 8)])
def test_persistent_decode_crosses_key_split_boundaries(cuda_device, rows, S, steps):
    """Teacher-forced decoding (same forced tokens on both paths) long enough for the context to cross a 96-key split
    boundary of the persistent kernel's attention - the current token's K/V then live in a different split than the
    prompt tail: every step's logits must match the per-op path."""
    from oracle.llama import random_llama_state_dict
    from procyon_b200 import _lib

    oc, pc = _cfgs("gq4", max_pos=512)
    sd = random_llama_state_dict(oc, seed=21)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, rows, S, seed=S, pad_left=0)
    forced = torch.randint(0, oc.vocab, (rows, steps), generator=torch.Generator().manual_seed(S + rows)).cuda()
    lib = _lib.load()

    def run(max_rows):
        lib.pcy_set_decode_megakernel(max_rows)
        out = m(input_embeds=emb.cuda(), use_cache=True)
        sess, logs = out.past_key_values, []
        for i in range(steps):
            o = m(input_ids=forced[:, i:i + 1], past_key_values=sess)
            logs.append(o.logits[:, 0].clone())
        return torch.stack(logs, 1)

    try:
        a = run(4)
        b = run(0)
    finally:
        lib.pcy_set_decode_megakernel(1)
    assert torch.isfinite(a).all()
    torch.testing.assert_close(a, b, rtol=3e-2, atol=4e-2)
    ref = _forced_oracle(sd, oc, emb, None, forced.cpu())[:, 1:]
    torch.testing.assert_close(a.cpu(), ref, rtol=3e-2, atol=4e-2)
    torch.testing.assert_close(b.cpu(), ref, rtol=3e-2, atol=4e-2)
    _assert_argmax_where_clear(a.cpu(), ref)


@pytest.mark.parametrize("rows", [6, 10, 16])
def test_programmatic_dependent_launch_changes_nothing(cuda_device, rows):
    """The per-op decode chain launched with the programmatic-dependent-launch attribute (each kernel may become
    resident while its predecessor drains and prefetches weights before `griddepcontrol.wait`) against plain stream
    edges: bit-identical logits, eagerly and through CUDA-graph replays, and both against the oracle."""
    from oracle.llama import random_llama_state_dict
    from procyon_b200 import _lib
    from procyon_b200.model.generation import generate_beam_search

    oc, pc = _cfgs("gq4", max_pos=512)
    sd = random_llama_state_dict(oc, seed=31)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, rows, 150, seed=rows, pad_left=5)
    forced = torch.randint(0, oc.vocab, (rows, 6), generator=torch.Generator().manual_seed(rows))
    lib = _lib.load()
    try:
        lib.pcy_set_pdl(1)
        a = _forced_decode(m, emb, mask, forced, 0)
        lib.pcy_set_pdl(0)
        b = _forced_decode(m, emb, mask, forced, 0)
    finally:
        lib.pcy_set_pdl(1)
    assert torch.equal(a, b)
    ref = _forced_oracle(sd, oc, emb, mask, forced)
    torch.testing.assert_close(a, ref, rtol=3e-2, atol=4e-2)
    if rows <= 16:
        i1, e1, m1 = _inputs(oc, sd, 1, 150, seed=rows)
        kw = dict(max_len=12, beam_size=rows, beam_group_size=rows // 2, diversity_penalty=0.8, eos_token_id=-5)
        try:
            lib.pcy_set_decode_rows_megakernel(0)  # the per-op chain is what the attribute applies to
            lib.pcy_set_pdl(1)
            o1, lp1, lg1 = generate_beam_search(m, e1.cuda(), None, **kw)   # CUDA-graph replays
            lib.pcy_set_pdl(0)
            m.__dict__.pop("_sessions", None)  # (the cached session holds the step graph captured with the attribute)
            o2, lp2, lg2 = generate_beam_search(m, e1.cuda(), None, **kw)
        finally:
            lib.pcy_set_pdl(1)
            lib.pcy_set_decode_rows_megakernel(1)
        assert torch.equal(o1, o2) and torch.equal(lp1, lp2) and torch.equal(lg1, lg2)


@pytest.mark.parametrize("kind,B,S,pad", [("gq4", 2, 128, 0), ("gq4", 2, 200, 37), ("gq4", 1, 333, 0), ("gq4", 3, 257, 130),
                                          ("gq2", 2, 192, 5), ("gq4", 1, 1024, 0)])
def test_prefill_tcgen05_causal_attention_matches_oracle_and_mma_sync(cuda_device, kind, B, S, pad):
    """Prefill attention on tcgen05 (prompts of >= 128 positions, head_dim 128: causal mask inside the diagonal key
    steps, GQA, left-pad key mask, shifted last query tile when S % 128 != 0, key steps that end mid-tile) against the
    mma.sync kernel it replaces and against the oracle: hidden states, last-position logits and the KV cache it leaves
    for decoding."""
    from oracle.llama import llama_forward, random_llama_state_dict
    from procyon_b200 import _lib

    oc, pc = _cfgs(kind, max_pos=1024)
    sd = random_llama_state_dict(oc, seed=61)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, B, S, seed=S + B, pad_left=pad)
    am = mask.cuda() if pad else None
    lib = _lib.load()
    outs = {}
    try:
        for tc in (1, 0):
            lib.pcy_set_llama_tc_attention(tc)
            n0 = lib.pcy_launch_count()
            o = m(input_embeds=emb.cuda(), attn_masks=am)
            outs[tc] = (o.hidden_states[-1].float().cpu(), lib.pcy_launch_count() - n0)
            again = m(input_embeds=emb.cuda(), attn_masks=am).hidden_states[-1].float().cpu()
            assert torch.equal(outs[tc][0], again)  # deterministic
    finally:
        lib.pcy_set_llama_tc_attention(1)
    keep = mask.bool()
    assert torch.isfinite(outs[1][0][keep]).all()

    def close(a, b, what):
        # two bf16 pipelines: the DESIGN tolerance band (3e-2 + 3e-2 |ref|) for all but a 1e-3 tail of single-ulp flips
        # of the larger values, nothing off by more than 0.1
        err = (a - b).abs()
        band = 3e-2 + 3e-2 * b.abs()
        assert (err > band).float().mean().item() < 1e-3, (what, (err > band).float().mean().item())
        assert err.max().item() < 0.1, (what, err.max().item())

    close(outs[1][0][keep], outs[0][0][keep], "tcgen05 vs mma.sync")
    ref = llama_forward(sd, oc, inputs_embeds=emb.float(), attention_mask=mask if pad else None, act_round="bf16")
    close(outs[1][0][keep], ref["hidden_states"][-1][keep], "tcgen05 vs oracle")
    close(outs[0][0][keep], ref["hidden_states"][-1][keep], "mma.sync vs oracle")
    # decoding from the cache the tcgen05 prefill left behind
    forced = torch.randint(0, oc.vocab, (B, 3), generator=torch.Generator().manual_seed(S))
    got = _forced_decode(m, emb, mask if pad else None, forced, 4)
    want = _forced_oracle(sd, oc, emb, mask if pad else None, forced)
    torch.testing.assert_close(got, want, rtol=3e-2, atol=4e-2)

"""The persistent decode-step kernel for 3..16 rows (csrc/decode_rows_megakernel.cu: beam search, the path every shipped
caller of the reference runs - procyon/evaluate/framework/procyon.py:71-76) against the CPU oracle and against the
one-launch-per-op path on identical device state."""
import pytest
import torch

from test_gpu_llama import _assert_argmax_where_clear, _build, _cfgs, _forced_oracle, _inputs

pytestmark = pytest.mark.gpu


def _forced(m, emb, mask, forced, rows_kernel):
    """Prefill + teacher-forced decode steps (every row its own input); logits of every step [rows, steps + 1, V]."""
    from procyon_b200 import _lib

    lib = _lib.load()
    lib.pcy_set_decode_rows_megakernel(1 if rows_kernel else 0)
    try:
        out = m(input_embeds=emb.cuda(), attn_masks=mask.cuda() if mask is not None else None, use_cache=True)
        sess = out.past_key_values
        logs = [sess.logits_cur.clone().cpu()]
        for i in range(forced.shape[1]):
            o = m(input_ids=forced[:, i:i + 1].cuda(), past_key_values=sess)
            logs.append(o.logits[:, 0].cpu())
    finally:
        lib.pcy_set_decode_rows_megakernel(1)
    return torch.stack(logs, 1)


@pytest.mark.parametrize("kind,rows,S,pad", [("gq4", 2, 90, 0), ("gq4", 3, 50, 4), ("gq4", 5, 70, 0), ("gq4", 8, 33, 3), ("gq4", 9, 64, 0),
                                             ("gq4", 16, 40, 5), ("gq4wide", 10, 50, 4), ("gq4wide", 4, 130, 0)])
def test_rows_kernel_teacher_forced_matches_oracle_and_per_op(cuda_device, kind, rows, S, pad):
    """Independent rows (beams = 1, every row its own prompt): k-parts of the down projection (gq4wide: 9 parts, the
    last one shorter), one and two 8-row MMA column tiles, left-padded prompts; every step's logits against the oracle
    and against the per-op path, exact arg-max ids wherever the oracle's margin is clear."""
    from oracle.llama import random_llama_state_dict

    oc, pc = _cfgs(kind)
    sd = random_llama_state_dict(oc, seed=3)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, rows, S, seed=7 + rows, pad_left=pad)
    forced = torch.randint(0, oc.vocab, (rows, 9), generator=torch.Generator().manual_seed(rows))
    ref = _forced_oracle(sd, oc, emb, mask, forced)
    a = _forced(m, emb, mask, forced, True)
    b = _forced(m, emb, mask, forced, False)
    assert torch.isfinite(a).all()
    torch.testing.assert_close(a, ref, rtol=3e-2, atol=4e-2)
    torch.testing.assert_close(a, b, rtol=3e-2, atol=4e-2)
    _assert_argmax_where_clear(a, ref)
    # bit-reproducible: pieces are added in a fixed order whatever CTA arrives last
    a2 = _forced(m, emb, mask, forced, True)
    assert torch.equal(a, a2)


def test_rows_kernel_single_row(cuda_device):
    """`pcy_set_decode_megakernel(-1)` sends every row count to the tile-streaming kernel, 1 row included (greedy
    decoding is faster on the row-streaming kernel - 3.07 vs 3.50 ms per token at Llama-3-8B size - so this is not the
    default): logits of every step against the oracle and against the default greedy kernel."""
    from oracle.llama import random_llama_state_dict
    from procyon_b200 import _lib

    oc, pc = _cfgs("gq4wide")
    sd = random_llama_state_dict(oc, seed=3)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, 1, 140, seed=5)
    forced = torch.randint(0, oc.vocab, (1, 9), generator=torch.Generator().manual_seed(1))
    ref = _forced_oracle(sd, oc, emb, None, forced)
    lib = _lib.load()
    try:
        lib.pcy_set_decode_megakernel(-1)
        n0 = lib.pcy_launch_count()
        a = _forced(m, emb, None, forced, True)
        lib.pcy_set_decode_megakernel(1)
        b = _forced(m, emb, None, forced, True)
    finally:
        lib.pcy_set_decode_megakernel(1)
    assert torch.isfinite(a).all()
    torch.testing.assert_close(a, ref, rtol=3e-2, atol=4e-2)
    torch.testing.assert_close(a, b, rtol=3e-2, atol=4e-2)
    _assert_argmax_where_clear(a, ref)


@pytest.mark.parametrize("n,beams,S,pad,steps", [(1, 10, 300, 0, 7), (2, 6, 260, 9, 6), (1, 16, 129, 0, 5), (1, 5, 64, 0, 14),
                                                 (3, 4, 100, 11, 6), (1, 3, 1100, 0, 4),
                                                 # 4 inputs x 2 kv heads x 19 key splits = 152 attention items on 148 SMs:
                                                 # some CTAs take two items (a second Q staging, tile reuse)
                                                 (4, 4, 1200, 0, 3),
                                                 # 40 steps: the generated entries (10 x t keys) outgrow the prompt
                                                 (1, 10, 70, 0, 40)])
def test_rows_kernel_beam_steps_match_per_op_on_identical_state(cuda_device, n, beams, S, pad, steps):
    """Real beam-search state (ancestry through `slots`, keys shared between beams, this step's rows appended inside the
    kernel): after every selection step both paths run on the SAME session state - logits of every beam row must agree,
    and the K / V rows the persistent kernel appended must be the ones the per-op path writes."""
    from oracle.llama import random_llama_state_dict
    from procyon_b200 import _lib
    from procyon_b200.model.pmc_llama import SELECT_BEAM

    oc, pc = _cfgs("gq4", max_pos=2048)
    sd = random_llama_state_dict(oc, seed=13)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, n, S, seed=beams + S, pad_left=pad)
    am = mask.cuda() if pad else None
    lib = _lib.load()
    sess = m.get_session(n, beams, S, max(16, steps + 2), torch.device("cuda"), pad > 0, False)
    sel = torch.tensor([(i + 1) * S - 1 for i in range(n)], device="cuda", dtype=torch.int32)
    _, _, logits, valid = m.prefill(emb.cuda(), am, want_cache=True, want_hidden=False, sel_rows=sel,
                                    kv_out=sess.kv_prompt)
    if pad:
        sess.prompt_valid.copy_(valid)
    sess.reset(logits)
    group = max(1, beams // 2) if beams % 2 == 0 else beams
    try:
        for step in range(steps):
            sess.select(SELECT_BEAM, group, 0.8, -5, False)
            lib.pcy_set_decode_rows_megakernel(0)
            sess.forward()
            lb = sess.logits_cur.clone()
            kv_b = sess.kv_gen.clone()
            lib.pcy_set_decode_rows_megakernel(1)
            sess.forward()
            la = sess.logits_cur.clone()
            assert torch.isfinite(la).all(), f"step {step}"
            torch.testing.assert_close(la, lb, rtol=3e-2, atol=4e-2, msg=lambda t: f"step {step}: {t}")
            # (generation slots 0 .. step hold data; the rest of the cache is uninitialised memory)
            torch.testing.assert_close(sess.kv_gen[:, :, :, :step + 1].float(), kv_b[:, :, :, :step + 1].float(),
                                       rtol=3e-2, atol=3e-2, msg=lambda t: f"step {step} kv: {t}")
    finally:
        lib.pcy_set_decode_rows_megakernel(1)


@pytest.mark.parametrize("n,beams,group,S,pad", [(1, 10, 2, 150, 0), (2, 5, 5, 90, 6), (1, 4, 2, 200, 0)])
def test_rows_kernel_beam_generation_matches_oracle(cuda_device, n, beams, group, S, pad):
    """Whole diverse-beam generations through the captured step graph (persistent kernel + selection kernels): every
    beam identical to the oracle's on all steps before the oracle's first sub-noise decision margin, equal scores for
    the sequences both found, and identical results from the per-op path."""
    from oracle.generate import generate_beam_search as oracle_beam
    from oracle.llama import random_llama_state_dict
    from procyon_b200 import _lib
    from procyon_b200.model.generation import generate_beam_search
    from test_gpu_llama import _clear_steps

    oc, pc = _cfgs("gq4", max_pos=512)
    sd = random_llama_state_dict(oc, seed=17)
    m = _build(sd, pc)
    ids, emb, mask = _inputs(oc, sd, n, S, seed=beams * 3 + S, pad_left=pad)
    am = mask.cuda() if pad else None
    kw = dict(max_len=12, beam_size=beams, beam_group_size=group, diversity_penalty=0.8, eos_token_id=-5)
    o1, lp1, lg1 = generate_beam_search(m, emb.cuda(), am, **kw)
    tr = []
    ro, rlp, rlogits = oracle_beam(sd, oc, emb.float(), mask if pad else torch.ones_like(mask), max_len=12,
                                   beam_size=beams, beam_group_size=group, diversity_penalty=0.8, eos_id=-5,
                                   act_round="bf16", mask_pads_in_decode=True, trace=tr)
    clear = _clear_steps(tr)
    assert clear >= 0
    assert torch.equal(o1[..., :clear].cpu(), ro[..., :clear]), f"beams differ within the {clear} clear-margin steps"
    for i in range(n):
        ours = {tuple(o1[i, b].tolist()): float(lp1[i, b]) for b in range(beams)}
        ref = {tuple(ro[i, b].tolist()): float(rlp[i, b]) for b in range(beams)}
        for k in set(ours) & set(ref):
            assert abs(ours[k] - ref[k]) < 6e-2 + 1e-2 * abs(ref[k])
    lib = _lib.load()
    try:
        lib.pcy_set_decode_rows_megakernel(0)
        m.__dict__.pop("_sessions", None)  # (cached sessions hold the step graph captured with the other setting)
        o2, lp2, lg2 = generate_beam_search(m, emb.cuda(), am, **kw)
    finally:
        lib.pcy_set_decode_rows_megakernel(1)
    assert torch.equal(o1[..., :clear], o2[..., :clear])

"""ESM2 encoder + pooler on the GPU vs the CPU oracle (same seeded weights and tokens)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(num_params, sd, pooling="mean", correction=False, custom=None, max_len=1024):
    from procyon_b200.model.esm import ESM_PLM

    m = ESM_PLM(num_params=num_params, pooling_method=pooling, protein_pooling_correction_option=correction,
                custom_config=custom, max_protein_len=max_len)
    missing = m.model.load_state_dict(sd, strict=True)
    return m.cuda()


@pytest.mark.parametrize("cfg", [
    dict(name="tiny16", custom=(2, 64, 4), lengths=[20, 9, 31, 1]),        # head_dim 16
    dict(name="tiny24", custom=(2, 96, 4), lengths=[50, 70, 3]),           # head_dim 24 (ESM2-35M style)
    dict(name="tiny64", custom=(3, 256, 4), lengths=[130, 64, 100, 129]),  # head_dim 64 (650M style)
])
def test_esm_states_and_pool(cuda_device, cfg):
    from oracle import esm2 as O

    L, d, H = cfg["custom"]
    sd = O.random_esm_state_dict(L, d, seed=7)
    toks = O.random_protein_tokens(len(cfg["lengths"]), 0, seed=3, lengths=cfg["lengths"])
    toks[0, 3] = O.MASK_IDX  # exercise the mask-dropout rescale
    ref = O.esm2_forward(sd, toks, L, H, act_round="bf16")
    m = _build("custom", sd, custom=cfg["custom"])
    got = m.encode_tokens(toks.cuda()).float().cpu()
    nonpad = toks != O.PAD_IDX
    # bf16 activations through L layers: tolerance = a few bf16 ulps of O(1) LayerNorm outputs
    torch.testing.assert_close(got[nonpad], ref[nonpad], rtol=3e-2, atol=3e-2)
    for pooling, corr in (("mean", False), ("mean", True), ("max", False)):
        m.pooler.pooling_method = pooling
        m.pooler.protein_pooling_correction_option = corr
        z, logits = m(toks.cuda())
        assert logits is None
        refp = O.protein_pooler(ref, torch.arange(toks.shape[0]), ~nonpad, pooling, corr)
        torch.testing.assert_close(z.float().cpu(), refp, rtol=3e-2, atol=3e-2)


def test_esm_config1_35m(cuda_device):
    """BASELINE config 1: ESM2-35M encode of 16 synthetic 256-residue proteins, pooled (mean)."""
    from oracle import esm2 as O

    L, d, H = O.ESM_SIZES["35m"]
    sd = O.random_esm_state_dict(L, d, seed=0)
    toks = O.random_protein_tokens(16, 256, seed=1234)
    ref = O.esm_plm_forward(sd, toks, L, H, pooling="mean", act_round="bf16")
    m = _build("35m", sd, pooling="mean")
    z, _ = m(toks.cuda())
    assert z.shape == (16, d)
    torch.testing.assert_close(z.float().cpu(), ref, rtol=3e-2, atol=2e-2)
    # determinism + micro-batching invariance: pooled rows do not depend on the pass they were encoded in
    m.max_tokens_per_pass = 3 * toks.shape[1]
    z2, _ = m(toks.cuda())
    assert torch.equal(z, z2)


def test_esm_long_protein_split(cuda_device):
    """Proteins longer than max_protein_len are chunked into extra rows and re-pooled across chunks."""
    from oracle import esm2 as O

    L, d, H = 2, 64, 4
    sd = O.random_esm_state_dict(L, d, seed=11)
    toks = O.random_protein_tokens(3, 0, seed=5, lengths=[100, 30, 75])
    for pooling in ("mean", "max"):
        ref = O.esm_plm_forward(sd, toks, L, H, pooling=pooling, max_protein_len=32, act_round="bf16")
        m = _build("custom", sd, pooling=pooling, custom=(L, d, H), max_len=32)
        z, _ = m(toks.cuda())
        assert z.shape == ref.shape == (3, d)
        torch.testing.assert_close(z.float().cpu(), ref, rtol=3e-2, atol=2e-2)
        zs, _ = m(toks.cuda(), aggregate=False)
        refs = O.esm_plm_forward(sd, toks, L, H, max_protein_len=32, act_round="bf16", aggregate=False)
        assert zs.shape == refs.shape
        keep = (toks != O.PAD_IDX)[:, : refs.shape[1]]
        torch.testing.assert_close(zs.float().cpu()[keep], refs[keep], rtol=3e-2, atol=3e-2)


def test_esm_empty_batch(cuda_device):
    from procyon_b200.model.esm import ESM_PLM

    m = ESM_PLM(num_params="custom", custom_config=(1, 64, 4), pooling_method="mean").cuda()
    out = m.encode_tokens(torch.zeros((0, 10), dtype=torch.int64, device="cuda"))
    assert out.shape == (0, 10, 64)


@pytest.mark.parametrize("lengths", [[298, 131, 260], [512, 300, 130, 64]])
def test_esm_tcgen05_attention_matches_mma_sync_and_oracle(cuda_device, lengths):
    """head_dim 64: the three tcgen05 kernels (64-key steps with Q and P in TMEM; 64-key steps with double-buffered
    S / P / P.V; 128-key steps) against the mma.sync kernel and the oracle, with padded and un-padded rows, key tiles that end mid-tile (T = 300: 44 keys
    in the last step; T = 514: 2 keys) and the shifted last query tile whose repeated rows are skipped."""
    from oracle import esm2 as O
    from procyon_b200 import _lib

    L, d, H = 2, 256, 4
    sd = O.random_esm_state_dict(L, d, seed=21)
    toks = O.random_protein_tokens(len(lengths), 0, seed=9, lengths=lengths)
    m = _build("custom", sd, custom=(L, d, H))
    lib = _lib.load()
    outs = {}
    try:
        for steps64 in (7, 6, 5, 4, 3, 2, 1, 0):
            lib.pcy_set_esm_tc_attention(1)
            lib.pcy_set_esm_attention_kernel(steps64)
            lib.pcy_set_fused_rope(1 if steps64 == 0 else 0)
            outs[steps64] = m.encode_tokens(toks.cuda()).float().cpu()
            # deterministic (kernel 5 first failed exactly here: a softmax thread that skipped phases of the P.V barrier
            # took "P.V(n-2) still running" for "P.V(n-1) done" in rows whose last key step is a fast, fully masked one)
            for _ in range(3 if steps64 >= 5 else 1):
                again = m.encode_tokens(toks.cuda()).float().cpu()
                assert torch.equal(outs[steps64], again)
            if steps64 in (2, 4, 5):  # Q rotated inside the attention kernel vs by the RoPE pass (default): same arithmetic
                lib.pcy_set_esm_attention_q_rope(1)
                inside = m.encode_tokens(toks.cuda()).float().cpu()
                lib.pcy_set_esm_attention_q_rope(0)
                torch.testing.assert_close(outs[steps64][toks != O.PAD_IDX], inside[toks != O.PAD_IDX], rtol=1e-2,
                                           atol=1e-2)
        # the rows beyond the last full 128-row tile on the mma.sync kernel instead of one more tcgen05 CTA
        lib.pcy_set_esm_attention_kernel(5)
        lib.pcy_set_esm_attention_tail_rows(16)
        tail = m.encode_tokens(toks.cuda()).float().cpu()
        lib.pcy_set_esm_attention_tail_rows(0)
        lib.pcy_set_esm_tc_attention(0)
        lib.pcy_set_fused_rope(0)
        b = m.encode_tokens(toks.cuda()).float().cpu()
    finally:
        lib.pcy_set_esm_attention_tail_rows(0)
        lib.pcy_set_esm_tc_attention(1)
        lib.pcy_set_esm_attention_kernel(5)  # the default
        lib.pcy_set_esm_attention_q_rope(0)  # the default
        lib.pcy_set_fused_rope(0)
    nonpad = toks != O.PAD_IDX
    ref = O.esm2_forward(sd, toks, L, H, act_round="bf16")
    outs["tail"] = tail
    for steps64, a in outs.items():
        torch.testing.assert_close(a[nonpad], b[nonpad], rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(a[nonpad], ref[nonpad], rtol=3e-2, atol=3e-2)


def test_return_mlm_logits(cuda_device):
    """`return_mlm` path: residue states + masked-LM head, incl. proteins longer than
    max_protein_len (chunk -> encode -> stitch back for states AND logits).  Golden: HF EsmForMaskedLM."""
    import os

    from oracle import esm2 as O

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "hf_esm_lm_head.pt"), weights_only=False)
    L, d, H = g["n_layers"], g["d"], g["n_heads"]
    m = _build("custom", {k: v.bfloat16() for k, v in g["state_dict"].items()}, custom=(L, d, H))
    z, logits = m(g["tokens"].cuda(), aggregate=False)
    keep = g["tokens"] != O.PAD_IDX
    assert logits.shape == g["logits"].shape and logits.dtype == z.dtype
    # bf16 weights + activations vs the fp32 HF run: a few bf16 ulps of O(1) values
    torch.testing.assert_close(logits.float().cpu()[keep], g["logits"][keep], rtol=5e-2, atol=5e-2)
    # oracle with the same roundings, longer-than-max proteins
    sd = O.random_esm_state_dict(2, 64, seed=11)
    toks = O.random_protein_tokens(3, 0, seed=4, lengths=[70, 30, 45])
    m2 = _build("custom", sd, custom=(2, 64, 4), max_len=32)
    z2, lg2 = m2(toks.cuda(), aggregate=False)
    bt, keys, eos = O.batched_split_long_seq(toks, max_protein_len=32)
    zr = O.esm2_forward(sd, bt, 2, 4, act_round="bf16")
    lr = O.esm2_lm_head(sd, zr, act_round="bf16")
    zr, lr = O.reverse_batched_split(zr, keys, eos), O.reverse_batched_split(lr, keys, eos)
    keep2 = toks != O.PAD_IDX
    assert lg2.shape == lr.shape
    torch.testing.assert_close(z2.float().cpu()[keep2], zr[keep2], rtol=3e-2, atol=3e-2)
    # logits are O(5) here (embedding rows of std 0.5 against LayerNorm outputs): the 3e-2 state tolerance, carried
    # through dense -> gelu -> LayerNorm -> 64-term dot products, is a few % of the logit scale at worst and far
    # less on average
    scale = lr[keep2].abs().max().item()
    err = (lg2.float().cpu()[keep2] - lr[keep2]).abs()
    assert err.max().item() < 5e-2 * scale and err.mean().item() < 5e-3 * scale, (err.max().item(), err.mean().item(), scale)

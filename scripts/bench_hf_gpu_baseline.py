"""GPU reference baseline — the denominator of BASELINE.json's ">= 10x the reference GPU (HF+ESM) forward" target
(BASELINE.md "Measurement plan" item 2): the reference's phenotype-generation path restated with the STOCK
HuggingFace classes it is built on, in bf16 on one GPU, keeping the reference's loop semantics:

  * ESM2-650M-shaped `EsmModel` (eager attention) on one protein -> mean pool over all tokens -> 3-layer MLP projector
    (procyon/model/esm.py:504-541, model_utils.py:13-41);
  * `LlamaForCausalLM` (Llama-3-8B shape, V = 128263, eager attention, RoPE base 10000 as under transformers 4.31),
    prefill on `inputs_embeds` with `output_hidden_states=True` (pmc_llama.py:575,584), then one forward per generated
    token with the growing `past_key_values`, `logits[:, -1].cpu()` and the argmax on the host every step
    (model_unified.py:769-773, 887-897).

Random-init weights (no checkpoints offline), synthetic inputs of the bench's shapes (1024 residues, 1024-token
prompt, 128 generated tokens).  Prints one JSON line.  `--tiny` runs a small shape (CPU works) to check the script.

    python scripts/bench_hf_gpu_baseline.py [--steps 3] [--warmup 1] [--gen 128]
"""
import argparse
import json
import time

import torch


def build(tiny: bool, device, dtype):
    from transformers import EsmConfig, EsmModel, LlamaConfig, LlamaForCausalLM

    if tiny:
        e = dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256)
        l = dict(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, max_position_embeddings=512)
    else:
        e = dict(hidden_size=1280, num_hidden_layers=33, num_attention_heads=20, intermediate_size=5120)
        l = dict(vocab_size=128263, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192)
    ecfg = EsmConfig(vocab_size=33, mask_token_id=32, pad_token_id=1, position_embedding_type="rotary",
                     token_dropout=True, emb_layer_norm_before=False, layer_norm_eps=1e-5, attn_implementation="eager",
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **e)
    lcfg = LlamaConfig(rms_norm_eps=1e-5, rope_parameters={"rope_type": "default", "rope_theta": 10000.0},
                       attn_implementation="eager", tie_word_embeddings=False, **l)
    torch.manual_seed(0)
    with torch.device(device):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            esm = EsmModel(ecfg, add_pooling_layer=False).eval()
            llama = LlamaForCausalLM(lcfg).eval()
            d_e, d_t = e["hidden_size"], l["hidden_size"]
            proj = torch.nn.Sequential(torch.nn.Linear(d_e, d_e), torch.nn.ReLU(), torch.nn.Dropout(0.0),
                                       torch.nn.Linear(d_e, d_e), torch.nn.ReLU(), torch.nn.Dropout(0.0),
                                       torch.nn.Linear(d_e, d_t)).eval()
        finally:
            torch.set_default_dtype(prev)
    for m in (esm, llama, proj):
        for p in m.parameters():
            if p.dim() > 1:
                p.data.normal_(std=0.02)
    return esm, llama, proj, l["vocab_size"]


@torch.no_grad()
def generate(esm, llama, proj, protein, prompt_ids, soft_pos, gen):
    """One pass of the reference's greedy generate loop; returns the generated ids (host list)."""
    z = esm(input_ids=protein, attention_mask=torch.ones_like(protein)).last_hidden_state  # (1, T, d)
    pooled = z.mean(dim=1)                                                               # CLS / EOS included
    soft = proj(pooled)
    x = llama.get_input_embeddings()(prompt_ids)
    x[0, soft_pos] = soft[0].to(x.dtype)                                                 # soft-token splice
    out = llama(inputs_embeds=x, use_cache=True, output_hidden_states=True)
    past = out.past_key_values
    tok = int(out.logits[:, -1].float().cpu().argmax(-1))                                 # per-step host round trip
    toks = [tok]
    for _ in range(gen - 1):
        out = llama(input_ids=torch.tensor([[tok]], device=prompt_ids.device), past_key_values=past, use_cache=True,
                    output_hidden_states=True)
        past = out.past_key_values
        tok = int(out.logits[:, -1].float().cpu().argmax(-1))
        toks.append(tok)
    return toks


def _layer_kv(cache, l):
    """(keys, values) tensors of layer l of whatever cache object this transformers version returns (4.31: tuple of
    tuples; 5.x: DynamicCache with .layers[l].keys/.values)."""
    if hasattr(cache, "layers"):
        return cache.layers[l].keys, cache.layers[l].values
    return cache[l][0], cache[l][1]


@torch.no_grad()
def generate_beam(esm, llama, proj, protein, prompt_ids, soft_pos, gen, beam_size=10, beam_group_size=2,
                  diversity_penalty=0.8, eos_id=-1):
    """The reference's `_generate_beam_search` loop (procyon/model/model_unified.py:701-842) restated on the stock HF
    modules, statement for statement: prompt embeddings repeated `beam_size` times for the prefill (:751-752), one
    forward per token with `output_hidden_states=True` (pmc_llama.py:575,584), `logits.clone().cpu()` appended to a
    growing host tensor every step (:773-781), LogSoftmax + per-group bincount penalty + ravel().topk on the device
    (:783-822), re-indexing of `out`, the host logits history (:827) and every layer's K and V (:830-832) per group,
    `.item()` EOS check per step (:833).  beam_size / beam_group_size default to the evaluation framework's
    (procyon/evaluate/framework/procyon.py:71-76: 5 captions x group size 2 = 10 beams)."""
    device = prompt_ids.device
    z = esm(input_ids=protein, attention_mask=torch.ones_like(protein)).last_hidden_state
    soft = proj(z.mean(dim=1))
    x = llama.get_input_embeddings()(prompt_ids)
    x[0, soft_pos] = soft[0].to(x.dtype)
    n = x.shape[0]
    bb = n * beam_size
    V = llama.config.vocab_size
    groups = beam_size // beam_group_size
    embeds = torch.repeat_interleave(x, repeats=beam_size, dim=0)
    mask = torch.ones(embeds.shape[:2], dtype=torch.int64, device=device)
    cur = torch.zeros((bb,), device=device)
    out = torch.zeros(bb, gen, dtype=torch.int64, device=device)
    past, output_logits = None, None
    sm = torch.nn.LogSoftmax(dim=-1)
    n_layers = llama.config.num_hidden_layers
    for i in range(gen):
        if i == 0:
            o = llama(inputs_embeds=embeds, attention_mask=mask, use_cache=True, output_hidden_states=True)
        else:
            o = llama(input_ids=out[:, i - 1].unsqueeze(-1), past_key_values=past, use_cache=True,
                      output_hidden_states=True)
        logits = o.logits[:, -1, :]
        past = o.past_key_values
        it = logits.detach().clone().cpu().unsqueeze(1)
        output_logits = it if output_logits is None else torch.cat([output_logits, it], dim=1)
        log_probs = sm(logits.float()) + cur[:, None]
        for inp in range(n):
            b0 = inp * beam_size
            for g in range(groups):
                inc = 1 if i == 0 else beam_group_size
                gs = b0 + g * beam_group_size
                ge = gs + beam_group_size
                lp = log_probs[gs:gs + inc]
                if g != 0:
                    lp -= diversity_penalty * torch.bincount(out[b0:gs, i], minlength=V).to(device)
                vals, idx = lp.ravel().topk(beam_group_size)
                toks = idx % V
                src = (idx // V) + gs
                out[gs:ge] = out[src]
                out[torch.arange(gs, ge), i] = toks
                cur[gs:ge] = vals
                output_logits[gs:ge] = output_logits[src.cpu()]
                for l in range(n_layers):
                    k, v = _layer_kv(past, l)
                    k[gs:ge] = k[src]
                    v[gs:ge] = v[src]
        if torch.all((out == eos_id).any(dim=1)).item():
            break
    return out.cpu().unflatten(0, (n, beam_size)), cur.cpu().unflatten(0, (n, beam_size)), output_logits


def measure(steps=2, warmup=1, gen=128, n_res=1024, n_prompt=1024, beam_size=10, beam_group_size=2, esm_proteins=256,
            esm_residues=512, tiny=False, device=None, beam_steps=1):
    """Everything bench.py's `gpu_reference` object needs, as one dict (all times wall-clock with a device
    synchronize on both sides: the loops are host-driven, so that IS their cost)."""
    cuda = torch.cuda.is_available()
    device = device or (torch.device("cuda", 0) if cuda else torch.device("cpu"))
    dtype = torch.bfloat16 if device.type == "cuda" else torch.float32
    esm, llama, proj, vocab = build(tiny, device, dtype)
    g = torch.Generator().manual_seed(1234)
    protein = torch.cat([torch.tensor([0]), torch.randint(4, 24, (n_res,), generator=g), torch.tensor([2])])[None].to(device)
    prompt = torch.randint(0, min(vocab, 128000), (1, n_prompt), generator=torch.Generator().manual_seed(4321)).to(device)

    def sync():
        if device.type == "cuda":
            torch.cuda.synchronize(device)

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        sync()
        t0 = time.perf_counter()
        for _ in range(k):
            r = fn()
        sync()
        return (time.perf_counter() - t0) / k, r

    res = {"impl": "hf-eager", "transformers": __import__("transformers").__version__, "dtype": str(dtype),
           "shape": {"residues": n_res, "prompt_tokens": n_prompt, "gen": gen}}
    dt, toks = timed(lambda: generate(esm, llama, proj, protein, prompt, 16, gen), steps, warmup)
    res["greedy"] = {"tokens_per_s": gen / dt, "ms_per_generate": dt * 1e3, "n_generated": len(toks)}
    # prefill alone and one decode step alone (same calls as inside the loop)
    with torch.no_grad():
        x = llama.get_input_embeddings()(prompt)
        dt_p, o = timed(lambda: llama(inputs_embeds=x, use_cache=True, output_hidden_states=True), 2, 1)
        res["greedy"]["ms_prefill"] = dt_p * 1e3
        res["greedy"]["ms_per_decode_step"] = (dt - dt_p) * 1e3 / max(gen - 1, 1)
    del o
    dt, (out, lp, lg) = timed(lambda: generate_beam(esm, llama, proj, protein, prompt, 16, gen, beam_size,
                                                     beam_group_size), beam_steps, 0 if not tiny else 0)
    res["beam"] = {"beam_size": beam_size, "beam_group_size": beam_group_size, "ms_per_generate": dt * 1e3,
                   "tokens_per_s_aggregate": beam_size * gen / dt, "sequences_tokens_per_s": gen / dt,
                   "steps_run": int(lg.shape[1])}
    del lg
    micro = 4 if tiny else 32
    pps, ms, shape = esm_batch(esm, proj, esm_proteins, esm_residues, device, max(steps, 1), warmup, micro=micro)
    res["esm2_encode"] = {"proteins_per_s": pps, "ms_per_step": ms, "proteins": esm_proteins, "residues": esm_residues,
                          "micro_batch": micro}
    del esm, llama, proj
    if device.type == "cuda":
        torch.cuda.empty_cache()
    return res


@torch.no_grad()
def esm_batch(esm, proj, n_prot, n_res, device, steps, warmup, micro=32):
    """Batch encode the way the reference's evaluation loop does it (evaluate/framework/procyon.py:296-321): fixed-size
    batches through the encoder, mean pool, projector; returns proteins/s."""
    g = torch.Generator().manual_seed(1234)
    toks = torch.full((n_prot, n_res + 2), 1, dtype=torch.int64)
    toks[:, 0] = 0
    toks[:, 1:n_res + 1] = torch.randint(4, 24, (n_prot, n_res), generator=g)
    toks[:, n_res + 1] = 2
    toks = toks.to(device)

    def run():
        outs = []
        for i in range(0, n_prot, micro):
            t = toks[i:i + micro]
            z = esm(input_ids=t, attention_mask=torch.ones_like(t)).last_hidden_state
            outs.append(proj(z.mean(dim=1)).float().cpu())
        return torch.cat(outs)

    for _ in range(warmup):
        run()
    if device.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run()
    if device.type == "cuda":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return n_prot / dt, dt * 1e3, tuple(out.shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--all", action="store_true", help="greedy + reference beam loop + ESM batch as ONE json object")
    args = ap.parse_args()
    if args.all:
        kw = dict(gen=6, n_res=64, n_prompt=48, esm_proteins=8, esm_residues=30, beam_size=4) if args.tiny else \
            dict(gen=args.gen)
        print(json.dumps(measure(steps=args.steps, warmup=args.warmup, tiny=args.tiny, **kw)))
        return
    cuda = torch.cuda.is_available()
    if not cuda and not args.tiny:
        raise SystemExit("no CUDA device: use --tiny for the CPU check of the script")
    device = torch.device("cuda", 0) if cuda else torch.device("cpu")
    dtype = torch.bfloat16 if cuda else torch.float32
    esm, llama, proj, vocab = build(args.tiny, device, dtype)
    n_res, n_prompt, gen = (64, 48, 6) if args.tiny else (1024, 1024, args.gen)
    g = torch.Generator().manual_seed(1234)
    protein = torch.cat([torch.tensor([0]), torch.randint(4, 24, (n_res,), generator=g), torch.tensor([2])])[None].to(device)
    prompt = torch.randint(0, min(vocab, 128000), (1, n_prompt), generator=torch.Generator().manual_seed(4321)).to(device)
    for _ in range(args.warmup):
        generate(esm, llama, proj, protein, prompt, 16, gen)
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        toks = generate(esm, llama, proj, protein, prompt, 16, gen)
    if cuda:
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"impl": "hf_eager_gpu_reference" if cuda else "hf_eager_cpu_tiny_check",
                      "metric": "phenotype_gen_tokens_per_s", "value": gen / dt, "unit": "tokens/s",
                      "ms_per_step": dt * 1e3, "steps": args.steps, "warmup": args.warmup, "dtype": str(dtype),
                      "config": {"workload": f"HF EsmModel ({n_res} residues) + projector + HF LlamaForCausalLM eager: "
                                             f"prefill S={n_prompt} + {gen} greedy tokens, reference loop semantics",
                                 "tiny": args.tiny}, "n_generated": len(toks)}))
    n_prot, n_res2 = (8, 30) if args.tiny else (256, 512)
    pps, ms, shape = esm_batch(esm, proj, n_prot, n_res2, device, args.steps, args.warmup, micro=4 if args.tiny else 32)
    print(json.dumps({"impl": "hf_eager_gpu_reference" if cuda else "hf_eager_cpu_tiny_check",
                      "metric": "esm2_encode_proteins_per_s", "value": pps, "unit": "proteins/s", "ms_per_step": ms,
                      "dtype": str(dtype), "config": {"workload": f"HF EsmModel eager, {n_prot} proteins x {n_res2} "
                                                                  "residues in batches, mean pool + projector",
                                                      "tiny": args.tiny}, "out_shape": list(shape)}))


if __name__ == "__main__":
    main()

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
    args = ap.parse_args()
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

"""How far are the CUDA path's per-step log-probs from the oracle's on the tiny test models (teacher-forced), overall
and on the top-20 candidates of every row (the ones beam search decides between)?  Sets the margin the strict beam
tests rely on."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle.llama import llama_forward
from test_gpu_beam_strict import _inputs, _tiny

for kind in ("gq2", "gq4"):
    for structured in (False, True):
        oc, sd, m = _tiny(kind, structured_head=structured)
        errs, toperrs, gaps = [], [], []
        for seed in range(6):
            rows, S, steps = 4, 24, 6
            ids, emb, mask = _inputs(oc, sd, rows, S, seed=seed)
            forced = torch.randint(0, oc.vocab, (rows, steps), generator=torch.Generator().manual_seed(seed))
            out = m(input_embeds=emb.cuda(), use_cache=True)
            sess = out.past_key_values
            ours = [sess.logits_cur.clone().cpu()]
            for i in range(steps):
                ours.append(m(input_ids=forced[:, i:i + 1].cuda(), past_key_values=sess).logits[:, 0].cpu())
            r = llama_forward(sd, oc, inputs_embeds=emb.float(), act_round="bf16")
            refs = [r["logits"][:, -1]]
            for i in range(steps):
                r = llama_forward(sd, oc, input_ids=forced[:, i:i + 1], past=r["past"], act_round="bf16")
                refs.append(r["logits"][:, -1])
            for a, b in zip(ours, refs):
                la, lb = torch.log_softmax(a, -1), torch.log_softmax(b, -1)
                errs.append((la - lb).abs().max().item())
                top = lb.topk(20, dim=-1)
                toperrs.append((la.gather(1, top.indices) - top.values).abs().max().item())
                gaps.append((top.values[:, :-1] - top.values[:, 1:]).median().item())
        e, t, g = torch.tensor(errs), torch.tensor(toperrs), torch.tensor(gaps)
        print(f"{kind} structured_head={structured}: log-prob |err| all tokens max {e.max():.4f} median {e.median():.4f}; "
              f"top-20 candidates max {t.max():.4f} median {t.median():.4f}; median gap between successive top-20 "
              f"{g.median():.3f}; top-5 logits {refs[-1][0].topk(5).values.tolist()}", flush=True)

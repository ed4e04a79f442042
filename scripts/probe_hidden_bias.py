"""Is the CUDA path's deviation from the oracle random (bf16 flips) or systematic (a scale)?  Compares final hidden
states (norm ratio, cosine) and the top logits on the structured-head tiny model, prefill and decode."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle.llama import llama_forward
from test_gpu_beam_strict import _inputs, _tiny

oc, sd, m = _tiny("gq4", structured_head=True)
ids, emb, mask = _inputs(oc, sd, 2, 24, seed=5)
ref = llama_forward(sd, oc, inputs_embeds=emb.float(), act_round="bf16")
out = m(input_embeds=emb.cuda(), attn_masks=None)
h = out.hidden_states[-1].float().cpu()
hr = ref["hidden_states"][-1]
print("prefill hidden: norm ratio", (h.norm(dim=-1) / hr.norm(dim=-1)).flatten()[-6:].tolist())
print("prefill hidden: 1-cos", (1 - torch.nn.functional.cosine_similarity(h, hr, dim=-1)).flatten()[-6:].tolist())
print("prefill hidden: frac elements equal", (h == hr).float().mean().item(), "max abs diff", (h - hr).abs().max().item())
lg = out.logits.cpu()
top = ref["logits"][:, -1].topk(5)
print("prefill top5 ref", top.values[0].tolist(), "ours", lg[:, -1].gather(1, top.indices)[0].tolist())
# lm head alone on the ORACLE's hidden: isolates the head GEMM
lg2 = m.lm_head_logits(hr[:, -1].bfloat16().cuda()).cpu()
print("head-only top5 ours", lg2.gather(1, top.indices)[0].tolist())
# decode step
o0 = m(input_embeds=emb.cuda(), use_cache=True)
sess = o0.past_key_values
print("prefill(sel rows, fused norm+head) top5", sess.logits_cur.cpu().gather(1, top.indices)[0].tolist())
nxt = torch.tensor([[5], [77]])
r1 = llama_forward(sd, oc, input_ids=nxt, past=ref["past"], act_round="bf16")
o1 = m(input_ids=nxt.cuda(), past_key_values=sess)
t1 = r1["logits"][:, -1].topk(5)
print("decode top5 ref", t1.values[0].tolist(), "ours", o1.logits[:, 0].cpu().gather(1, t1.indices)[0].tolist())

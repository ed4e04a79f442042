"""Diagnostic (needs the temporary STAMP build of esm_attention_row_kernel): clock64 stamps of
the MMA thread and of one softmax thread of one CTA of kernel 6 inside a real ESM2-650M encode."""
import ctypes, json, sys
import torch
sys.path.insert(0, ".")
from procyon_b200 import _lib
from procyon_b200.model.esm import ESM_PLM

dev = torch.device("cuda", 0)
torch.manual_seed(0)
with torch.device(dev):
    m = ESM_PLM(num_params="650m", pooling_method="mean").bfloat16().eval()
for p in m.parameters():
    if p.dim() > 1:
        p.data.normal_(std=0.02)
N = 128
g = torch.Generator().manual_seed(1)
toks = torch.full((N, 514), 1, dtype=torch.int64)
toks[:, 0] = 0
toks[:, 1:513] = torch.randint(4, 24, (N, 512), generator=g)
toks[:, 513] = 2
toks = toks.to(dev)
lib = _lib.load()
lib.pcy_set_esm_attention_kernel(6)
for _ in range(2):
    m(toks)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
assert lib.pcy_debug_read_attn_stamps(buf) == 0
st = [[[buf[(k * 16 + s) * 8 + h] for h in range(8)] for s in range(16)] for k in range(2)]
t0 = st[0][0][0]
out = {"mma": [], "softmax": []}
for j in range(9):
    a = st[0][j]
    out["mma"].append({"step": j, "start": a[0] - t0, "issue_s_next": a[1] - a[0], "to_p_ready_wait": a[2] - a[1],
                       "p_ready_wait": a[3] - a[2], "pv_issue": a[4] - a[3]})
    b = st[1][j]
    out["softmax"].append({"step": j, "start": b[0] - t0, "s_full_wait": b[1] - b[0], "tmem_ld": b[2] - b[1],
                           "max_exp_store_issue": b[3] - b[2], "o_full_wait": b[4] - b[3], "wait_st": b[5] - b[4],
                           "arrive": b[6] - b[5]})
for k in ("mma", "softmax"):
    for row in out[k]:
        print(k, json.dumps(row))

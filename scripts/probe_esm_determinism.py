"""Does the ESM2 encode read memory it did not write?  Encode, poison freed GPU memory with NaN / huge patterns, encode
again with a fresh workspace, compare bit for bit; report where the oracle error concentrates."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import esm2 as OE
from procyon_b200.model.esm import ESM_PLM

L, d, H = OE.ESM_SIZES["650m"]
sd = OE.random_esm_state_dict(L, d, seed=6, dtype=torch.bfloat16)
toks = OE.random_protein_tokens(2, 512, seed=22, lengths=[512, 130])


def enc_once(poison):
    if poison is not None:
        junk = [torch.full((256 * 1024 * 1024,), poison, device="cuda", dtype=torch.bfloat16) for _ in range(8)]
        torch.cuda.synchronize()
        del junk
    enc = ESM_PLM(num_params="650m", pooling_method="mean", max_protein_len=1024)
    enc.model.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval()
    out = enc.encode_tokens(toks.cuda()).float().cpu()
    enc.release()
    return out


a = enc_once(None)
b = enc_once(float("nan"))
c = enc_once(3e38)
keep = toks != 1
print("finite:", torch.isfinite(a[keep]).all().item(), torch.isfinite(b[keep]).all().item(), torch.isfinite(c[keep]).all().item())
print("bit-identical after NaN poison:", torch.equal(a[keep], b[keep]), " after 3e38 poison:", torch.equal(a[keep], c[keep]))
print("max |a-b| valid rows:", (a[keep] - b[keep]).abs().max().item() if torch.isfinite(b[keep]).all() else "nan present")
ref = OE.esm2_forward(sd, toks, L, H, act_round="bf16")
for p in range(2):
    k = keep[p]
    err = (a[p][k] - ref[p][k]).abs()
    i = err.argmax()
    r, col = divmod(int(i), d)
    print(f"protein {p}: max err {err.max():.4f} at token {r} dim {col} (ref value {ref[p][k][r, col]:.3f}); "
          f"mean err {err.mean():.5f}; |ref| max {ref[p][k].abs().max():.2f}")

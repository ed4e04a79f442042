"""Diagnostic: attention time inside a real ESM2-650M batch encode (256 x 512 residues) for the kernels named in
PCY_ESM_ATTN (comma-separated).  Used at the end of round 2 with two temporary edits of esm_attention_row_kernel
(max32 / exp32 replaced by a bit pack; one P.V MMA per step instead of four) to split the kernel's time into softmax
arithmetic, MMA issue and the TMEM / mbarrier skeleton: DESIGN.md section 5."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, ".")
from procyon_b200 import _lib
from procyon_b200.model.esm import ESM_PLM

def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        m = ESM_PLM(num_params="650m", pooling_method="mean").bfloat16().eval()
    for p in m.parameters():
        if p.dim() > 1:
            p.data.normal_(std=0.02)
    N = 256
    g = torch.Generator().manual_seed(1)
    toks = torch.full((N, 514), 1, dtype=torch.int64)
    toks[:, 0] = 0
    toks[:, 1:513] = torch.randint(4, 24, (N, 512), generator=g)
    toks[:, 513] = 2
    toks = toks.to(dev)
    lib = _lib.load()
    for kern in [int(k) for k in os.environ.get("PCY_ESM_ATTN", "6,5").split(",")]:
        lib.pcy_set_esm_attention_kernel(kern)
        for _ in range(2):
            m(toks)
        lib.pcy_esm_profile(1)
        for _ in range(3):
            m(toks)
        buf = (ctypes.c_double * 8)()
        lib.pcy_esm_profile_read(buf, 8)
        lib.pcy_esm_profile(0)
        print(json.dumps({"kernel": kern, "attention_ms": round(buf[4] / 3, 2), "qkv_ms": round(buf[2] / 3, 2), "fc2_ms": round(buf[7] / 3, 2)}), flush=True)
    lib.pcy_set_esm_attention_kernel(5)

if __name__ == "__main__":
    main()

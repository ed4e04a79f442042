"""One ESM2-650M encode pass (default 64 proteins x 512 residues) for ncu: only the encoder is built.

    ncu --set full --clock-control none --import-source on -k regex:esm_attention_tc64 -c 1 -o gpurun_out/prof_attn64 -f \
        python scripts/profile_esm_kernels.py
"""
import sys

import torch

sys.path.insert(0, ".")
from procyon_b200.model.esm import ESM_PLM  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        m = ESM_PLM(num_params="650m", pooling_method="mean").bfloat16().eval()
    for p in m.parameters():
        if p.dim() > 1:
            p.data.normal_(std=0.02)
    g = torch.Generator().manual_seed(1)
    toks = torch.full((n, 514), 1, dtype=torch.int64)
    toks[:, 0] = 0
    toks[:, 1:513] = torch.randint(4, 24, (n, 512), generator=g)
    toks[:, 513] = 2
    import os
    if os.environ.get("PCY_ESM_ATTN") is not None:
        from procyon_b200 import _lib
        _lib.load().pcy_set_esm_attention_kernel(int(os.environ["PCY_ESM_ATTN"]))
    out, _ = m(toks.to(dev))
    torch.cuda.synchronize()
    print("done", tuple(out.shape))


if __name__ == "__main__":
    main()

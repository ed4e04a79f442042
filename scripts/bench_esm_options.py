"""ESM2-650M encode throughput (256 proteins x 512 residues, pooled + projected) with the optional kernels toggled.
Run on the B200 box."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from procyon_b200 import _lib  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    lib = _lib.load()
    g = torch.Generator().manual_seed(1234)
    N, L = 256, 512
    toks = torch.full((N, L + 2), 1, dtype=torch.int64)
    toks[:, 0] = 0
    toks[:, 1:L + 1] = torch.randint(4, 24, (N, L), generator=g)
    toks[:, L + 1] = 2
    toks = toks.to(dev)

    def run(n=3):
        for _ in range(2):
            model.forward_sequences(toks)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            model.forward_sequences(toks)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    base = run()
    print(json.dumps({"config": "default", "ms": base, "proteins_per_s": N / base * 1e3}), flush=True)
    lib.pcy_set_fused_rope(1)
    t = run()
    print(json.dumps({"config": "fused_rope", "ms": t, "proteins_per_s": N / t * 1e3}), flush=True)
    lib.pcy_set_fused_rope(0)
    lib.pcy_set_gemm_cluster(1)
    t = run()
    print(json.dumps({"config": "gemm_cluster", "ms": t, "proteins_per_s": N / t * 1e3}), flush=True)
    lib.pcy_set_gemm_cluster(0)
    lib.pcy_set_esm_tc_attention(0)
    t = run()
    print(json.dumps({"config": "mma.sync attention", "ms": t, "proteins_per_s": N / t * 1e3}), flush=True)
    lib.pcy_set_esm_tc_attention(1)


if __name__ == "__main__":
    main()

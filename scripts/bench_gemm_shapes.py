"""Isolated timing of the tcgen05 GEMM at the ESM2-650M / Llama-3-8B shapes, the cta_group::2 pair MMA on and off, next to
torch.matmul (cuBLAS) on the same box.  CUDA events, 20 iterations after 5 warm-ups, operands rotated over buffers
larger than L2.  Run on the B200 box."""
import json
import sys

import torch

sys.path.insert(0, ".")
from procyon_b200 import _lib, ops  # noqa: E402

SHAPES = [  # (name, M, N, K)
    ("esm qkv", 64 * 514, 3840, 1280), ("esm out", 64 * 514, 1280, 1280), ("esm fc1", 64 * 514, 5120, 1280),
    ("esm fc2", 64 * 514, 1280, 5120), ("llama qkv", 1024, 6144, 4096), ("llama o", 1024, 4096, 4096),
    ("llama gate_up", 1024, 28672, 4096), ("llama down", 1024, 4096, 14336), ("square 8192", 8192, 8192, 8192),
]


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    lib = _lib.load()
    rows = []
    for name, M, N, K in SHAPES:
        nbuf = max(2, int(300e6 // ((M * K + N * K + M * N) * 2)) + 1)
        A = [(torch.randn(M, K, device="cuda") * 0.5).bfloat16() for _ in range(nbuf)]
        W = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(nbuf)]
        i = [0]

        def ours():
            i[0] = (i[0] + 1) % nbuf
            return ops.linear(A[i[0]], W[i[0]], force="tc")

        def cublas():
            i[0] = (i[0] + 1) % nbuf
            return A[i[0]] @ W[i[0]].t()

        fl = 2.0 * M * N * K
        lib.pcy_set_gemm_cluster(0)
        lib.pcy_set_gemm_pair_mma(0)
        t0 = timeit(ours)
        r0 = ops.linear(A[0], W[0], force="tc")
        lib.pcy_set_gemm_pair_mma(2)
        t2 = timeit(ours)
        same = bool(torch.equal(r0, ops.linear(A[0], W[0], force="tc")))
        lib.pcy_set_gemm_pair_mma(1)
        t1 = timeit(ours)  # the default heuristics (tile width by wave efficiency, pair MMA from three waves)
        tc = timeit(cublas)
        rows.append({"shape": name, "M": M, "N": N, "K": K, "single_cta_tflops": round(fl / t0 / 1e9, 1),
                     "pair_mma_tflops": round(fl / t2 / 1e9, 1), "pair_bit_identical": same,
                     "default_tflops": round(fl / t1 / 1e9, 1), "cublas_tflops": round(fl / tc / 1e9, 1),
                     "default_vs_cublas": round(tc / t1, 3)})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()

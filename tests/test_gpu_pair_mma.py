"""tcgen05 GEMM with cta_group::2 (one MMA of M = 256 issued by the leader of a 2-CTA cluster, each CTA staging its
rows of A and half of the W tile) against the one-CTA kernel and the fp32 CPU reference."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias=None):
    out = a.float() @ w.float().t()
    return out + bias if bias is not None else out


@pytest.mark.parametrize("M,N,K", [(129, 256, 64), (300, 1000, 200), (514, 3840, 1280), (257, 520, 1288),
                                   (4096, 1280, 5120), (1280, 2560, 320), (1024, 6144, 4096)])
def test_pair_mma_gemm_matches_single_cta(cuda_device, M, N, K):
    """Odd numbers of row-blocks (the peer works on padding), N and K tails, bias + GELU + residual and fp32-output
    epilogues. Every output element is the same chain of k-steps in both kernels."""
    from procyon_b200 import _lib, ops

    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).bfloat16().cuda()
    try:
        lib.pcy_set_gemm_pair_mma(0)
        o0 = ops.linear(a, w, bias, residual=res, act=1, force="tc")
        f0 = ops.linear(a, w, bias, force="tc", out_fp32=True)
        lib.pcy_set_gemm_pair_mma(2)
        o1 = ops.linear(a, w, bias, residual=res, act=1, force="tc")
        f1 = ops.linear(a, w, bias, force="tc", out_fp32=True)
        f2 = ops.linear(a, w, bias, force="tc", out_fp32=True)  # back to back: barrier phases / TMEM reuse
    finally:
        lib.pcy_set_gemm_pair_mma(1)
    ref = _ref(a.cpu(), w.cpu(), bias.cpu())
    torch.testing.assert_close(f1.cpu(), ref, rtol=2e-4, atol=2e-4 * ref.abs().max().item())
    assert torch.equal(f1, f2)
    torch.testing.assert_close(f1, f0, rtol=1e-5, atol=1e-5 * ref.abs().max().item())
    torch.testing.assert_close(o1.float(), o0.float(), rtol=1e-2, atol=1e-2)


def test_pair_mma_swiglu_epilogue(cuda_device):
    """gate/up interleaved in 16-row groups + SwiGLU epilogue (the Llama prefill gate_up GEMM) through the pair MMA."""
    from procyon_b200 import _lib, ops

    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    M, F, K = 384, 512, 256
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
    w = (torch.randn(2 * F, K, generator=g) / math.sqrt(K)).bfloat16().cuda()
    try:
        lib.pcy_set_gemm_pair_mma(0)
        o0 = ops.linear(a, w, act=ops.ACT_SWIGLU, force="tc")
        lib.pcy_set_gemm_pair_mma(2)
        o1 = ops.linear(a, w, act=ops.ACT_SWIGLU, force="tc")
    finally:
        lib.pcy_set_gemm_pair_mma(1)
    assert o1.shape == o0.shape == (M, F)
    torch.testing.assert_close(o1.float(), o0.float(), rtol=1e-2, atol=1e-2)

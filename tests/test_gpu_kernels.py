"""Parity of the op-level C ABI (called through ctypes) against the CPU oracle / plain fp32 torch on CPU."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cpu_linear(a, w, bias=None, residual=None, act=0, scale=1.0, scale_ncols=0):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    if scale_ncols:
        v[:, :scale_ncols] *= scale
    if act == 1:
        v = F.gelu(v)
    elif act == 2:
        M, N = v.shape
        v = v.view(M, N // 32, 2, 16)
        v = (F.silu(v[:, :, 0]) * v[:, :, 1]).reshape(M, N // 2)
    if residual is not None:
        v = v + residual.float()
    return v


@pytest.mark.parametrize("M,N,K,force", [
    (128, 256, 64, "tc"), (300, 1000, 200, "tc"), (17, 264, 72, "tc"), (514, 3840, 1280, "tc"),
    (1, 4096, 4096, "skinny"), (3, 1000, 512, "skinny"), (10, 2560, 1280, "skinny"), (16, 520, 264, "skinny"),
    (5, 4096, 14336, "skinny"),
])
@pytest.mark.parametrize("epi", ["plain", "bias_gelu", "bias_res_scale"])
def test_linear_parity(cuda_device, M, N, K, force, epi):
    from procyon_b200 import ops

    g = torch.Generator().manual_seed(M + 13 * N + 7 * K)
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, generator=g) if epi != "plain" else None
    res = torch.randn(M, N, generator=g).bfloat16() if epi == "bias_res_scale" else None
    act = 1 if epi == "bias_gelu" else 0
    sc = 64 if epi == "bias_res_scale" else 0
    ref = _cpu_linear(a, w, bias, res, act, 0.125, sc)
    out = ops.linear(a.cuda(), w.cuda(), bias.cuda() if bias is not None else None,
                     residual=res.cuda() if res is not None else None, act=act, scale=0.125, scale_ncols=sc,
                     force=force)
    # tolerance: one bf16 rounding of the output (2^-8 relative) + fp32 accumulation-order noise
    torch.testing.assert_close(out.float().cpu(), ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item())
    out32 = ops.linear(a.cuda(), w.cuda(), bias.cuda() if bias is not None else None,
                       residual=res.cuda() if res is not None else None, act=act, scale=0.125, scale_ncols=sc,
                       force=force, out_fp32=True)
    torch.testing.assert_close(out32.cpu(), ref, rtol=2e-4, atol=2e-4 * ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(129, 256, 64), (300, 1000, 200), (514, 3840, 1280), (257, 520, 1288),
                                   (4096, 1280, 5120), (1280, 2560, 320)])
def test_cluster_multicast_gemm_is_bit_identical(cuda_device, M, N, K):
    """2-CTA clusters sharing the weight tile by TMA multicast run the same MMAs in the same order as independent
    CTAs: results must be bit-identical (odd numbers of row-blocks pair the last one with padding)."""
    from procyon_b200 import _lib, ops

    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).bfloat16().cuda()
    try:
        lib.pcy_set_gemm_cluster(0)
        o0 = ops.linear(a, w, bias, residual=res, act=1, force="tc")
        f0 = ops.linear(a, w, bias, force="tc", out_fp32=True)
        lib.pcy_set_gemm_cluster(1)
        o1 = ops.linear(a, w, bias, residual=res, act=1, force="tc")
        f1 = ops.linear(a, w, bias, force="tc", out_fp32=True)
    finally:
        lib.pcy_set_gemm_cluster(0)
    assert torch.equal(o0, o1)
    assert torch.equal(f0, f1)
    ref = _cpu_linear(a.cpu(), w.cpu(), bias.cpu())
    torch.testing.assert_close(f1.cpu(), ref, rtol=2e-4, atol=2e-4 * ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(5, 4096, 4096), (10, 6144, 4096), (16, 520, 1280), (7, 1003, 192), (10, 4096, 14336),
                                   (12, 33, 64)])
def test_skinny_tensor_core_kernel(cuda_device, M, N, K):
    """5..16 activation rows: weights streamed through mma.sync vs the scalar-FMA kernel and vs fp32 torch, with
    bias / GELU / residual / fp32-output epilogues, N not a multiple of the 16-row tile, and the SwiGLU layout."""
    from procyon_b200 import _lib, ops

    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 131 + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).bfloat16().cuda()
    try:
        lib.pcy_set_skinny_mma(0)
        o0 = ops.linear(a, w, bias, residual=res, act=1, force="skinny")
        f0 = ops.linear(a, w, bias, force="skinny", out_fp32=True, scale=0.125, scale_ncols=min(N, 32))
        lib.pcy_set_skinny_mma(1)
        o1 = ops.linear(a, w, bias, residual=res, act=1, force="skinny")
        f1 = ops.linear(a, w, bias, force="skinny", out_fp32=True, scale=0.125, scale_ncols=min(N, 32))
    finally:
        lib.pcy_set_skinny_mma(1)
    ref = _cpu_linear(a.cpu(), w.cpu(), bias.cpu(), None, 0, 0.125, min(N, 32))
    torch.testing.assert_close(f1.cpu(), ref, rtol=2e-4, atol=2e-4 * ref.abs().max().item())
    torch.testing.assert_close(f1, f0, rtol=2e-4, atol=2e-4 * ref.abs().max().item())
    torch.testing.assert_close(o1.float(), o0.float(), rtol=1e-2, atol=1e-2 * ref.abs().max().item())
    if N % 32 == 0:
        gate, up = w[: N // 2].contiguous(), w[N // 2:].contiguous()
        packed = ops.pack_gate_up(gate, up)
        s1 = ops.linear(a, packed, act=ops.ACT_SWIGLU, force="skinny")
        sref = F.silu(a.float().cpu() @ gate.float().cpu().t()) * (a.float().cpu() @ up.float().cpu().t())
        torch.testing.assert_close(s1.float().cpu(), sref, rtol=1e-2, atol=1e-2 * sref.abs().max().item())


@pytest.mark.parametrize("M,force", [(1, "skinny"), (4, "skinny"), (9, "skinny"), (200, "tc")])
def test_swiglu_parity(cuda_device, M, force):
    from procyon_b200 import ops

    g = torch.Generator().manual_seed(M)
    Fdim, K = 1024, 512
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16()
    gate = (torch.randn(Fdim, K, generator=g) / math.sqrt(K)).bfloat16()
    up = (torch.randn(Fdim, K, generator=g) / math.sqrt(K)).bfloat16()
    ref = F.silu(a.float() @ gate.float().t()) * (a.float() @ up.float().t())
    packed = ops.pack_gate_up(gate.cuda(), up.cuda())
    out = ops.linear(a.cuda(), packed, act=ops.ACT_SWIGLU, force=force)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item())


def test_linear_is_linear_at_full_size(cuda_device):
    """size-independent property at Llama shapes: f(a1 + a2) == f(a1) + f(a2) (inputs chosen so a1 + a2 is exact)."""
    from procyon_b200 import ops

    torch.manual_seed(0)
    w = (torch.randn(6144, 4096, device="cuda") / 64).bfloat16()
    a1 = (torch.randint(-8, 9, (1024, 4096), device="cuda").float() / 8).bfloat16()
    a2 = (torch.randint(-8, 9, (1024, 4096), device="cuda").float() / 8).bfloat16()
    s = (a1.float() + a2.float()).bfloat16()
    assert torch.equal(s.float(), a1.float() + a2.float())
    o1 = ops.linear(a1, w, out_fp32=True, force="tc")
    o2 = ops.linear(a2, w, out_fp32=True, force="tc")
    os_ = ops.linear(s, w, out_fp32=True, force="tc")
    torch.testing.assert_close(os_, o1 + o2, rtol=1e-3, atol=2e-3)
    # tensor-core and weight-streaming kernels agree on the same rows
    o_sk = ops.linear(a1[:8], w, out_fp32=True, force="skinny")
    torch.testing.assert_close(o_sk, o1[:8], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("d", [320, 480, 1280, 2560, 4096, 5120])
def test_norms(cuda_device, d):
    from procyon_b200 import _lib
    from procyon_b200._lib import c_float, c_i64, c_int, check, ptr, stream_ptr

    lib = _lib.load()
    g = torch.Generator().manual_seed(d)
    x = (torch.randn(37, d, generator=g) * 2 + 0.3).bfloat16()
    w = (1 + 0.1 * torch.randn(d, generator=g)).bfloat16()
    b = (0.1 * torch.randn(d, generator=g)).bfloat16()
    xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
    y = torch.empty_like(xc)
    check(lib.pcy_layernorm_bf16(ptr(xc), ptr(wc), ptr(bc), ptr(y), c_i64(37), c_int(d), c_float(1e-5), stream_ptr()))
    ref = F.layer_norm(x.float(), (d,), w.float(), b.float(), 1e-5)
    torch.testing.assert_close(y.float().cpu(), ref.bfloat16().float(), rtol=1.6e-2, atol=1e-2)
    check(lib.pcy_rmsnorm_bf16(ptr(xc), ptr(wc), ptr(y), c_i64(37), c_int(d), c_float(1e-5), stream_ptr()))
    xf = x.float()
    ref = (w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16().float()).bfloat16()
    torch.testing.assert_close(y.float().cpu(), ref.float(), rtol=1.6e-2, atol=1e-2)


@pytest.mark.parametrize("hd,H,KVH,Tq,Tk,causal,masked", [
    (64, 4, 4, 70, 70, 0, True), (24, 5, 5, 33, 33, 0, True), (16, 4, 4, 130, 130, 0, False),
    (128, 8, 2, 200, 200, 1, True), (128, 4, 1, 64, 64, 1, False), (64, 2, 2, 514, 514, 0, True),
    (128, 8, 2, 5, 133, 1, False),
])
def test_attention_parity(cuda_device, hd, H, KVH, Tq, Tk, causal, masked):
    from procyon_b200 import _lib
    from procyon_b200._lib import c_float, c_i64, c_int, check, ptr, stream_ptr

    lib = _lib.load()
    B = 2
    g = torch.Generator().manual_seed(hd + Tq)
    q = torch.randn(B, Tq, H, hd, generator=g).bfloat16()
    k = torch.randn(B, Tk, KVH, hd, generator=g).bfloat16()
    v = torch.randn(B, Tk, KVH, hd, generator=g).bfloat16()
    valid = torch.ones(B, Tk, dtype=torch.uint8)
    if masked:
        if causal:
            valid[1, :7] = 0  # left padding
        else:
            valid[1, Tk - 9:] = 0  # right padding
    scale = 1.0 / math.sqrt(hd)
    # oracle: eager attention in fp32 (pmc_llama.py:221-247 / fair-esm MultiheadAttention)
    rep = H // KVH
    kk = k.float().repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    vv = v.float().repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    s = (q.float().permute(0, 2, 1, 3) @ kk.transpose(-1, -2)) * scale
    mask = (valid == 0)[:, None, None, :].expand(B, H, Tq, Tk).clone()
    if causal:
        i = torch.arange(Tq)[:, None] + (Tk - Tq)
        j = torch.arange(Tk)[None, :]
        mask |= (j > i)[None, None]
    s = s.masked_fill(mask, float("-inf"))
    pr = torch.softmax(s, dim=-1)
    pr = torch.nan_to_num(pr, nan=0.0)
    ref = (pr @ vv).permute(0, 2, 1, 3)  # B,Tq,H,hd
    qc, kc, vc, vm = q.cuda(), k.cuda(), v.cuda(), valid.cuda()
    o = torch.zeros(B, Tq, H, hd, device="cuda", dtype=torch.bfloat16)
    check(lib.pcy_attention_bf16(ptr(qc), ptr(kc), ptr(vc), ptr(o), c_i64(Tq * H * hd), c_i64(H * hd), c_int(hd),
                                 c_i64(Tk * KVH * hd), c_i64(KVH * hd), c_int(hd), c_i64(Tk * KVH * hd),
                                 c_i64(KVH * hd), c_int(hd), c_i64(Tq * H * hd), c_i64(H * hd), c_int(hd), c_int(B),
                                 c_int(H), c_int(KVH), c_int(Tq), c_int(Tk), c_int(hd), ptr(vm), c_i64(Tk),
                                 c_float(scale), c_int(causal), stream_ptr()))
    got = o.float().cpu()
    rows_ok = ~mask.all(dim=-1).permute(0, 2, 1)  # fully masked query rows are undefined in the reference
    torch.testing.assert_close(got[rows_ok], ref[rows_ok], rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("M,N,K", [(1024, 6144, 4096), (300, 1000, 200), (129, 192, 64), (514, 3840, 1280), (257, 520, 1288),
                                   (1024, 4096, 512)])
@pytest.mark.parametrize("epi", ["plain", "bias_gelu", "bias_res_scale", "swiglu"])
def test_gemm_tile_widths_agree(cuda_device, M, N, K, epi):
    """The one-CTA tcgen05 GEMM with 128x128, 128x192 and 128x256 tiles (`pcy_set_gemm_tile`): the same chain of k-steps
    per output element, so bf16 / fp32 results must be bit-identical across tile widths, and right against fp32 torch.
    (1024, 6144, 4096) is the Llama-3-8B q/k/v projection of a 1024-token prefill, where the heuristic picks 192.)"""
    from procyon_b200 import _lib, ops

    g = torch.Generator().manual_seed(M + 13 * N + 7 * K)
    if epi == "swiglu":
        N = (N // 32) * 32
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().cuda()
    bias = torch.randn(N, generator=g).cuda() if epi in ("bias_gelu", "bias_res_scale") else None
    res = torch.randn(M, N, generator=g).bfloat16().cuda() if epi == "bias_res_scale" else None
    act = {"bias_gelu": 1, "swiglu": 2}.get(epi, 0)
    sc = 64 if epi == "bias_res_scale" else 0
    lib = _lib.load()
    outs = {}
    try:
        lib.pcy_set_gemm_pair_mma(0)
        for width in (256, 192, 128, 0):
            lib.pcy_set_gemm_tile(width)
            outs[width] = (ops.linear(a, w, bias, residual=res, act=act, scale=0.125, scale_ncols=sc, force="tc"),
                           ops.linear(a, w, bias, residual=res, act=act, scale=0.125, scale_ncols=sc, force="tc",
                                      out_fp32=True) if epi != "swiglu" else None)
    finally:
        lib.pcy_set_gemm_tile(0)
        lib.pcy_set_gemm_pair_mma(1)
    for width in (192, 128, 0):
        assert torch.equal(outs[width][0], outs[256][0]), f"tile {width} differs from 256"
        if outs[width][1] is not None:
            assert torch.equal(outs[width][1], outs[256][1])
    ref = _cpu_linear(a.cpu(), w.cpu(), bias.cpu() if bias is not None else None, res.cpu() if res is not None else None,
                      act, 0.125, sc)
    torch.testing.assert_close(outs[192][0].float().cpu(), ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item())

"""GPU bring-up of the GEMM kernels: correctness vs torch fp32 matmul + timing. Run on the B200 box.

    python scripts/bringup_gemm.py [--quick]
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from procyon_b200 import ops  # noqa: E402


def ref_linear(a, w, bias, residual, act, scale, scale_ncols):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    if scale_ncols:
        v[:, :scale_ncols] *= scale
    if act == ops.ACT_GELU:
        v = torch.nn.functional.gelu(v)
    elif act == ops.ACT_SWIGLU:
        M, N = v.shape
        v = v.view(M, N // 32, 2, 16)
        v = (torch.nn.functional.silu(v[:, :, 0]) * v[:, :, 1]).reshape(M, N // 2)
    if residual is not None:
        v = v + residual.float()
    return v


def check(name, M, N, K, *, bias=False, residual=False, act=0, scale_ncols=0, out_fp32=False, force=None, rms=False):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    n_out = N // 2 if act == ops.ACT_SWIGLU else N
    r = torch.randn(M, n_out, device="cuda", generator=g).bfloat16() if residual else None
    rw = None
    a_ref = a
    if rms:
        rw = (1 + 0.1 * torch.randn(K, device="cuda", generator=g)).bfloat16()
        af = a.float()
        rstd = torch.rsqrt(af.pow(2).mean(-1, keepdim=True) + 1e-5)
        a_ref = (rw.float() * (af * rstd).bfloat16().float()).bfloat16()
    out = ops.linear(a, w, b, residual=r, act=act, scale=0.125, scale_ncols=scale_ncols, out_fp32=out_fp32,
                     force=force, rms_weight=rw)
    torch.cuda.synchronize()
    ref = ref_linear(a_ref, w, b, r, act, 0.125, scale_ncols)
    err = (out.float() - ref).abs()
    tol = 2e-2 * ref.abs().max().item() + 1e-3
    ok = bool(err.max().item() <= tol) and bool(torch.isfinite(out.float()).all())
    info = {"name": name, "M": M, "N": N, "K": K, "ok": ok, "max_err": err.max().item(), "tol": tol,
            "ref_absmax": ref.abs().max().item()}
    if not ok:
        bad = (err > tol).nonzero()
        info["n_bad"] = int(bad.shape[0])
        info["first_bad"] = bad[:8].tolist()
        rows = torch.unique(bad[:, 0])
        cols = torch.unique(bad[:, 1])
        info["bad_rows"] = [int(rows.min()), int(rows.max()), int(rows.numel())]
        info["bad_cols"] = [int(cols.min()), int(cols.max()), int(cols.numel())]
        info["sample_out"] = out.float()[:2, :8].tolist()
        info["sample_ref"] = ref[:2, :8].tolist()
    print(json.dumps(info), flush=True)
    return ok


def bench(M, N, K, iters=20, force=None):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(a, w, out=out, force=force)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.linear(a, w, out=out, force=force)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # torch/cuBLAS reference timing
    for _ in range(3):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K + M * N)
    print(json.dumps({"bench": [M, N, K], "force": force, "ms": ms, "tflops": fl / ms / 1e9, "gbs": by / ms / 1e6,
                      "cublas_ms": ms_ref, "cublas_tflops": fl / ms_ref / 1e9}), flush=True)


def main():
    quick = "--quick" in sys.argv
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    # tensor-core path
    ok &= check("tc_one_tile", 128, 256, 64, force="tc")
    ok &= check("tc_one_tile_k256", 128, 256, 256, force="tc")
    ok &= check("tc_bn128", 128, 128, 128, force="tc")
    ok &= check("tc_multi", 512, 1024, 512, force="tc")
    ok &= check("tc_ragged", 300, 1000, 200, force="tc")
    ok &= check("tc_small_m", 17, 264, 72, force="tc")
    ok &= check("tc_bias_gelu", 384, 5120, 1280, bias=True, act=ops.ACT_GELU, force="tc")
    ok &= check("tc_bias_res", 384, 1280, 5120, bias=True, residual=True, force="tc")
    ok &= check("tc_qkv_scale", 514, 3840, 1280, bias=True, scale_ncols=1280, force="tc")
    ok &= check("tc_swiglu", 256, 2048, 512, act=ops.ACT_SWIGLU, force="tc")
    ok &= check("tc_fp32_out", 200, 520, 256, out_fp32=True, force="tc")
    ok &= check("tc_persistent", 4096, 4096, 1024, force="tc")
    # skinny path
    for m in (1, 2, 3, 4, 5, 8, 10, 16):
        ok &= check(f"sk_m{m}", m, 4096, 4096, bias=True, residual=True, force="skinny")
    ok &= check("sk_lmhead", 1, 128263, 4096, out_fp32=True, force="skinny")
    ok &= check("sk_bigk", 10, 4096, 14336, residual=True, force="skinny")
    ok &= check("sk_swiglu", 1, 28672, 4096, act=ops.ACT_SWIGLU, force="skinny")
    ok &= check("sk_swiglu4", 4, 28672, 4096, act=ops.ACT_SWIGLU, force="skinny")
    ok &= check("sk_rms", 4, 6144, 4096, force="skinny", rms=True)
    ok &= check("sk_gelu", 2, 2560, 1280, bias=True, act=ops.ACT_GELU, force="skinny")
    print(json.dumps({"all_ok": bool(ok)}), flush=True)
    if not quick:
        for shape in [(8192, 8192, 8192), (32896, 3840, 1280), (32896, 5120, 1280), (32896, 1280, 5120),
                      (1024, 6144, 4096), (1024, 28672, 4096), (1024, 4096, 14336)]:
            bench(*shape, force="tc")
        for shape in [(1, 6144, 4096), (1, 28672, 4096), (1, 4096, 14336), (1, 128263, 4096), (4, 28672, 4096),
                      (10, 28672, 4096)]:
            bench(*shape, force="skinny")
    return 0 if ok else 1


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(f"elapsed {time.time() - t0:.1f}s", flush=True)
    sys.exit(rc)

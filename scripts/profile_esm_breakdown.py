"""ESM2-650M batch encode (256 proteins x 512 residues, pooled), on the B200 box:
  * throughput for several micro-batch budgets (`max_tokens_per_pass`),
  * in-situ time per kernel class from CUDA events between the kernels of a real encode (`pcy_esm_profile`),
  * SM clock / power sampled while the encode loop runs (is the step power-capped?).
Prints one JSON object per line."""
import ctypes
import json
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from procyon_b200 import _lib  # noqa: E402
from procyon_b200.model.esm import ESM_PLM  # noqa: E402

N, L = 256, 512
NAMES = ["embed", "layernorm", "qkv", "rope", "attention", "out_proj", "fc1", "fc2"]


def sample_clocks(stop, out):
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits",
                                "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            out.append((float(r[0]), float(r[1]), r[2].strip()))
        except Exception:
            pass
        time.sleep(0.05)


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        m = ESM_PLM(num_params="650m", pooling_method="mean").bfloat16().eval()
    for p in m.parameters():
        if p.dim() > 1:
            p.data.normal_(std=0.02)
    lib = _lib.load()
    import os
    if os.environ.get("PCY_PAIR_MMA") == "0":
        lib.pcy_set_gemm_pair_mma(0)
        print(json.dumps({"gemm": "one CTA per tile (cta_group::2 pair MMA off)"}), flush=True)
    g = torch.Generator().manual_seed(1234)
    toks = torch.full((N, L + 2), 1, dtype=torch.int64)
    toks[:, 0] = 0
    toks[:, 1:L + 1] = torch.randint(4, 24, (N, L), generator=g)
    toks[:, L + 1] = 2
    toks = toks.to(dev)
    T, d, layers = L + 2, 1280, 33
    flops = N * T * layers * (24 * d * d + 4 * T * d)

    def run(n=4, warm=2):
        for _ in range(warm):
            m(toks)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            m(toks)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    for budget in (256 * 1024,):
        m.max_tokens_per_pass = budget
        ms = run()
        print(json.dumps({"max_tokens_per_pass": budget, "ms": round(ms, 2), "proteins_per_s": round(N / ms * 1e3, 1),
                          "tflops": round(flops / ms / 1e9, 1)}), flush=True)
    m.max_tokens_per_pass = 256 * 1024

    # clocks / power under a 3 s encode loop
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, samples))
    th.start()
    t0 = time.time()
    while time.time() - t0 < 3.0:
        m(toks)
    torch.cuda.synchronize()
    stop.set()
    th.join()
    if samples:
        mhz = sorted(s[0] for s in samples)
        print(json.dumps({"clock_samples": len(samples), "sm_mhz_median": mhz[len(mhz) // 2], "sm_mhz_min": mhz[0],
                          "power_w_max": max(s[1] for s in samples), "reasons": sorted({s[2] for s in samples})}), flush=True)

    # per-class breakdown (events between kernels; adds a sync per encode, so the total is a little above `ms`)
    import os
    full = os.environ.get("PCY_ESM_BREAKDOWN_ALL", "0") == "1"
    # (attention kernel, Q roped inside it, rows beyond the last full query tile sent to the mma.sync kernel)
    variants = [(5, False, 0), (6, False, 0)]
    if full:
        variants += [(4, True, 0), (2, True, 0), (1, False, 0), (0, False, 0)]
    for steps64, qr, tail in variants:
        lib.pcy_set_esm_attention_q_rope(1 if qr else 0)
        lib.pcy_set_esm_attention_tail_rows(tail)
        lib.pcy_set_esm_attention_kernel(steps64)
        ms = run(n=3, warm=1)
        lib.pcy_esm_profile(1)
        reps = 3
        for _ in range(reps):
            m(toks)
        buf = (ctypes.c_double * 8)()
        lib.pcy_esm_profile_read(buf, 8)
        lib.pcy_esm_profile(0)
        per = {k: round(buf[i] / reps, 3) for i, k in enumerate(NAMES)}
        tot = sum(per.values())
        gemm_fl = {"qkv": 6, "out_proj": 2, "fc1": 8, "fc2": 8}
        tf = {k: round(N * T * layers * v * d * d / per[k] / 1e9, 1) for k, v in gemm_fl.items() if per[k] > 0}
        tf["attention"] = round(N * T * layers * 4 * T * d / per["attention"] / 1e9, 1) if per["attention"] > 0 else None
        print(json.dumps({"tail_rows_to_mma_sync": tail, "q_rope_in_attention": bool(steps64 >= 2 and qr), "attention_kernel": ["128-key steps", "64-key steps, double-buffered", "64-key steps, Q and P in TMEM", "64-key steps, Q and P in TMEM, ALU pack", "64-key steps, Q and P in TMEM, pair barriers", "64-key steps, Q and P in TMEM, pair barriers, O accumulated in TMEM", "64-key steps, one thread per query row, Q / P / O in TMEM", "64-key steps, one thread per query row, persistent CTAs"][steps64],
                          "ms_untraced": round(ms, 2), "proteins_per_s": round(N / ms * 1e3, 1), "ms_per_class": per,
                          "sum_ms": round(tot, 2), "share": {k: round(v / tot, 3) for k, v in per.items()},
                          "tflops_per_class": tf}), flush=True)
    lib.pcy_set_esm_attention_kernel(5)
    lib.pcy_set_esm_attention_q_rope(0)
    lib.pcy_set_esm_attention_tail_rows(0)


if __name__ == "__main__":
    main()

"""BASELINE config 4 on one GPU: 1 text(+protein) query against a 20 000-protein embedding database.

  * scoring kernel alone (`pcy_cosine_scores`: normalise + dot in one pass over the fp32 database): CUDA events, the
    database rotated over 5 copies (512 MB > the 126 MB L2) so that every launch streams from HBM; GB/s of the
    algorithmic N*d*4 bytes against the measured HBM peak;
  * end-to-end query latency through the public API: `model(inputs, retrieval=True)` (tokenise, ESM2 encode of the
    query protein, splice, Llama prefill of a 1024-token prompt ending in [PROT], aaseq_lm_projector) +
    `get_proteins_from_embedding(..., top_k=20)` with the result on the host.
Run on the B200 box; prints one JSON object per line.  (Not part of bench.py: it has not run on a GPU yet.)"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from procyon_b200.data.inference_utils import cosine_scores, get_proteins_from_embedding  # noqa: E402

N_DB = 20000


def main():
    dev = torch.device("cuda", 0)
    hbm_peak = bench._peaks()[0]
    from procyon_b200.data.inference_utils import retrieval_scores_topk

    for d in (1280, 2560):
        g = torch.Generator().manual_seed(99)
        n_copies = 6
        dbs = [torch.randn(N_DB, d, generator=g).to(dev) for _ in range(n_copies)]
        q = torch.randn(1, d, generator=g).to(dev)
        out = torch.empty((1, N_DB), device=dev, dtype=torch.float32)
        ref = torch.nn.functional.normalize(q.float()) @ torch.nn.functional.normalize(dbs[n_copies - 1]).T
        for k in (0, 20):
            # one CUDA graph of n_copies launches (one per database copy: 6 x 102 MB > L2): kernel time without the
            # host-side launch cost of the Python wrapper
            def launch_all():
                for i in range(n_copies):
                    if k:
                        retrieval_scores_topk(q, dbs[i], k, scores_out=out)
                    else:
                        cosine_scores(q, dbs[i], out=out)
            launch_all()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                with torch.cuda.graph(gr, stream=side):
                    launch_all()
            torch.cuda.current_stream(dev).wait_stream(side)
            for _ in range(3):
                gr.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            a.record()
            for _ in range(iters):
                gr.replay()
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / (iters * n_copies) * 1e3
            gbs = N_DB * d * 4 / us / 1e3
            print(json.dumps({"what": "retrieval scoring kernel" + (f" + fused top-{k}" if k else " (scores only)"),
                              "n_db": N_DB, "d": d, "us_per_query": round(us, 2), "achieved_gbs": round(gbs, 1),
                              "hbm_peak_gbs": hbm_peak, "frac": round(gbs / hbm_peak, 3), "timing": "CUDA-graph replay",
                              "max_abs_err_vs_torch": float((out - ref).abs().max())}), flush=True)
        del dbs

    model = bench.build_model(dev)
    inputs = bench.synth_inputs(model)
    inputs["instructions"] = [inputs["instructions"][0].rsplit(" ", 2)[0] + " [ANSWER] [PROT]"]
    d = model.protein_embed_dim
    db = torch.randn(N_DB, d, generator=torch.Generator().manual_seed(99)).to(dev)

    def query():
        out = model(inputs, retrieval=True, aaseq_type="protein")
        return get_proteins_from_embedding(db, out, top_k=20)

    for _ in range(3):
        df = query()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        df = query()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    n_tok = len(model.tokenizer(inputs["instructions"][0], add_special_tokens=True)["input_ids"])
    print(json.dumps({"what": "end-to-end retrieval query (host inputs -> top-20 DataFrame)", "n_db": N_DB, "d": d,
                      "prompt_tokens": n_tok, "ms_per_query": round(ms, 2), "top1_index": int(df["index"].iloc[0])}),
          flush=True)


if __name__ == "__main__":
    main()

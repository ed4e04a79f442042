"""Short driver for ncu: one ESM2-650M encode micro-batch, one Llama-3-8B prefill, a few eager decode steps.

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pcy --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_paths.py [--what esm,prefill,decode]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="esm,prefill,decode")
    ap.add_argument("--proteins", type=int, default=64)
    ap.add_argument("--decode-steps", type=int, default=2)
    ap.add_argument("--beams", type=int, default=1)
    args = ap.parse_args()
    what = set(args.what.split(","))
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    inputs = bench.synth_inputs(model)
    torch.cuda.synchronize()
    if "esm" in what:
        g = torch.Generator().manual_seed(1)
        toks = torch.full((args.proteins, 514), 1, dtype=torch.int64)
        toks[:, 0] = 0
        toks[:, 1:513] = torch.randint(4, 24, (args.proteins, 512), generator=g)
        toks[:, 513] = 2
        torch.cuda.nvtx.range_push("esm_encode")
        out = model.forward_sequences(toks.to(dev))
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    if "prefill" in what or "decode" in what:
        from procyon_b200.model.pmc_llama import SELECT_BEAM, SELECT_GREEDY

        (x, ids, am, _, _, _) = model._preprocessing(inputs, crop_off=True, no_pad=True, left_pad=True)
        te = model.text_encoder
        sess = te.get_session(1, args.beams, x.shape[1], bench.GEN_LEN, dev, False, False)
        sel = torch.tensor([x.shape[1] - 1], device=dev, dtype=torch.int32)
        torch.cuda.nvtx.range_push("prefill")
        _, _, logits, _ = te.prefill(x, None, want_cache=True, want_hidden=False, sel_rows=sel, kv_out=sess.kv_prompt)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        if "decode" in what:
            mode = SELECT_GREEDY if args.beams == 1 else SELECT_BEAM
            sess.reset(logits)
            sess.select(mode, args.beams, 0.8, -1, False)
            torch.cuda.nvtx.range_push("decode")
            for _ in range(args.decode_steps):
                sess.forward()
                sess.select(mode, args.beams, 0.8, -1, False)
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
    print("done")


if __name__ == "__main__":
    main()

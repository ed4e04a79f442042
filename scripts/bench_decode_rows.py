"""ms per decode step (forward + selection, CUDA graph replay) for several row counts: 1 (greedy), 4 (small beam:
persistent kernel) and 10 (the reference's evaluation default, beam_size=10: one launch per op).  Run on the B200 box."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from procyon_b200.model.pmc_llama import SELECT_BEAM, SELECT_GREEDY  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    inputs = bench.synth_inputs(model)
    (x, ids, am, _, _, _) = model._preprocessing(inputs, crop_off=True, no_pad=True, left_pad=True)
    te = model.text_encoder
    for beams in [int(b) for b in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "4", "10"])]:
        sess = te.get_session(1, beams, x.shape[1], bench.GEN_LEN, dev, False, False)
        sel = torch.tensor([x.shape[1] - 1], device=dev, dtype=torch.int32)
        _, _, logits, _ = te.prefill(x, None, want_cache=True, want_hidden=False, sel_rows=sel, kv_out=sess.kv_prompt)
        mode = SELECT_GREEDY if beams == 1 else SELECT_BEAM
        group = 1 if beams == 1 else (beams // 2 if beams % 2 == 0 else beams)
        from procyon_b200 import _lib

        lib = _lib.load()
        if os.environ.get("PCY_BENCH_MAX_ROWS"):  # rows up to which the GREEDY persistent kernel is used (default 2)
            lib.pcy_set_decode_megakernel(int(os.environ["PCY_BENCH_MAX_ROWS"]))
        # rows <= 2: the greedy persistent kernel; more rows: the persistent beam kernel, then the per-op chain
        for rows_kernel, pdl in (((1, 1), (0, 1)) if beams > 2 else ((1, 1),)):
            lib.pcy_set_pdl(pdl)
            lib.pcy_set_decode_rows_megakernel(rows_kernel)
            sess.reset(logits)
            sess.select(mode, group, 0.8, -1, False)
            sess._graph = None
            g = sess.step_graph(mode, group, 0.8, -1, False)
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 60
            a.record()
            for _ in range(n):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n
            print(json.dumps({"beams": beams, "persistent_kernel": bool(rows_kernel) or beams <= 2,
                              "programmatic_dependent_launch": bool(pdl), "ms_per_step": ms,
                              "tokens_per_s_aggregate": beams / ms * 1e3, "weights_gb_per_s": 15.01 / ms * 1e3}),
                  flush=True)
        lib.pcy_set_pdl(1)
        lib.pcy_set_decode_rows_megakernel(1)


if __name__ == "__main__":
    main()

"""Does capturing several decode steps in ONE CUDA graph beat replaying a one-step graph?  (launch gaps between graph
launches vs between kernel nodes).  Run on the B200 box: python scripts/bench_multistep_graph.py [beams]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from procyon_b200.model.pmc_llama import SELECT_BEAM, SELECT_GREEDY  # noqa: E402


def main():
    beams = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    inputs = bench.synth_inputs(model)
    (x, ids, am, _, _, _) = model._preprocessing(inputs, crop_off=True, no_pad=True, left_pad=True)
    te = model.text_encoder
    sess = te.get_session(1, beams, x.shape[1], bench.GEN_LEN, dev, False, False)
    sel = torch.tensor([x.shape[1] - 1], device=dev, dtype=torch.int32)
    _, _, logits, _ = te.prefill(x, None, want_cache=True, want_hidden=False, sel_rows=sel, kv_out=sess.kv_prompt)
    mode = SELECT_GREEDY if beams == 1 else SELECT_BEAM
    group = 1 if beams == 1 else 2
    for n_steps in (1, 4, 16):
        sess.reset(logits)
        sess.select(mode, group, 0.8, -1, False)
        sess.forward()
        sess.select(mode, group, 0.8, -1, False)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(n_steps):
                    sess.forward()
                    sess.select(mode, group, 0.8, -1, False)
        torch.cuda.current_stream(dev).wait_stream(side)
        reps = 96 // n_steps
        g.replay()
        torch.cuda.synchronize()
        sess.reset(logits)
        sess.select(mode, group, 0.8, -1, False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        print(json.dumps({"beams": beams, "steps_per_graph": n_steps, "ms_per_step": a.elapsed_time(b) / (reps * n_steps)}), flush=True)


if __name__ == "__main__":
    main()

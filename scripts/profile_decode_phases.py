"""Phase timing of the persistent decode kernel (globaltimer stamps at every grid barrier). Run on the B200 box."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from procyon_b200 import _lib  # noqa: E402
from procyon_b200.model.pmc_llama import SELECT_GREEDY  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    inputs = bench.synth_inputs(model)
    (x, ids, am, _, _, _) = model._preprocessing(inputs, crop_off=True, no_pad=True, left_pad=True)
    te = model.text_encoder
    sess = te.get_session(1, 1, x.shape[1], bench.GEN_LEN, dev, False, False)
    sel = torch.tensor([x.shape[1] - 1], device=dev, dtype=torch.int32)
    _, _, logits, _ = te.prefill(x, None, want_cache=True, want_hidden=False, sel_rows=sel, kv_out=sess.kv_prompt)
    sess.reset(logits)
    sess.select(SELECT_GREEDY, 1, 0.0, -1, False)
    L = te.model.config.num_hidden_layers
    # stamps per layer: P1 [stage, stream, epi+kv request, bar], P2 [q/rope, scores, softmax, pv, partials+merge, tail,
    # bar], P3 [4], P4 [4], P5 [4] = 23; + lm head [stage, stream, end]
    per_layer = 24
    n_stamps = 1 + per_layer * L + 3
    buf = torch.zeros(max(n_stamps + 8, 4200), device=dev, dtype=torch.int64)  # (word 4095 = optional "SKEW" tag)
    lib = _lib.load()
    for _ in range(3):
        sess.forward()
        sess.select(SELECT_GREEDY, 1, 0.0, -1, False)
    lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(buf.data_ptr()))
    acc = torch.zeros(n_stamps - 1, dtype=torch.float64)
    n = 10
    for _ in range(n):
        sess.forward()
        sess.select(SELECT_GREEDY, 1, 0.0, -1, False)
        torch.cuda.synchronize()
        t = buf.cpu().double()[:n_stamps]
        acc += (t[1:] - t[:-1])
    lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(0))
    acc /= n * 1e3  # us
    per = acc[: per_layer * L].view(L, per_layer)[1:].mean(0)  # skip layer 0 (different P1)
    names = ["P1 stage", "P1 stream", "P1 epilogue", "P2 kv request", "P1 barrier", "P2 q load + rope", "P2 scores", "P2 softmax",
             "P2 p.v", "P2 partials + ticket merge", "P2 tail", "P2 barrier", "P3 stage", "P3 stream", "P3 epilogue",
             "P3 barrier", "P4 stage", "P4 stream", "P4 epilogue", "P4 barrier", "P5 stage", "P5 stream", "P5 epilogue",
             "P5 barrier"]
    for nm, v in zip(names, per.tolist()):
        print(f"{nm:22s} {v:8.2f} us")
    print(f"layer total            {per.sum():8.2f} us   (ideal streaming 66.7 us)")
    print("lm head [stage, stream, epilogue]", [round(x, 2) for x in acc[per_layer * L:].tolist()])
    print(f"step total             {acc.sum():8.2f} us")


if __name__ == "__main__":
    main()

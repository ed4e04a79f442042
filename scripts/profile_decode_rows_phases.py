"""Phase timing of the persistent beam-search decode kernel (csrc/decode_rows_megakernel.cu): %globaltimer stamps of CTA
0 at every phase boundary, averaged over layers 1.. and over 10 steps.  Run on the B200 box.
usage: python scripts/profile_decode_rows_phases.py [beams=10] [steps_before=1]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from procyon_b200 import _lib  # noqa: E402
from procyon_b200.model.pmc_llama import SELECT_BEAM  # noqa: E402

NAMES = ["qkv stage", "qkv stream", "qkv pieces/epilogue",
         "att: K/V request", "att: barrier (qkv done)", "att: item loads + Q/K RoPE",
         "att: wait for the K/V tiles", "att: S, mask, softmax", "att: P.V + partials", "att: barrier (partials)",
         "att: merge", "att: barrier (attn done)",
         "o stage", "o stream", "o pieces/epilogue", "o barrier",
         "gate/up stage", "gate/up stream", "gate/up pieces/epilogue", "gate/up barrier",
         "down stage", "down stream", "down pieces/epilogue", "down barrier"]


def main():
    beams = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    inputs = bench.synth_inputs(model)
    (x, ids, am, _, _, _) = model._preprocessing(inputs, crop_off=True, no_pad=True, left_pad=True)
    te = model.text_encoder
    sess = te.get_session(1, beams, x.shape[1], bench.GEN_LEN, dev, False, False)
    sel = torch.tensor([x.shape[1] - 1], device=dev, dtype=torch.int32)
    _, _, logits, _ = te.prefill(x, None, want_cache=True, want_hidden=False, sel_rows=sel, kv_out=sess.kv_prompt)
    sess.reset(logits)
    group = beams // 2 if beams % 2 == 0 else beams
    sess.select(SELECT_BEAM, group, 0.8, -1, False)
    L = te.model.config.num_hidden_layers
    per_layer = len(NAMES)
    n_stamps = 1 + per_layer * L + 3
    buf = torch.zeros(n_stamps + 64, device=dev, dtype=torch.int64)
    lib = _lib.load()
    for _ in range(warm):
        sess.forward()
        sess.select(SELECT_BEAM, group, 0.8, -1, False)
    lib.pcy_set_decode_rows_timing_buffer(ctypes.c_void_p(buf.data_ptr()))
    acc = torch.zeros(n_stamps - 1, dtype=torch.float64)
    n = 10
    for _ in range(n):
        sess.forward()
        sess.select(SELECT_BEAM, group, 0.8, -1, False)
        torch.cuda.synchronize()
        t = buf.cpu().double()[:n_stamps]
        acc += (t[1:] - t[:-1])
    lib.pcy_set_decode_rows_timing_buffer(ctypes.c_void_p(0))
    acc /= n * 1e3  # us
    per = acc[: per_layer * L].view(L, per_layer)[1:].mean(0)  # skip layer 0 (embedding rows staged from the table)
    for nm, v in zip(NAMES, per.tolist()):
        print(f"{nm:34s} {v:8.2f} us")
    print(f"layer total                        {per.sum():8.2f} us   (ideal streaming 66.7 us)")
    print("lm head [stage, stream, pieces/epilogue]", [round(v, 2) for v in acc[per_layer * L:].tolist()])
    print(f"step total                         {acc.sum():8.2f} us  (beams {beams}, steps {warm + 1}..{warm + n})")


if __name__ == "__main__":
    main()

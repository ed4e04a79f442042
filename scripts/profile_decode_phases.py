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
    buf = torch.zeros(5 * L + 2, device=dev, dtype=torch.int64)
    lib = _lib.load()
    for _ in range(3):
        sess.forward()
        sess.select(SELECT_GREEDY, 1, 0.0, -1, False)
    lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(buf.data_ptr()))
    acc = torch.zeros(5 * L + 1, dtype=torch.float64)
    n = 10
    for _ in range(n):
        sess.forward()
        sess.select(SELECT_GREEDY, 1, 0.0, -1, False)
        torch.cuda.synchronize()
        t = buf.cpu().double()
        acc += (t[1:] - t[:-1])
    lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(0))
    acc /= n * 1e3  # us
    names = ["P1 qkv", "P2 attention", "P3 o_proj", "P4 gate/up", "P5 down"]
    per = acc[:-1].view(L, 5)
    ideal = [50.3e6, 0, 33.6e6, 234.9e6, 117.4e6]
    print("phase                mean us   min us   max us   ideal us @6.54TB/s")
    for i, nm in enumerate(names):
        print(f"{nm:18s} {per[:, i].mean():9.2f} {per[:, i].min():8.2f} {per[:, i].max():8.2f} {ideal[i] / 6538.9e3:9.2f}")
    print(f"lm_head            {acc[-1]:9.2f}                    {128263 * 4096 * 2 / 6538.9e3:9.2f}")
    print(f"total              {acc.sum():9.2f} us")


if __name__ == "__main__":
    main()

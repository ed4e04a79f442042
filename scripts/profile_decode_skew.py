"""Per-CTA finish times of the weight phases of the persistent decode kernel: is the skew at the grid barriers
systematic per SM?  Run on the B200 box."""
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
    n_ph = 4 * L + 1
    G = 148
    buf = torch.zeros(4096 + n_ph * G * 2 + 64, device=dev, dtype=torch.int64)
    buf[4095] = 0x534B4557  # "SKEW": per-CTA phase-end stamps after the first 4096 words
    lib = _lib.load()
    for _ in range(3):
        sess.forward()
        sess.select(SELECT_GREEDY, 1, 0.0, -1, False)
    lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(buf.data_ptr()))
    runs = []
    for _ in range(4):
        sess.forward()
        sess.select(SELECT_GREEDY, 1, 0.0, -1, False)
        torch.cuda.synchronize()
        t = buf[4096:4096 + n_ph * G * 2].cpu().view(n_ph, G, 2)
        runs.append(t.clone())
    lib.pcy_set_decode_timing_buffer(ctypes.c_void_p(0))
    names = {0: "P1 qkv", 1: "P3 o", 2: "P4 gate/up", 3: "P5 down"}
    for ph in (4 * 10 + 2, 4 * 10 + 3, 4 * L):
        print("phase", ph, names.get(ph % 4, "") if ph < 4 * L else "lm head")
        lat = []
        for r in runs:
            tt = r[ph, :, 0].double()
            smid = r[ph, :, 1]
            rel = (tt - tt.min()) / 1e3
            order = torch.argsort(smid)
            lat.append(rel[order])
            print(f"  spread: max-min {rel.max():.2f} us, mean lag {rel.mean():.2f} us, last SM {int(smid[rel.argmax()])}, "
                  f"first SM {int(smid[rel.argmin()])}")
        lat = torch.stack(lat)
        c = torch.corrcoef(lat)
        print("  run-to-run correlation of per-SM lag:", [round(float(c[0, i]), 2) for i in range(1, lat.shape[0])])
        m = lat.mean(0)
        top = torch.argsort(m, descending=True)[:12]
        print("  slowest SMs (sorted smid index: mean lag us):", [(int(i), round(float(m[i]), 2)) for i in top])
        # lag by SM parity / range
        print("  mean lag SM 0-73: %.2f us, 74-147: %.2f us" % (float(m[:74].mean()), float(m[74:].mean())))


if __name__ == "__main__":
    main()

"""Run-to-run differences of ESM attention kernel 5 (O accumulated in TMEM) on a one-layer encoder."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import esm2 as O  # noqa: E402  (debug script, not product)
from procyon_b200 import _lib  # noqa: E402
from procyon_b200.model.esm import ESM_PLM  # noqa: E402

L, d, H = 1, 256, 4
for lengths in ([298, 131, 260], [512, 300, 130, 64], [126, 126], [512, 300, 130, 64, 257, 511, 129, 40] * 6):
    sd = O.random_esm_state_dict(L, d, seed=21)
    toks = O.random_protein_tokens(len(lengths), 0, seed=9, lengths=lengths)
    m = ESM_PLM(num_params="custom", pooling_method="mean", protein_pooling_correction_option=False, custom_config=(L, d, H), max_protein_len=1024)
    m.model.load_state_dict(sd, strict=True)
    m = m.cuda()
    lib = _lib.load()
    lib.pcy_set_esm_attention_kernel(4)
    ref = m.encode_tokens(toks.cuda()).float().cpu()
    lib.pcy_set_esm_attention_kernel(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
    outs = [m.encode_tokens(toks.cuda()).float().cpu() for _ in range(12)]
    lib.pcy_set_esm_attention_kernel(4)
    nonpad = toks != O.PAD_IDX
    print("lengths", lengths, "T", toks.shape[1], "max |k5 - k4| over valid rows", (outs[0] - ref)[nonpad].abs().max().item())
    bad = torch.zeros_like(nonpad)
    for o in outs[1:]:
        diff = (o - outs[0]).abs().amax(-1) > 0
        bad |= diff
    for b in range(toks.shape[0]):
        rows = torch.nonzero(bad[b] & nonpad[b]).flatten().tolist()
        rows_pad = torch.nonzero(bad[b] & ~nonpad[b]).flatten().tolist()
        print("  protein", b, "len", lengths[b] + 2, "rows differing run to run (valid):", rows[:40], "n=", len(rows), "| in padding:", len(rows_pad))
        if rows:
            t = rows[0]
            vals = torch.stack([o[b, t] for o in outs])
            cols = torch.nonzero((vals - vals[0]).abs().amax(0) > 0).flatten().tolist()
            print("    row", t, "columns differing:", cols[:24], "n=", len(cols), "max diff", (vals - vals[0]).abs().max().item(), "vs k4", (vals[0] - ref[b, t]).abs().max().item())

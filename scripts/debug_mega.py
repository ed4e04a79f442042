"""Debug helper: run the persistent and the per-op decode paths separately on the wide-ffn test model."""
import os
import sys

os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_llama import _build, _cfgs, _inputs  # noqa: E402

from oracle.llama import random_llama_state_dict  # noqa: E402
from procyon_b200 import _lib  # noqa: E402
from procyon_b200.model.generation import generate_greedy  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4
which = sys.argv[2] if len(sys.argv) > 2 else "both"
oc, pc = _cfgs("gq4wide")
sd = random_llama_state_dict(oc, seed=3)
m = _build(sd, pc)
ids, emb, mask = _inputs(oc, sd, rows, 50, seed=7, pad_left=4)
lib = _lib.load()
if which in ("both", "mega"):
    lib.pcy_set_decode_megakernel(4)
    o1, lp1, lg1 = generate_greedy(m, emb.cuda(), mask.cuda(), max_len=10)
    torch.cuda.synchronize()
    print("megakernel ok", o1[0].tolist())
if which in ("both", "perop"):
    lib.pcy_set_decode_megakernel(0)
    o2, lp2, lg2 = generate_greedy(m, emb.cuda(), mask.cuda(), max_len=10, use_graph=False)
    torch.cuda.synchronize()
    print("per-op ok", o2[0].tolist())

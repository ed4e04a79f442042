"""Where the end-to-end retrieval query (BASELINE config 4) spends its host time: cProfile of 10 queries."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from procyon_b200.data.inference_utils import ShardedProteinIndex  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    instr = bench._synth_instruction(model, bench.PROMPT_LEN, "Which protein is described by :", "[ANSWER] [PROT]", seed=777)
    inputs = {"data": {"seq": None, "seq_idx": None, "text": [], "text_idx": [], "drug": None},
              "input": {"seq": None, "text": [[]], "drug": None},
              "target": {"seq": None, "text": None, "drug": None},
              "instructions": [instr], "reference_indices": {"input": {"seq": [[]]}, "target": {"text": [0]}}}
    db = torch.randn(bench.N_DB, model.protein_embed_dim, generator=torch.Generator().manual_seed(99))
    index = ShardedProteinIndex(db, dev)

    def query():
        out = model(inputs, retrieval=True, aaseq_type="protein")
        val, idx = index.topk(out["contrastive_out"]["positive"]["text"][:1].float(), 20)
        return val.cpu(), idx.cpu()

    for _ in range(3):
        query()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        query()
    torch.cuda.synchronize()
    print("S padded to", model.config.max_text_len); print("ms per query", (time.perf_counter() - t0) / 10 * 1e3)
    # GPU time alone
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        query()
    b.record()
    torch.cuda.synchronize()
    print("event ms per query", a.elapsed_time(b) / 10)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        query()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()

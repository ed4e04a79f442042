"""The one exchange step of the hot path outside the encode / scoring shards: the in-batch
contrastive loss of BASELINE config 5 with `contrastive_global` — every rank all-gathers the normalised embeddings and
the id vectors, builds the global negatives mask and scores its own b rows against all W*b.

  * CPU, gloo, world 2 and 3: the id all-gathers + conflict matrix of `UnifiedProCyon._conflict_matrix` against the
    oracle's restatement on the concatenated ids (host logic; no kernel involved);
  * GPU, NCCL, world 2 (needs 2 GPUs: `gpurun --gpus 2`): `UnifiedProCyon.forward(retrieval=True)` in train mode on
    every rank — loss of each rank against the oracle composed from the gathered embeddings and ids.
"""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _retrieval_inputs(rank, b=4, with_text=True):
    """A retrieval batch of b pairs per rank with id collisions inside and ACROSS ranks (false negatives)."""
    g = torch.Generator().manual_seed(50 + rank)
    seq_idx = torch.tensor([7, 8, 9, 7, 11, 12][:b + 2]) + (0 if rank == 0 else torch.tensor([0, 0, 3, 5, 0, 1][:b + 2]))
    text_idx = [100, 101, 100, 103][:b] if rank == 0 else [100, 104, 101, 103][:b]
    inputs = {
        "data": {"seq": None, "seq_idx": seq_idx, "text": [f"t{i}" for i in range(b)], "text_idx": text_idx,
                 "drug": None},
        "input": {"seq": [[] for _ in range(b)] if with_text else [[i] for i in range(b)],
                  "text": [[i] for i in range(b)] if with_text else [[] for _ in range(b)], "drug": None},
        "target": {"seq": {"positive": list(range(b)), "negative": None}, "text": None, "drug": None},
        "dataset_id": torch.tensor([0, 0, 4, 2][:b]) if rank == 0 else torch.tensor([0, 4, 0, 2][:b]),
    }
    return inputs


def _conflict_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.fusion import conflict_matrix
        from procyon_b200.model.model_unified import UnifiedProCyon

        for with_text, with_dset, aaseq in ((True, True, "protein"), (True, False, "domain"), (False, True, "protein")):
            stub = types.SimpleNamespace(config=types.SimpleNamespace(contrastive_global=True))
            mine = _retrieval_inputs(rank % 2, with_text=with_text)
            if not with_dset:
                mine.pop("dataset_id")
            got = UnifiedProCyon._conflict_matrix(stub, mine, aaseq, torch.device("cpu"))
            # oracle on the concatenation of every rank's ids, in rank order
            text_ids, prot_ids, dset = [], [], []
            for r in range(world):
                inp = _retrieval_inputs(r % 2, with_text=with_text)
                if with_text:
                    text_ids += [inp["data"]["text_idx"][row[-1]] for row in inp["input"]["text"]]
                else:
                    text_ids += [-1 - int(inp["data"]["seq_idx"][row[-1]]) for row in inp["input"]["seq"]]
                prot_ids += [int(inp["data"]["seq_idx"][i]) for i in inp["target"]["seq"]["positive"]]
                dset += inp["dataset_id"].tolist()
            ref = conflict_matrix(torch.tensor(text_ids), torch.tensor(prot_ids),
                                  {"protein": 0, "domain": 1, "peptide": 2}[aaseq],
                                  torch.tensor(dset) if with_dset else None)
            G = world * 4
            assert got.shape == ref.shape == (G, G)
            assert torch.equal(got, ref), (with_text, with_dset, aaseq)
            assert bool(ref.diagonal().all())  # a sample never conflicts with itself
            if with_text:
                assert not bool(ref.all())  # the id vectors hold real false negatives, inside and across ranks
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_conflict_matrix_id_gathers_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_conflict_worker, args=(world, 29640 + world, ret), nprocs=world, join=True)
    assert all(ret.get(r) == 1 for r in range(world))


# --------------------------------------------------------------------------------------------------- NCCL, 2 GPUs
def _nccl_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import torch.nn.functional as F

        import test_gpu_unified as TU
        from oracle.esm2 import random_protein_tokens
        from oracle.fusion import conflict_matrix, infonce

        m = TU._tiny_model()  # same seed on every rank -> identical replicas, like DDP
        m = m.to(dev)
        m.config.contrastive_global = True
        m.config.filter_negatives_by_id_contrastive = True
        m.contrastive_head.all_gather_version = True
        b = 3
        toks = random_protein_tokens(4, 0, seed=8 + rank, lengths=[30, 12, 21, 17])
        seq_idx = torch.tensor([11, 12, 13, 14]) if rank == 0 else torch.tensor([11, 22, 13, 24])
        text_idx = [4, 9, 4] if rank == 0 else [4, 5, 9]
        inputs = {
            "data": {"seq": toks, "seq_idx": seq_idx, "text": ["binds atp", "membrane transport", "binds atp"]
                     if rank == 0 else ["binds atp", "kinase", "membrane transport"], "text_idx": text_idx, "drug": None},
            "input": {"seq": [[], [], []], "text": [[0], [1], [2]], "drug": None},
            "target": {"seq": {"positive": [0, 1, 2], "negative": None}, "text": None, "drug": None},
            "instructions": ["Context : [EXT] Which protein does this ? [ANSWER] [PROT]"] * b,
            "reference_indices": {"input": {"seq": [[], [], []]}, "target": {"text": [0, 1, 2]}},
            "dataset_id": torch.tensor([0, 0, 0]) if rank == 0 else torch.tensor([0, 2, 0]),
        }
        m.train()
        out = m(inputs, retrieval=True, aaseq_type="protein")
        loss = out["contrastive_loss"].float().reshape(1)
        zs = out["contrastive_out"]["positive"]["sequence"].float().contiguous()
        zt = out["contrastive_out"]["positive"]["text"].float().contiguous()

        def gather(t):
            buf = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(buf, t.contiguous())
            return torch.cat(buf, 0).cpu()

        all_zs, all_zt, losses = gather(zs), gather(zt), gather(loss)
        text_ids = gather(torch.tensor(text_idx, device=dev))
        prot_ids = gather(seq_idx[:b].to(dev))
        dsets = gather(inputs["dataset_id"].to(dev))
        mask = conflict_matrix(text_ids, prot_ids, 0, dsets)
        assert not bool(mask.all()), "the batch is meant to hold cross-rank false negatives"
        all_s, all_t = F.normalize(all_zs, dim=-1), F.normalize(all_zt, dim=-1)
        temp = float(m.contrastive_head.temperature)
        for r in range(world):
            ref = infonce(all_zs[r * b:(r + 1) * b], all_zt[r * b:(r + 1) * b], temperature=temp, all_s=all_s,
                          all_t=all_t, mask=mask, rank=r)
            assert abs(float(losses[r]) - float(ref)) < 2e-3 * max(1.0, abs(float(ref))), (r, float(losses[r]), float(ref))
        # ... and the un-gathered loss of the same rows is a different number (the exchange really happened)
        local = infonce(all_zs[rank * b:(rank + 1) * b], all_zt[rank * b:(rank + 1) * b], temperature=temp)
        assert abs(float(local) - float(losses[rank])) > 1e-3
        ret[rank] = float(losses[rank])
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_gathered_infonce_forward_nccl_two_ranks(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, 29671, ret), nprocs=2, join=True)
    assert len(ret) == 2 and ret[0] != ret[1]

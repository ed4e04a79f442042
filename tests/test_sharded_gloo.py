"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: block partition, wrap-around padding, gather order.
The per-rank compute is injected (oracle functions), so this runs without a GPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_prot, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.fusion import cosine_scores
        from procyon_b200.inference.sharded import encode_proteins_sharded, shard_bounds, sharded_scores

        torch.manual_seed(0)
        toks = torch.randint(4, 24, (n_prot, 6))
        proj = torch.randn(6, 5)
        enc = lambda t: t.float() @ proj  # stands in for the encoder: a deterministic function of the tokens
        full = encode_proteins_sharded(enc, toks)
        assert full.shape == (n_prot, 5)
        torch.testing.assert_close(full, enc(toks))
        lo, hi, per = shard_bounds(n_prot, world, rank)
        local = encode_proteins_sharded(enc, toks, gather=False)
        torch.testing.assert_close(local, enc(toks[lo:hi]))
        db = torch.randn(n_prot, 5)
        q = torch.randn(3, 5)
        s = sharded_scores(cosine_scores, q, db)
        torch.testing.assert_close(s, cosine_scores(q, db))
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_prot", [8, 7, 1])
def test_sharded_encode_and_scores_world2(n_prot):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + n_prot
    mp.spawn(_worker, args=(2, port, n_prot, ret), nprocs=2, join=True)
    assert ret.get(0) == 1 and ret.get(1) == 1

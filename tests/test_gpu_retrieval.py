"""Retrieval scoring + fused top-k (BASELINE config 4: 1 query vs a 20 000-protein embedding database) on the GPU
against the oracle's `cosine_scores` (procyon/data/inference_utils.py:955-970 restated) and a host ranking."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rank(scores, k):
    """Reference ranking: argsort descending (ties: lower row first), first k."""
    order = sorted(range(scores.numel()), key=lambda i: (-float(scores[i]), i))
    return order[:k]


@pytest.mark.parametrize("N,d,Q,dtype,k", [
    (20000, 1280, 1, torch.float32, 20), (20000, 2560, 1, torch.float32, 20), (20000, 1280, 1, torch.bfloat16, 20),
    (20000, 2560, 3, torch.float32, 32), (5000, 640, 2, torch.float32, 7), (3001, 96, 5, torch.float32, 20),
    (777, 1280, 9, torch.bfloat16, 1), (33, 1280, 1, torch.float32, 32), (1, 2560, 4, torch.float32, 5),
])
def test_scores_and_fused_topk_match_oracle(cuda_device, N, d, Q, dtype, k):
    from oracle.fusion import cosine_scores as o_cos
    from procyon_b200.data.inference_utils import cosine_scores, retrieval_scores_topk

    g = torch.Generator().manual_seed(N + d + Q)
    db = torch.randn(N, d, generator=g).to(dtype)
    if N > 40:
        db[17] = 0  # F.normalize eps path
        db[29] = db[3] * 2.5  # same direction as row 3: equal cosine up to rounding
    q = torch.randn(Q, d, generator=g)
    ref = o_cos(q, db.float())
    dbc = db.cuda()
    plain = cosine_scores(q, dbc)
    tol = 1e-5 if dtype == torch.float32 else 1e-4
    torch.testing.assert_close(plain.cpu(), ref, rtol=1e-4, atol=tol)
    for rep in range(3):  # the ticket must come back to zero: repeated launches rank correctly
        scores, val, idx = retrieval_scores_topk(q, dbc, k, index_base=1000 * rep)
        assert torch.equal(scores, plain)  # the fused launch computes the very same scores
        val, idx = val.cpu(), idx.cpu()
        for qi in range(Q):
            want = _rank(scores[qi].cpu(), k)  # ranking of the device's own scores: must be EXACT
            n_real = min(k, N)
            assert idx[qi, :n_real].tolist() == [w + 1000 * rep for w in want], (rep, qi)
            assert torch.equal(val[qi, :n_real], scores[qi].cpu()[want])
            assert (idx[qi, n_real:] == -1).all() and torch.isinf(val[qi, n_real:]).all()
            # and against the oracle's scores wherever its own ranking is numerically unambiguous
            ro = ref[qi].sort(descending=True)
            gaps = ro.values[: n_real] - ro.values[1 : n_real + 1] if N > n_real else None
            if gaps is not None and bool((gaps > 10 * tol).all()):
                assert idx[qi, :n_real].tolist() == [int(i) + 1000 * rep for i in ro.indices[:n_real]]


def test_get_proteins_from_embedding_uses_fused_ranking(cuda_device):
    """Public API: top_k <= 32 -> one launch for scores + ranking; top_k=None -> the whole ranking."""
    from oracle.fusion import cosine_scores as o_cos
    from procyon_b200 import _lib
    from procyon_b200.data.inference_utils import get_proteins_from_embedding

    g = torch.Generator().manual_seed(99)
    db = torch.randn(20000, 1280, generator=g)
    q = torch.randn(1, 1280, generator=g)
    ref = o_cos(q, db)[0]
    dbc = db.cuda()
    lib = _lib.load()
    n0 = lib.pcy_launch_count()
    df = get_proteins_from_embedding(dbc, query_embeddings=q, top_k=20)
    assert lib.pcy_launch_count() - n0 == 1
    assert df["index"].tolist() == ref.argsort(descending=True)[:20].tolist()
    torch.testing.assert_close(torch.tensor(df["sim_score"].tolist()), ref.sort(descending=True).values[:20],
                               rtol=1e-4, atol=1e-5)
    full = get_proteins_from_embedding(dbc, query_embeddings=q, top_k=None)
    assert len(full) == 20000 and full["index"].tolist()[:50] == ref.argsort(descending=True)[:50].tolist()
    many = get_proteins_from_embedding(dbc, query_embeddings=q, top_k=100)
    assert many["index"].tolist() == ref.argsort(descending=True)[:100].tolist()


def test_sharded_index_topk_single_rank_and_merge(cuda_device):
    """ShardedProteinIndex.topk without a process group, and the candidate merge (`pcy_topk_merge`) on candidates laid
    out as 8 shards would send them (incl. a short last shard padded with -1)."""
    import ctypes

    from oracle.fusion import cosine_scores as o_cos
    from procyon_b200 import _lib
    from procyon_b200._lib import c_int, check, ptr, stream_ptr
    from procyon_b200.data.inference_utils import ShardedProteinIndex, retrieval_scores_topk

    g = torch.Generator().manual_seed(5)
    N, d, k, W = 20003, 1280, 20, 8
    db = torch.randn(N, d, generator=g)
    q = torch.randn(2, d, generator=g)
    ref = o_cos(q, db)
    index = ShardedProteinIndex(db, "cuda")
    val, idx = index.topk(q, k)
    for qi in range(2):
        assert idx[qi].tolist() == ref[qi].argsort(descending=True)[:k].tolist()
    # 8 shards on one GPU: per-shard fused ranking with the shard's global row offset, then the merge
    per = (N + W - 1) // W
    vals, idxs = [], []
    for r in range(W):
        lo, hi = r * per, min(N, (r + 1) * per)
        _, v, i = retrieval_scores_topk(q, db[lo:hi].cuda(), k, index_base=lo)
        vals.append(v), idxs.append(i)
    cand_val = torch.stack(vals, 1).reshape(2, W * k).contiguous()
    cand_idx = torch.stack(idxs, 1).reshape(2, W * k).contiguous()
    out_val = torch.empty((2, k), device="cuda")
    out_idx = torch.empty((2, k), device="cuda", dtype=torch.int32)
    check(_lib.load().pcy_topk_merge(ptr(cand_val), ptr(cand_idx), c_int(2), c_int(W * k), c_int(k), ptr(out_val),
                                     ptr(out_idx), stream_ptr(torch.device("cuda"))), "pcy_topk_merge")
    assert torch.equal(out_idx.cpu().long(), idx.cpu()) and torch.equal(out_val, val)

"""Integer bookkeeping of the persistent beam-search decode kernel (csrc/decode_rows_megakernel.cu), replayed on the CPU
by scripts/sim_rows_partition.py: the stream-K / group-aligned cut of every weight matrix over CTAs and warps must cover
every 16 x 256 tile exactly once, finish every output group exactly once, and stay inside the shared-memory pool and the
piece slots the host plan promises - for the SM count of a B200 and for others."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("sim_rows_partition", os.path.join(ROOT, "scripts", "sim_rows_partition.py"))
sim = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sim)

SHAPES = {
    "llama8b": dict(d=4096, f=14336, H=32, KVH=8, V=16 * 1000 + 7),   # (a shortened vocabulary keeps the test fast)
    "gq4": dict(d=1024, f=1024, H=8, KVH=2, V=1003),
    "gq4wide": dict(d=1024, f=8448, H=8, KVH=2, V=1003),
}


@pytest.mark.parametrize("name", list(SHAPES))
@pytest.mark.parametrize("nb", [148, 132, 160])
def test_every_tile_once_every_group_once(name, nb):
    c = SHAPES[name]
    kcap = max(c["d"], c["H"] * 128)
    qkv = (c["H"] + 2 * c["KVH"]) * 128
    mats = dict(qkv=(qkv, c["d"], False), o=(c["d"], c["H"] * 128, False), gu=(2 * c["f"], c["d"], True),
                down=(c["d"], c["f"], False), lm=(c["V"], c["d"], False))
    ran = 0
    for label, (N, K, swiglu) in mats.items():
        if not sim.plan_ok(N, K, kcap, swiglu, nb):
            continue  # the host plan sends this shape to the per-op path on such a device
        r = sim.simulate(N, K, kcap, swiglu, nb)   # (asserts coverage, single finish, pool and slot bounds inside)
        assert r["max_pool_per_warp"] <= sim.PW
        if not swiglu:
            sim.simulate(N, K, kcap, swiglu, nb, aligned=True)
        ran += 1
    assert ran >= 3


def test_llama8b_on_b200_is_fully_supported():
    c = dict(d=4096, f=14336, H=32, KVH=8, V=128263)
    kcap = 4096
    for N, K, swiglu in ((6144, 4096, False), (4096, 4096, False), (28672, 4096, True), (4096, 14336, False),
                         (c["V"], 4096, False)):
        assert sim.plan_ok(N, K, kcap, swiglu, 148)
    g = sim.make_geom(4096, 14336, kcap)
    assert (g["KQ"], g["ckq"]) == (4, 14)  # the down projection is cut into 4 k-parts of 14 chunks

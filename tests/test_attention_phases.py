"""CPU replay of the mbarrier protocol of the ESM2 attention kernels with O accumulated in TMEM (kernels 5 / 6 and the
persistent kernel 7 of procyon_b200/csrc/attention_tc.cu; scripts/sim_attention_phases.py): the shipped waiting rules
never alias a phase, dead-lock or let a TMEM region be overwritten under a reader, whatever the interleaving; the
variant that skipped phases of the P.V barrier (first version of kernel 5, seen on B200 as run-to-run differences) is
caught by the same replay."""
import importlib.util
import os
import sys

import pytest

_spec = importlib.util.spec_from_file_location(
    "sim_attention_phases", os.path.join(os.path.dirname(os.path.dirname(__file__)), "scripts", "sim_attention_phases.py"))
sim = importlib.util.module_from_spec(_spec)
sys.modules[_spec.name] = sim  # dataclasses resolve the defining module through sys.modules
_spec.loader.exec_module(sim)


@pytest.mark.parametrize("n_kv", [1, 2, 3, 5, 9, 16, 17])
def test_one_item_per_cta_rules_are_clean(n_kv):
    # kernels 5 / 6: one (protein, head, query tile) item per CTA; 9 steps = 512 residues + BOS + EOS
    for seed in range(12):
        sim.run(n_kv=n_kv, n_items=1, persistent=False, seed=seed)


@pytest.mark.parametrize("n_kv,n_items", [(1, 6), (2, 5), (3, 4), (9, 4), (5, 7)])
def test_persistent_cta_rules_are_clean(n_kv, n_items):
    # kernel 7: running phase counters across items, o_free before the next item's first P.V, Q rewritten per item
    for seed in range(12):
        sim.run(n_kv=n_kv, n_items=n_items, persistent=True, seed=seed)


def test_skipping_phases_of_the_pv_barrier_is_caught():
    caught, messages = 0, set()
    for seed in range(60):
        try:
            sim.run(n_kv=9, n_items=1, persistent=False, wait_every_o_phase=False, seed=seed)
        except sim.ProtocolError as e:
            caught += 1
            messages.add(str(e).split(":")[0])
    assert caught >= 30, caught
    assert "aliasing" in messages  # "P.V(n-2) still running" taken for "P.V(n-1) done"


def test_parity_wait_semantics():
    b = sim.Barrier("b", 2)
    assert not b.parity_passes(0) and b.parity_passes(1)  # phase 0 in progress
    b.arrive()
    assert not b.parity_passes(0)
    b.arrive()  # phase 0 complete, phase 1 in progress
    assert b.parity_passes(0) and not b.parity_passes(1)
    b.arrive(); b.arrive()  # two phases on: a waiter for phase 0 would now block (same parity as phase 2)
    assert not b.parity_passes(0)


@pytest.mark.parametrize("persistent,n_items", [(False, 1), (True, 4)])
def test_kv_ring_refill_distance(persistent, n_items):
    """KV2_AHEAD: tiles requested 3 (shipped) or 2 steps ahead of the S = Q K^T that consumes them never overwrite a
    stage under a pending P.V and always find the right tile; 1 dead-locks, because S(j+1) is issued before the
    iteration's load; 4 = the ring depth waits for a P.V that cannot have been issued yet."""
    for ahead in (3, 2):
        for n_kv in (1, 2, 3, 4, 5, 9, 17):
            for seed in range(6):
                sim.run(n_kv=n_kv, n_items=n_items, persistent=persistent, seed=seed, ahead=ahead)
    for ahead in (1, 4):
        with pytest.raises(sim.ProtocolError):
            sim.run(n_kv=9, n_items=n_items, persistent=persistent, seed=0, ahead=ahead)

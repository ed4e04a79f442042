"""CPU replay of the mbarrier protocol of the ESM2 attention kernels with O accumulated in TMEM (kernels 5 / 6 and the
persistent kernel 7, procyon_b200/csrc/attention_tc.cu): one MMA / TMA thread and the softmax threads of a CTA as
coroutines under a random scheduler, the tensor pipe as an in-order queue whose instructions retire after random delays.

An mbarrier parity wait can only tell the CURRENT phase from the one before it.  The replay keeps the true phase index of
every barrier next to the parity bit the kernel sees and reports
  * aliasing   - a wait that passes although the phase it was meant for has not completed (round 2, first version of
                 kernel 5: softmax threads that skipped phases of o_full took "P.V(n-2) still running" for "P.V(n-1)
                 done" and read O early; seen on B200 as run-to-run differences),
  * deadlock   - a wait that can never pass because its barrier has moved two phases on,
  * hazards    - a TMEM region written while a reader of its previous contents has not finished (S / P buffers, O, Q),
so that the bookkeeping of a new variant can be checked before it costs GPU time.  `wait_every_o_phase=False` replays the
broken kernel-5 variant; tests/test_attention_phases.py asserts that it is caught and that the shipped rules are clean.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

STAGES = 4  # K / V ring (KV2_STAGES)


class ProtocolError(AssertionError):
    pass


@dataclass
class Barrier:
    name: str
    count: int
    phase: int = 0  # index of the phase in progress
    pending: int = 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending = 0
            self.phase += 1

    def parity_passes(self, parity: int) -> bool:
        # try_wait.parity: true iff the phase with this parity is not the one in progress
        return (self.phase & 1) != (parity & 1)


@dataclass
class Cta:
    n_kv: int
    n_items: int
    n_soft: int
    wait_every_o_phase: bool
    persistent: bool
    rng: random.Random
    ahead: int = STAGES - 1  # K / V tiles requested ahead of the S = Q K^T that consumes them
    stage_tile: dict = field(default_factory=dict)  # ring stage -> global tile it holds
    bars: dict = field(default_factory=dict)
    pipe: list = field(default_factory=list)  # in-order tensor pipe: (kind, payload)
    # TMEM bookkeeping: who still has to read what
    s_readers: dict = field(default_factory=dict)  # global step -> threads that have not read S(g) yet
    p_written: dict = field(default_factory=dict)  # global step -> threads that have stored P(g)
    o_item_readers: dict = field(default_factory=dict)  # item -> threads that have not read its final O
    o_steps_done: int = 0  # P.V MMAs retired (global count)
    s_done: int = 0  # S MMAs retired (global count)
    q_item: int = -1  # item whose Q is complete in TMEM
    q_writers: dict = field(default_factory=dict)

    def __post_init__(self):
        n = self.n_soft
        self.bars = {
            "q_full": Barrier("q_full", n), "o_full": Barrier("o_full", 1), "o_free": Barrier("o_free", n),
            "s_full0": Barrier("s_full0", 1), "s_full1": Barrier("s_full1", 1),
            "p_ready0": Barrier("p_ready0", n), "p_ready1": Barrier("p_ready1", n),
        }
        for s in range(STAGES):
            self.bars[f"kv_full{s}"] = Barrier(f"kv_full{s}", 1)
            self.bars[f"kv_empty{s}"] = Barrier(f"kv_empty{s}", 1)

    # a wait: yields until the parity test passes, then checks that the intended phase really completed
    def wait(self, bar: str, parity: int, intended_phase: int, who: str):
        b = self.bars[bar]
        spins = 0
        while not b.parity_passes(parity):
            if b.phase > intended_phase + 1:
                raise ProtocolError(f"deadlock: {who} waits for phase {intended_phase} of {bar}, barrier is in {b.phase}")
            spins += 1
            if spins > 200000:
                raise ProtocolError(f"stuck: {who} at {bar} phase {intended_phase} (barrier in {b.phase})")
            yield
        if b.phase <= intended_phase:
            raise ProtocolError(f"aliasing: {who} passed {bar} parity {parity} meant for phase {intended_phase}, "
                                f"barrier is only in phase {b.phase}")

    # ---- tensor pipe: instructions retire in order, each after a random number of scheduler ticks ----
    def issue(self, kind: str, payload):
        self.pipe.append([kind, payload, self.rng.randint(0, 6)])

    def tick_pipe(self):
        if not self.pipe:
            return
        head = self.pipe[0]
        if head[2] > 0:
            head[2] -= 1
            return
        kind, payload, _ = self.pipe.pop(0)
        if kind == "S":  # S(g) written into buffer g & 1: the buffer's previous contents must be dead
            g, item = payload
            if g >= 2:
                if self.s_readers.get(g - 2):
                    raise ProtocolError(f"hazard: S({g}) overwrites S({g - 2}) before {self.s_readers[g - 2]} read it")
                if self.o_steps_done < g - 1:
                    raise ProtocolError(f"hazard: S({g}) overwrites P({g - 2}) before P.V({g - 2}) retired")
            if self.q_item != item:
                raise ProtocolError(f"hazard: S({g}) of item {item} reads Q of item {self.q_item}")
            if self.stage_tile.get(g % STAGES) != g:
                raise ProtocolError(f"hazard: S({g}) reads K of tile {self.stage_tile.get(g % STAGES)}")
            self.s_readers[g] = set(range(self.n_soft))
            self.s_done += 1
        elif kind == "PV":
            g, item, first = payload
            if len(self.p_written.get(g, ())) != self.n_soft:
                raise ProtocolError(f"hazard: P.V({g}) reads P before all threads stored it")
            if self.stage_tile.get(g % STAGES) != g:
                raise ProtocolError(f"hazard: P.V({g}) reads V of tile {self.stage_tile.get(g % STAGES)}")
            if first and item > 0 and self.o_item_readers.get(item - 1):
                raise ProtocolError(f"hazard: first P.V of item {item} overwrites O of item {item - 1} before "
                                    f"{self.o_item_readers[item - 1]} read it")
            self.o_steps_done += 1
        elif kind == "commit":
            self.bars[payload].arrive()

    # ---- the MMA / TMA thread ----
    def mma_thread(self):
        g0 = kl = ks = 0
        for item in range(self.n_items):
            k0 = ks

            def load_kv():
                nonlocal kl
                st = kl % STAGES
                if kl >= STAGES:
                    yield from self.wait(f"kv_empty{st}", (kl // STAGES - 1) & 1, kl // STAGES - 1, "mma")
                    if self.o_steps_done < kl - STAGES + 1:
                        raise ProtocolError(f"hazard: tile {kl} overwrites tile {kl - STAGES} before its P.V retired")
                self.stage_tile[st] = kl
                self.bars[f"kv_full{st}"].arrive()  # TMA completes the transaction count
                kl += 1

            def issue_s(t):
                nonlocal ks
                st = ks % STAGES
                yield from self.wait(f"kv_full{st}", (ks // STAGES) & 1, ks // STAGES, "mma")
                g = g0 + t
                self.issue("S", (g, item))
                self.issue("commit", f"s_full{g & 1}")
                ks += 1

            for t in range(min(self.ahead, self.n_kv)):
                yield from load_kv()
            yield from self.wait("q_full", item & 1, item, "mma")
            yield from issue_s(0)
            for j in range(self.n_kv):
                if j + 1 < self.n_kv:
                    yield from issue_s(j + 1)
                if j + self.ahead < self.n_kv:
                    yield from load_kv()
                g = g0 + j
                yield from self.wait(f"p_ready{g & 1}", (g >> 1) & 1, g >> 1, "mma")
                if j == 0 and item > 0:
                    yield from self.wait("o_free", (item - 1) & 1, item - 1, "mma")
                self.issue("PV", (g, item, j == 0))
                self.issue("commit", f"kv_empty{(k0 + j) % STAGES}")
                self.issue("commit", "o_full")
            g0 += self.n_kv

    # ---- one softmax thread (a warp in the kernel: all its lanes move together) ----
    def softmax_thread(self, tid: int):
        g0 = 0
        who = f"softmax{tid}"
        for item in range(self.n_items):
            # Q of this item -> TMEM: every S MMA of the previous item must have retired
            if item > 0 and self.s_done < g0:
                raise ProtocolError(f"hazard: {who} overwrites Q while S MMAs of item {item - 1} are in flight")
            self.q_writers.setdefault(item, set()).add(tid)
            if len(self.q_writers[item]) == self.n_soft:
                self.q_item = item
            self.bars["q_full"].arrive()
            self.o_item_readers.setdefault(item, set(range(self.n_soft)))
            for j in range(self.n_kv):
                g = g0 + j
                yield from self.wait(f"s_full{g & 1}", (g >> 1) & 1, g >> 1, who)
                self.s_readers[g].discard(tid)  # scores in registers
                yield
                self.p_written.setdefault(g, set()).add(tid)  # P(g) over this thread's own scores
                rescale = self.rng.random() < 0.15
                if j > 0 and (self.wait_every_o_phase or rescale):
                    yield from self.wait("o_full", (g - 1) & 1, g - 1, who)
                    if rescale and self.o_steps_done < g:
                        raise ProtocolError(f"hazard: {who} rescales O before P.V({g - 1}) retired")
                self.bars[f"p_ready{g & 1}"].arrive()
                for _ in range(self.rng.randint(0, 3)):
                    yield
            g_last = g0 + self.n_kv - 1
            # a fast thread (e.g. one whose last step was fully masked) gets here early
            yield from self.wait("o_full", g_last & 1, g_last, who)
            if self.o_steps_done < g_last + 1:
                raise ProtocolError(f"hazard: {who} reads the final O of item {item} before P.V({g_last}) retired")
            self.o_item_readers[item].discard(tid)
            if self.persistent:
                self.bars["o_free"].arrive()
            g0 += self.n_kv


def run(n_kv: int, n_items: int, n_soft: int = 4, wait_every_o_phase: bool = True, persistent: bool = True,
        seed: int = 0, max_ticks: int = 2_000_000, ahead: int = STAGES - 1) -> None:
    """Replays one CTA; raises ProtocolError on aliasing / deadlock / hazard."""
    rng = random.Random(seed)
    cta = Cta(n_kv=n_kv, n_items=n_items, n_soft=n_soft, wait_every_o_phase=wait_every_o_phase, persistent=persistent,
              rng=rng, ahead=ahead)
    threads = [cta.mma_thread()] + [cta.softmax_thread(t) for t in range(n_soft)]
    alive = list(range(len(threads)))
    for _ in range(max_ticks):
        if not alive and not cta.pipe:
            return
        # a random runnable thread advances to its next yield; the tensor pipe advances independently
        if alive and (not cta.pipe or rng.random() < 0.7):
            i = rng.choice(alive)
            try:
                next(threads[i])
            except StopIteration:
                alive.remove(i)
        else:
            cta.tick_pipe()
    raise ProtocolError("simulation did not finish")


if __name__ == "__main__":
    for n_kv in (1, 2, 3, 5, 9, 17):
        for seed in range(20):
            run(n_kv=n_kv, n_items=4, seed=seed)
            run(n_kv=n_kv, n_items=1, persistent=False, seed=seed)
    caught = 0
    for seed in range(200):
        try:
            run(n_kv=9, n_items=1, persistent=False, wait_every_o_phase=False, seed=seed)
        except ProtocolError as e:
            caught += 1
            last = str(e)
    print("shipped rules clean; phase-skipping variant caught in", caught, "of 200 seeds, e.g.:", last)

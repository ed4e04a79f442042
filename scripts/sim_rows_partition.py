"""CPU model of the work partition of csrc/decode_rows_megakernel.cu (chunk ranges, warp spans, pool pieces, cross-CTA
pieces and tickets): checks that every weight tile is consumed exactly once, that every output group is finished exactly
once, and that the pool / piece-slot bounds the host plan promises hold.  Pure integer arithmetic, no GPU."""
import sys
from collections import defaultdict

RW, PW, TR, TK = 8, 3, 16, 256


def make_geom(N, K, kcap, aligned=False):
    n_rg = (N + TR - 1) // TR
    ck = K // TK
    kq0 = (K + kcap - 1) // kcap
    ckq = (ck + kq0 - 1) // kq0
    KQ = (ck + ckq - 1) // ckq
    return dict(n_rg=n_rg, ck=ck, KQ=KQ, ckq=ckq, per=n_rg * ckq, T=n_rg * ck, unit=ck if (aligned and KQ == 1) else 1)


def part_len(g, q):
    return g["ck"] - q * g["ckq"] if q == g["KQ"] - 1 else g["ckq"]


def seg_start(g, q, rg):
    return q * g["per"] + rg * part_len(g, q)


def owner(c, T, nb):
    return ((c + 1) * nb - 1) // T


def lo_of(T, i, nb):
    return T * i // nb


def cta_lo(g, i, nb):
    """first chunk of CTA i: stream-K cut, or (unit = chunks per row group) at row-group boundaries"""
    return (g["T"] // g["unit"]) * i // nb * g["unit"]


def plan_ok(N, K, kcap, swiglu, nb):
    """the host-side guard of plan_matrix()"""
    g = make_geom(N, K, kcap)
    cmax = (g["T"] + nb - 1) // nb
    if g["KQ"] == 1:
        cmax = max(cmax, (g["n_rg"] + nb - 1) // nb * g["ck"])
    span = (cmax + RW - 1) // RW
    len_min = min(part_len(g, q) for q in range(g["KQ"]))
    pieces = 1 if span <= 1 else (span + len_min - 2) // len_min + 1
    if g["KQ"] == 1 and not swiglu:
        pieces = min(pieces, 2)
    return pieces <= PW


def simulate(N, K, kcap, swiglu, nb, aligned=False):
    g = make_geom(N, K, kcap, aligned)
    T = g["T"]
    seen = defaultdict(int)
    tickets = defaultdict(int)
    finished = defaultdict(int)
    pieces_written = set()
    maxcp = 0
    for q in range(g["KQ"]):
        ln = part_len(g, q)
        for rg in range(g["n_rg"]):
            s = seg_start(g, q, rg)
            maxcp = max(maxcp, owner(s + ln - 1, T, nb) - owner(s, T, nb) + 1)
    nseg = 2 if swiglu else g["KQ"]
    og_total = 2 * g["ck"] if swiglu else g["ck"]
    max_np = 0
    ring_positions = 0
    for bid in range(nb):
        lo, hi = cta_lo(g, bid, nb), cta_lo(g, bid + 1, nb)
        a = lo
        while a < hi:
            q = min(a // g["per"], g["KQ"] - 1)
            b = min(hi, T if q == g["KQ"] - 1 else (q + 1) * g["per"])
            ln, qbase = part_len(g, q), q * g["per"]
            C = b - a
            L = (C + RW - 1) // RW
            ring_positions += L * RW
            prg = {}
            direct = set()
            for w in range(RW):
                w_lo, w_hi = a + C * w // RW, a + C * (w + 1) // RW
                assert w_hi - w_lo <= L
                cur, cnt, np_ = -1, 0, 0

                def flush():
                    nonlocal np_
                    whole = cnt == ln
                    if whole and g["KQ"] == 1 and not swiglu:
                        direct.add(cur)
                        finished[cur] += 1
                    else:
                        assert np_ < PW, ("pool overflow", N, K, bid, w)
                        prg[w * PW + np_] = cur
                        np_ += 1

                for c in range(w_lo, w_hi):
                    rem = c - qbase
                    rgi, kl = rem // ln, rem % ln
                    kc = q * g["ckq"] + kl
                    assert 0 <= rgi < g["n_rg"] and kc < g["ck"]
                    seen[(rgi, kc)] += 1
                    if rgi != cur:
                        if cur >= 0:
                            flush()
                        cur, cnt = rgi, 0
                    cnt += 1
                if cur >= 0:
                    flush()
                max_np = max(max_np, np_)
            rg_first, rg_last = (a - qbase) // ln, (b - 1 - qbase) // ln
            og_first, og_last = (rg_first // 2, rg_last // 2) if swiglu else (rg_first, rg_last)
            for og in range(og_first, og_last + 1):
                any_ = False
                contrib = 0
                segn = []
                for s in range(2 if swiglu else 1):
                    rgi = 2 * og + s if swiglu else og
                    if rgi in prg.values():
                        any_ = True
                    s_lo = qbase + rgi * ln
                    n = max(0, min(b, s_lo + ln) - max(a, s_lo))
                    segn.append(n)
                    contrib += n
                if not any_:
                    assert not swiglu and og in direct
                    continue
                og_c0 = 2 * og * g["ck"] if swiglu else og * g["ck"]
                if g["KQ"] == 1 and og_c0 >= lo and og_c0 + og_total <= hi:
                    finished[og] += 1
                    continue
                assert g["unit"] == 1, "a group-aligned cut never splits a group between CTAs"
                for s in range(2 if swiglu else 1):
                    if segn[s] == 0:
                        continue
                    rgi = 2 * og + s if swiglu else og
                    sidx = s if swiglu else q
                    pi = bid - owner(qbase + rgi * ln, T, nb)
                    assert 0 <= pi < maxcp, ("piece slot", pi, maxcp)
                    key = (og * nseg + sidx, pi)
                    assert key not in pieces_written
                    pieces_written.add(key)
                tickets[og] += contrib
                assert tickets[og] <= og_total
                if tickets[og] == og_total:
                    # merger: every slot it reads must have been written
                    for sidx in range(nseg):
                        qq = 0 if swiglu else sidx
                        rgi = 2 * og + sidx if swiglu else og
                        s_lo, s_len = seg_start(g, qq, rgi), part_len(g, qq)
                        first, last = owner(s_lo, T, nb), owner(s_lo + s_len - 1, T, nb)
                        for pi in range(last - first + 1):
                            if lo_of(T, first + pi + 1, nb) <= lo_of(T, first + pi, nb):
                                continue
                            assert (og * nseg + sidx, pi) in pieces_written, ("missing piece", og, sidx, pi)
                    finished[og] += 1
            a = b
    n_og = g["n_rg"] // 2 if swiglu else g["n_rg"]
    assert all(seen[(rg, kc)] == 1 for rg in range(g["n_rg"]) for kc in range(g["ck"])), "tile coverage"
    assert len(seen) == g["n_rg"] * g["ck"]
    assert all(finished[og] == 1 for og in range(n_og)), [og for og in range(n_og) if finished[og] != 1][:5]
    return dict(geom=g, maxcp=maxcp, max_pool_per_warp=max_np, ring_positions_per_cta=ring_positions / nb, chunks_per_cta=T / nb)


if __name__ == "__main__":
    nb = 148
    shapes = {
        "llama8b": dict(d=4096, f=14336, H=32, KVH=8, V=128263),
        "gq4": dict(d=1024, f=1024, H=8, KVH=2, V=1003),
        "gq4wide": dict(d=1024, f=8448, H=8, KVH=2, V=1003),
        "fullwidth1layer": dict(d=4096, f=14336, H=32, KVH=8, V=128263),
    }
    for name, c in shapes.items():
        kcap = max(c["d"], c["H"] * 128)
        qkv = (c["H"] + 2 * c["KVH"]) * 128
        for label, (N, K, sw) in dict(qkv=(qkv, c["d"], False), o=(c["d"], c["H"] * 128, False), gu=(2 * c["f"], c["d"], True),
                                      down=(c["d"], c["f"], False), lm=(c["V"], c["d"], False)).items():
            for n in (nb, 160, 132, 64, 7):
                if not plan_ok(N, K, kcap, sw, n):
                    print(name, label, "not supported on", n, "SMs (per-op path)")
                    continue
                r = simulate(N, K, kcap, sw, n)
                if not sw:
                    simulate(N, K, kcap, sw, n, aligned=True)
                if n == nb:
                    print(name, label, r)
    print("ok")

"""The algorithm of the long-piece kernels (long_piece_warp / long_piece_block in csrc/b200bpe.cu), as a small
executable model, against the oracle's literal min-rank loop.

`_byte_pair_merge` (src/lib.rs:140-196) merges ONE pair per step: the smallest rank, leftmost on ties.  The kernels
merge, in one ROUND, every pair of the current minimum rank g that the sequential loop would merge before any
other rank gets its turn: inside a chain of overlapping g-pairs every second one (a merge destroys the g-pair to
its right), left to right -- but only up to and including the first merge that creates a NEW pair of rank < g
("violation"): the sequential loop would turn to that pair next, so the round commits what precedes it and the
next round restarts from the exact sequential state.  This model states exactly that and is what the CUDA code was
written from; the GPU parity tests check the kernels themselves."""
import random

import vocab_util as vu
from oracle import Oracle

MAX = 0xFFFFFFFF
PSEUDO = 0xFFFFFF00


def build_pairs(ranks):
    pair = {}
    known = set(ranks) | {bytes([i]) for i in range(256)}

    def idof(b):
        return ranks.get(b, PSEUDO + b[0] if len(b) == 1 else None)

    for t, r in ranks.items():
        for k in range(1, len(t)):
            a, b = t[:k], t[k:]
            if a in known and b in known:
                pair[(idof(a), idof(b))] = r
    return pair, idof


def rounds_encode(piece, ranks, pair, idof, stats):
    if piece in ranks:                                        # whole-piece probe, src/lib.rs:367-368
        return [ranks[piece]]
    ids = [idof(bytes([b])) for b in piece]
    rk = [pair.get((ids[i], ids[i + 1]), MAX) for i in range(len(ids) - 1)] + [MAX]
    while True:
        m, g = len(ids), min(rk)
        if g == MAX:
            return ids
        stats["rounds"] += 1
        sel, run = [False] * m, 0                             # every second pair of a chain of overlapping g-pairs
        for i in range(m):
            if rk[i] == g:
                run += 1
                sel[i] = (run & 1) == 1
            else:
                run = 0
        nl, nr, viol = [None] * m, [None] * m, None           # ranks of the pairs a merge at i would create
        for i in range(m):
            if not sel[i]:
                continue
            left = g if (i >= 2 and sel[i - 2]) else (ids[i - 1] if i >= 1 else None)
            nl[i] = pair.get((left, g), MAX) if i >= 1 else MAX
            nr[i] = pair.get((g, ids[i + 2]), MAX) if i + 2 < m else MAX
            if (nl[i] < g or nr[i] < g) and viol is None:
                viol = i
        stats["violations"] += viol is not None
        com = [sel[i] and (viol is None or i <= viol) for i in range(m)]
        nids, nrk = [], []
        for i in range(m):
            if i >= 1 and com[i - 1]:
                continue                                      # absorbed by the merge to its left
            if com[i]:
                nids.append(g)
                nrk.append(nl[i + 2] if (i + 2 < m and com[i + 2]) else nr[i])
            else:
                nids.append(ids[i])
                nrk.append(nl[i + 1] if (i + 1 < m and com[i + 1]) else rk[i])
        ids, rk = nids, nrk


def test_round_synchronous_merge_equals_the_sequential_loop():
    rnd = random.Random(1)
    stats = {"rounds": 0, "violations": 0}
    total = 0
    for _ in range(500):
        alpha = bytes(rnd.sample(range(97, 123), rnd.choice([1, 2, 3, 4])))
        ranks = {bytes([i]): i for i in range(256)}
        toks = set()
        for _ in range(rnd.choice([3, 8, 20, 60])):
            toks.add(bytes(rnd.choice(alpha) for _ in range(rnd.choice([2, 2, 3, 3, 4, 5, 6, 8]))))
        order = list(range(256, 256 + len(toks)))
        rnd.shuffle(order)                                    # non-monotone on purpose: ranks need not follow merge order
        for t, r in zip(sorted(toks), order):
            ranks[t] = r
        o = Oracle(ranks, {}, vu.R50K_PAT)
        pair, idof = build_pairs(ranks)
        for _ in range(20):
            piece = bytes(rnd.choice(alpha) for _ in range(rnd.choice([1, 2, 3, 5, 9, 17, 40, 99, 100, 101, 150, 400])))
            assert rounds_encode(piece, ranks, pair, idof, stats) == o.encode_single_piece(piece), (piece, ranks)
            total += 1
    assert stats["violations"] > 100                          # the early-commit rule was exercised, not just present
    assert stats["rounds"] < 40 * total                       # and rounds stay far below one per merge

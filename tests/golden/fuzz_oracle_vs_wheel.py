#!/usr/bin/env python
"""Pinning the oracle beyond the committed fixtures: random documents (tools/fuzz_cpu.py generators) split by
the oracle's literal backtracking matcher vs the REAL engine (the `tiktoken` wheel, Rust CoreBPE; splits are
read out with the all-substrings vocabulary of make_golden.py).  Build container only (needs the wheel); the
GPU box never runs this.  Usage: python tests/golden/fuzz_oracle_vs_wheel.py [seed] [seconds]"""
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), HERE]
import vocab_util as vu  # noqa: E402
from oracle import Oracle  # noqa: E402
import fuzz_cpu as F  # noqa: E402
import make_golden as G  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
    rnd = random.Random(seed)
    base = {bytes([i]): i for i in range(256)}
    pats = [vu.R50K_PAT, vu.CL100K_PAT, vu.O200K_PAT]
    oracles = [Oracle(base, {}, p) for p in pats]
    t0 = time.time()
    n = bad = 0
    while time.time() - t0 < secs and not bad:
        docs = []
        while len(docs) < 40:
            d = F.rnd_doc(rnd)
            if len(d.encode()) <= 90:
                docs.append(d)
        pid = rnd.randrange(3)
        split = G.wheel_splitter(pats[pid], [d.encode() for d in docs])
        for d in docs:
            exp, got = split(d), oracles[pid].split(d.encode())
            n += 1
            if exp != got:
                print("ORACLE != REAL ENGINE, pattern", pid, repr(d), exp, got)
                bad = 1
                break
    print(f"seed {seed}: {n} documents, {'MISMATCH' if bad else 'oracle == real engine'}")
    # the merge: random tiny-alphabet vocabularies (rank ties impossible, cascades and long tokens likely)
    from tiktoken import _tiktoken
    t0 = time.time()
    m = 0
    while time.time() - t0 < secs / 2 and not bad:
        alpha = bytes(rnd.sample(range(97, 123), rnd.choice([2, 2, 3, 4])))
        ranks = dict(base)
        toks = set()
        for _ in range(rnd.randint(3, 60)):
            toks.add(bytes(rnd.choice(alpha) for _ in range(rnd.choice([2, 2, 2, 3, 3, 4, 5, 8, 16, 17, 19, 30, 64]))))
        for t, r in zip(sorted(toks), rnd.sample(range(256, 2000), len(toks))):
            ranks[t] = r
        core = _tiktoken.CoreBPE(ranks, {}, vu.R50K_PAT)
        o = Oracle(ranks, {}, vu.R50K_PAT)
        for _ in range(60):
            piece = bytes(rnd.choice(alpha) for _ in range(rnd.choice([1, 2, 3, 9, 16, 17, 33, 64, 65, 100, 129, 200, 256, 500])))
            exp, got = core.encode_single_piece(piece), o.encode_single_piece(piece)
            m += 1
            if exp != got:
                print("ORACLE != REAL ENGINE (merge)", piece, sorted(toks), exp, got)
                bad = 1
                break
    print(f"seed {seed}: {m} pieces merged, {'MISMATCH' if bad else 'oracle == real engine'}")
    return bad


if __name__ == "__main__":
    sys.exit(main())

"""The algorithm of the segmented parallel merge kernel (csrc/kernels_pmerge.cuh), as a small executable model, against
the oracle's literal min-rank loop.

`_byte_pair_merge` (src/lib.rs:140-196) merges ONE pair per step: the smallest rank, leftmost on ties.  If merges
never created pairs, the loop would walk the pairs in (rank, position) order and take a pair unless a neighbour was
taken before it -- the greedy independent set of the path in key order, which has a closed form: pair e is taken iff
the increasing run of keys that ends at e from the left and the one that ends at e from the right both have even
length (valleys are taken, then every second pair up each slope, a peak only if both slopes agree).  The pairs that
merges DO create are the only thing that can disturb this order, so a round
  1. takes that independent set G,
  2. probes, for every member, the ranks of the two pairs its merge creates (its neighbour part being the merged
     token when the pair two places away is in G with a smaller key, else the part as it is),
  3. commits the members whose rank is strictly below T = the smallest rank any member would create -- a prefix of the
     sequential order in which no new pair can have come first -- plus, always, the global minimum (the sequential
     loop's next step whatever it creates),
and the next round starts from that exact sequential state.  This model states exactly that; the GPU parity tests
check the kernel itself."""
import random

import vocab_util as vu
from oracle import Oracle
from test_round_sync_merge_model import MAX, build_pairs


def parallel_encode(piece, ranks, pair, idof, stats):
    if piece in ranks:                                        # whole-piece probe, src/lib.rs:367-368
        return [ranks[piece]]
    ids = [idof(bytes([b])) for b in piece]
    rk = [pair.get((ids[i], ids[i + 1]), MAX) for i in range(len(ids) - 1)] + [MAX]     # rk[e]: pair (e, e+1)
    while True:
        m = len(ids)
        if min(rk) == MAX:
            return ids
        stats["rounds"] += 1
        up = [e >= 1 and rk[e - 1] <= rk[e] for e in range(m)]            # key(e-1) < key(e) in (rank, position) order
        dl, dr = [0] * m, [0] * m
        for e in range(1, m):
            dl[e] = dl[e - 1] + 1 if up[e] else 0
        for e in range(m - 2, -1, -1):
            dr[e] = 0 if up[e + 1] else dr[e + 1] + 1
        G = [rk[e] != MAX and dl[e] % 2 == 0 and dr[e] % 2 == 0 for e in range(m)]
        c1 = min(range(m), key=lambda e: (rk[e], e))
        assert G[c1]
        nl, nr = [MAX] * m, [MAX] * m
        for e in range(m):
            if not G[e]:
                continue
            M = rk[e]
            if e >= 1:
                L = rk[e - 2] if (e >= 2 and G[e - 2] and rk[e - 2] <= rk[e]) else ids[e - 1]
                nl[e] = pair.get((L, M), MAX)
            if e + 2 < m:
                R = rk[e + 2] if (G[e + 2] and rk[e + 2] < rk[e]) else ids[e + 2]
                nr[e] = pair.get((M, R), MAX)
        T = min(min(nl[e], nr[e]) for e in range(m) if G[e])
        com = [G[e] and (rk[e] < T or e == c1) for e in range(m)]
        stats["merges"] += sum(com)
        stats["held_back"] += sum(G) - sum(com)
        nids, nrk = [], []
        for e in range(m):
            if e >= 1 and com[e - 1]:
                continue                                      # absorbed by the merge to its left
            if com[e]:
                nids.append(rk[e])
                nrk.append(nl[e + 2] if (e + 2 < m and com[e + 2] and rk[e] <= rk[e + 2]) else nr[e])
            else:
                nids.append(ids[e])
                nrk.append(nl[e + 1] if (e + 1 < m and com[e + 1]) else rk[e])
        ids, rk = nids, nrk


def _random_vocab(rnd, alpha, n_tok, monotone):
    ranks = {bytes([i]): i for i in range(256)}
    toks = set()
    for _ in range(n_tok):
        toks.add(bytes(rnd.choice(alpha) for _ in range(rnd.choice([2, 2, 3, 3, 4, 5, 6, 8]))))
    order = list(range(256, 256 + len(toks)))
    if not monotone:
        rnd.shuffle(order)                                    # adversarial: ranks need not follow merge order
    for t, r in zip(sorted(toks, key=lambda t: (len(t), t)) if monotone else sorted(toks), order):
        ranks[t] = r
    return ranks


def test_parallel_merge_equals_the_sequential_loop():
    rnd = random.Random(7)
    stats = {"rounds": 0, "merges": 0, "held_back": 0}
    total = 0
    for it in range(600):
        alpha = bytes(rnd.sample(range(97, 123), rnd.choice([1, 2, 3, 4, 6])))
        ranks = _random_vocab(rnd, alpha, rnd.choice([3, 8, 20, 60, 200]), monotone=(it % 3 == 0))
        o = Oracle(ranks, {}, vu.R50K_PAT)
        pair, idof = build_pairs(ranks)
        for _ in range(20):
            piece = bytes(rnd.choice(alpha) for _ in range(rnd.choice([1, 2, 3, 5, 9, 17, 40, 99, 100, 101, 150, 400])))
            assert parallel_encode(piece, ranks, pair, idof, stats) == o.encode_single_piece(piece), (piece, ranks)
            total += 1
    assert stats["held_back"] > 1000                          # the threshold was exercised, not just present
    assert stats["merges"] > 1.5 * stats["rounds"]            # several merges per round even on these adversarial tiny alphabets (4-35 on text)


# ---- the kernel's own formulation: several pieces in one buffer, bitmaps, run parities by carry-propagating additions ----
M32 = 0xFFFFFFFF
SEP = 0xFFFFFFFE                                              # rank slot of the last part of a piece (PM_SEP)
def brev(x): return int('{:032b}'.format(x)[::-1],2)
def alt_from_start(words):
    n=32
    prev=[0]+words[:-1]
    starts=[w & ~(((w<<1)&M32) | (prev[t]>>31 if t else 0)) & M32 for t,w in enumerate(words)]
    s=[(w+(st&0x55555555)) for w,st in zip(words,starts)]
    G=0;P=0
    for t in range(n):
        if s[t]>M32: G|=1<<t
        s[t]&=M32
        if s[t]==M32: P|=1<<t
    cin=(((G|P)+G) ^ (P & ~G)) & M32
    out=[]
    for t in range(n):
        st=(s[t]+((cin>>t)&1))&M32
        ev=words[t] & ~st & M32
        out.append((ev&0x55555555)|(words[t]&~ev&0xAAAAAAAA))
    return out
def alt_from_end(words):
    r=[brev(words[31-t]) for t in range(32)]
    x=alt_from_start(r)
    return [brev(x[31-t]) for t in range(32)]


def batch_encode(pieces, ranks, pair, idof):
    ids=[];rk=[];seg=[]
    for s,p in enumerate(pieces):
        for j,b in enumerate(p):
            ids.append(idof(bytes([b]))); seg.append(s)
            rk.append(pair.get((idof(bytes([b])), idof(bytes([p[j+1]]))),MAX) if j+1<len(p) else SEP)
    m=len(ids); rounds=0
    while m:
        rk2p=rk+[MAX,MAX]
        nw=(m+31)//32
        U=[M32]*32; V=[0]*32
        for t in range(nw):
            ub=0;vb=0
            for lane in range(32):
                e=32*t+lane
                r0=rk[e] if e<m else MAX
                rm1=rk[e-1] if (e>=1 and e<m) else MAX
                if e>=m or (e>=1 and rm1<=r0): ub|=1<<lane
                if r0<SEP: vb|=1<<lane
            U[t]=ub;V[t]=vb
        if not any(V): break
        rounds+=1
        odd_l=alt_from_start(U)
        zs=[(~(((U[t]>>1)|(((U[t+1] if t<31 else M32)<<31)&M32)))&M32) for t in range(32)]
        odd_r=alt_from_end(zs)
        tb=[0]*(nw+2)
        for t in range(nw): tb[1+t]=V[t]&~odd_l[t]&~odd_r[t]&M32
        T={};c1={}
        nl=[None]*m;nr=[None]*m
        def bit(words,e): 
            if e<0: return 0
            return (words[1+(e>>5)]>>(e&31))&1 if (e>>5)<nw+1 else 0
        for e in range(m):
            if not bit(tb,e): continue
            r0=rk[e]
            has_l=e>=1 and rk[e-1]!=SEP
            has_r=rk2p[e+1]!=SEP
            vl=vr=MAX
            if has_l:
                r2=rk[e-2] if e>=2 else MAX
                L=r2 if (bit(tb,e-2) and r2<=r0) else ids[e-1]
                vl=pair.get((L,r0),MAX)
            if has_r:
                r2=rk2p[e+2]
                R=r2 if (bit(tb,e+2) and r2<r0) else ids[e+2]
                vr=pair.get((r0,R),MAX)
            nl[e]=vl; nr[e]=vr if has_r else SEP
            s=seg[e]; T[s]=min(T.get(s,MAX),vl,vr); c1[s]=min(c1.get(s,MAX),(r0<<10)|e)
        cb=[0]*(nw+2)
        for e in range(m):
            if bit(tb,e):
                s=seg[e]
                if rk[e]<T[s] or e==(c1[s]&1023): cb[1+(e>>5)]|=1<<(e&31)
        nids=[];nrk=[];nseg=[]
        for e in range(m):
            if bit(cb,e-1): continue
            r0=rk[e]
            if bit(cb,e):
                nids.append(r0); nrk.append(nl[e+2] if (nr[e]!=SEP and bit(cb,e+2) and r0<=rk2p[e+2]) else nr[e])
            else:
                nids.append(ids[e]); nrk.append(SEP if r0==SEP else (nl[e+1] if bit(cb,e+1) else r0))
            nseg.append(seg[e])
        ids,rk,seg=nids,nrk,nseg; m=len(ids)
    out=[[] for _ in pieces]
    for i,s in zip(ids,seg): out[s].append(i)
    return out


def test_bitmap_run_parities():
    """pm_alt_from_start / pm_alt_from_end (carry trick + ballot carry-lookahead over 32 words) against the definition."""
    def ref_start(bits):
        out, d = [0] * len(bits), 0
        for i, b in enumerate(bits):
            if b:
                out[i] = 1 if d % 2 == 0 else 0
                d += 1
            else:
                d = 0
        return out
    tobits = lambda words: [(words[i >> 5] >> (i & 31)) & 1 for i in range(1024)]
    rnd = random.Random(1)
    for it in range(1500):
        mode = rnd.random()
        if mode < 0.3:
            words = [rnd.getrandbits(32) for _ in range(32)]
        elif mode < 0.6:
            words = [rnd.getrandbits(32) | rnd.getrandbits(32) | rnd.getrandbits(32) for _ in range(32)]
        else:
            words = [M32 if rnd.random() < 0.7 else rnd.getrandbits(32) for _ in range(32)]
        b = tobits(words)
        assert tobits(alt_from_start(words)) == ref_start(b)
        assert tobits(alt_from_end(words)) == ref_start(b[::-1])[::-1]


def test_batched_rounds_equal_the_sequential_loop():
    """pmerge_class as the kernel runs it (pieces packed into one 512-part buffer, separator ranks, commit bitmaps)."""
    rnd = random.Random(3)
    for it in range(150):
        alpha = bytes(rnd.sample(range(97, 123), rnd.choice([1, 2, 3, 4, 6])))
        ranks = _random_vocab(rnd, alpha, rnd.choice([3, 8, 20, 60, 200]), monotone=(it % 3 == 0))
        o = Oracle(ranks, {}, vu.R50K_PAT)
        pair, idof = build_pairs(ranks)
        for _ in range(8):
            L = rnd.choice([32, 64, 128, 256])
            pieces = [bytes(rnd.choice(alpha) for _ in range(rnd.randint(L // 2 + 1, L))) for _ in range(rnd.randint(1, 512 // L))]
            pieces = [p for p in pieces if p not in ranks]
            for p, g in zip(pieces, batch_encode(pieces, ranks, pair, idof)):
                assert g == o.encode_single_piece(p), (p, ranks)

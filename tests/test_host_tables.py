"""CPU-only check of the table builder (bpe_tables.h) and of the per-thread short-piece path
(piece probe + merge_short in bpe_device.cuh) -- the same code the encode kernel runs."""
import json
import os
import random

import numpy as np

import vocab_util as vu
from oracle import Oracle
from oracle.oracle import _flatten

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAXR = 0xFFFFFFFF


def make_tables(H, ranks):
    import ctypes as C
    toks = list(ranks.keys())
    blob, off = _flatten(toks)
    rk = np.asarray([ranks[t] for t in toks], np.uint32)
    rc = C.c_int(0)
    h = H.hc_tables_new(blob.ctypes.data, off.ctypes.data, rk.ctypes.data, len(toks), C.byref(rc))
    return h, rc.value


def enc_short(H, h, piece):
    out = np.zeros(40, np.uint32)
    k = H.hc_encode_short(h, piece, len(piece), out.ctypes.data)
    return out[:k].tolist()


def test_short_path_on_reference_vectors(hostcheck):
    for voc in json.load(open(os.path.join(G, "bpe_adversarial.json"))):
        ranks = {bytes([i]): i for i in range(256)}
        ranks.update({k.encode(): v for k, v in voc["extra"].items()})
        h, rc = make_tables(hostcheck, ranks)
        assert rc == 0
        for piece, expected in voc["cases"]:
            if len(piece) <= 16:
                assert enc_short(hostcheck, h, piece.encode()) == expected
        for t, r in ranks.items():
            if len(t) > 16:
                assert hostcheck.hc_probe_long(h, t, len(t)) == r
        assert hostcheck.hc_probe_long(h, b"Q" * 20, 20) == MAXR
        hostcheck.hc_tables_free(h)


def test_short_path_with_missing_single_bytes(hostcheck):
    """A vocabulary may lack single bytes; merges THROUGH such a byte must still work
    (pseudo ids), and a lone missing byte is the reference's panic -> RANK_MAX marker here."""
    rnd = random.Random(3)
    for _ in range(300):
        alpha = bytes(rnd.sample(range(97, 123), 3))
        ranks = {bytes([i]): i for i in range(256)}
        del ranks[bytes([alpha[0]])]
        toks = set()
        for _ in range(12):
            toks.add(bytes(rnd.choice(alpha) for _ in range(rnd.choice([2, 2, 3, 4, 5]))))
        for t, r in zip(sorted(toks), rnd.sample(range(256, 400), len(toks))):
            ranks[t] = r
        o = Oracle(ranks, {}, vu.R50K_PAT)
        h, rc = make_tables(hostcheck, ranks)
        assert rc == 0
        for _ in range(25):
            piece = bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 16)))
            assert enc_short(hostcheck, h, piece) == o.encode_single_piece(piece), (piece, ranks)
        hostcheck.hc_tables_free(h)


def test_duplicate_ranks_rejected(hostcheck):
    ranks = {bytes([i]): i for i in range(256)}
    ranks[b"ab"] = 5                        # duplicate of byte 5's rank (reference: assert, lib.rs:636-641)
    h, rc = make_tables(hostcheck, ranks)
    assert h is None and rc == -3


def test_short_path_on_synthetic_cl100k(hostcheck):
    pat, ranks, special, _ = vu.load_encoding("cl100k_base", allow_real=False)
    from tools import corpus
    o = Oracle(ranks, special, pat)
    h, rc = make_tables(hostcheck, ranks)
    assert rc == 0
    pieces = {p for p in o.split(corpus.generate(corpus.ENGLISH, 9, 1 << 18).tobytes()) if len(p) <= 16}
    for p in pieces:
        assert enc_short(hostcheck, h, p) == o.encode_single_piece(p)
    hostcheck.hc_tables_free(h)


def enc_mid(H, h, piece, cap):
    out = np.zeros(cap + 8, np.uint32)
    k = H.hc_encode_mid(h, piece, len(piece), cap, out.ctypes.data)
    assert k >= 0
    return out[:k].tolist()


def test_mid_path_on_reference_vectors_and_adversarial_vocabularies(hostcheck):
    """merge_mid_conv (thread-per-piece path for 17..256-byte pieces) against the reference vectors
    and against the oracle on tiny-alphabet vocabularies that force ties, cascades and dead-run skips."""
    for voc in json.load(open(os.path.join(G, "bpe_adversarial.json"))):
        ranks = {bytes([i]): i for i in range(256)}
        ranks.update({k.encode(): v for k, v in voc["extra"].items()})
        h, rc = make_tables(hostcheck, ranks)
        assert rc == 0
        for piece, expected in voc["cases"]:
            b = piece.encode()
            if 2 <= len(b) <= 256:
                assert enc_mid(hostcheck, h, b, 256) == expected
        hostcheck.hc_tables_free(h)
    rnd = random.Random(11)
    for trial in range(120):
        alpha = bytes(rnd.sample(range(97, 123), rnd.choice([2, 3, 4])))
        ranks = {bytes([i]): i for i in range(256)}
        if trial % 5 == 0:
            del ranks[bytes([alpha[0]])]
        toks = set()
        for _ in range(rnd.randint(4, 40)):
            toks.add(bytes(rnd.choice(alpha) for _ in range(rnd.choice([2, 2, 2, 3, 3, 4, 5, 8, 19, 30]))))
        for t, r in zip(sorted(toks), rnd.sample(range(256, 600), len(toks))):
            ranks[t] = r
        o = Oracle(ranks, {}, vu.R50K_PAT)
        h, rc = make_tables(hostcheck, ranks)
        assert rc == 0
        for _ in range(30):
            n = rnd.choice([2, 3, 17, 18, 31, 33, 64, 65, 100, 128, 129, 200, 255, 256])
            piece = bytes(rnd.choice(alpha) for _ in range(n))
            cap = next(c for c in (64, 128, 256) if c >= n)
            assert enc_mid(hostcheck, h, piece, cap) == o.encode_single_piece(piece), (piece, ranks)
        hostcheck.hc_tables_free(h)


def test_mid_path_on_synthetic_o200k_cjk(hostcheck):
    pat, ranks, special, _ = vu.load_encoding("o200k_base", allow_real=False)
    from tools import corpus
    o = Oracle(ranks, special, pat)
    h, rc = make_tables(hostcheck, ranks)
    assert rc == 0
    text, off = corpus.config3(nbytes=1 << 20, seed=5)
    pieces = [p for p in o.split(text.tobytes().decode("utf-8"))]
    n = 0
    for p in pieces:
        b = p.encode() if isinstance(p, str) else p
        if 17 <= len(b) <= 256:
            cap = next(c for c in (64, 128, 256) if c >= len(b))
            assert enc_mid(hostcheck, h, b, cap) == o.encode_single_piece(b)
            n += 1
            if n >= 3000:
                break
    assert n > 500
    hostcheck.hc_tables_free(h)


def test_pair_table_lookup_is_complete_and_exact(hostcheck):
    """Every (left, right) split of every token is found with its rank by both probe forms (one at a
    time, two at a time); absent pairs miss."""
    for name in ("cl100k_base", "r50k_base"):
        pat, ranks, special, _ = vu.load_encoding(name, allow_real=False)
        h, rc = make_tables(hostcheck, ranks)
        assert rc == 0
        n_pairs = 0
        present = set()
        for t, r in ranks.items():
            for k in range(1, len(t)):
                a, b = ranks.get(t[:k]), ranks.get(t[k:])
                if a is not None and b is not None:
                    n_pairs += 1
                    present.add((a, b))
                    assert hostcheck.hc_pair_lookup(h, a, b) == r
        assert hostcheck.hc_tables_pairs(h) == n_pairs
        assert hostcheck.hc_pair_buckets(h) * 2 >= 3 * n_pairs           # load <= 1/3
        rnd = random.Random(5)
        ids = list(ranks.values())
        for _ in range(20000):
            a, b = rnd.choice(ids), rnd.choice(ids)
            if (a, b) not in present:
                assert hostcheck.hc_pair_lookup(h, a, b) == MAXR
        hostcheck.hc_tables_free(h)

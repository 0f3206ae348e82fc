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

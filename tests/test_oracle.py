"""Pins oracle/ (the CPU checker) to the reference: the Rust unit tests that need no vocabulary
file, and the fixtures generated from the real engine by tests/golden/make_golden.py."""
import hashlib
import itertools
import json
import os

import numpy as np
import pytest

import vocab_util as vu
from oracle import Oracle
from tools import corpus

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BYTES = {bytes([i]): i for i in range(256)}
PATS = {"r50k": vu.R50K_PAT, "cl100k": vu.CL100K_PAT, "o200k": vu.O200K_PAT}


def test_rust_unit_tests():
    # src/lib.rs:689-701: ranks {ab:0, cd:1}
    o = Oracle({b"ab": 0, b"cd": 1, **{bytes([i]): i + 2 for i in range(256)}}, {}, vu.R50K_PAT)
    assert o.byte_pair_split(b"abcd") == [b"ab", b"cd"]
    assert o.byte_pair_split(b"abab") == [b"ab", b"ab"]


def test_empty_string():
    # tests/test_encoding.py:83
    for pat in PATS.values():
        assert Oracle(BYTES, {}, pat).encode_ordinary("") == []


@pytest.mark.parametrize("name", ["r50k", "cl100k", "o200k"])
def test_exhaustive_split_digests(name):
    spec = json.load(open(os.path.join(G, "splits_exhaustive.json")))[name]
    o = Oracle(BYTES, {}, PATS[name])
    for l in range(1, spec["max_len"] + 1):
        h = hashlib.sha256()
        n = 0
        for tup in itertools.product(spec["alphabet"], repeat=l):
            b = "".join(tup).encode()
            h.update(b + b"\x00" + b"\x01".join(o.split(b)) + b"\x02")
            n += 1
        assert n == spec["per_len"][str(l)]["count"]
        assert h.hexdigest() == spec["per_len"][str(l)]["sha256"], f"{name}: split of some length-{l} string differs"


@pytest.mark.parametrize("name", ["r50k", "cl100k", "o200k"])
def test_random_unicode_splits(name):
    cases = json.load(open(os.path.join(G, "splits_random.json")))[name]
    o = Oracle(BYTES, {}, PATS[name])
    for text_hex, pieces_hex in cases:
        b = bytes.fromhex(text_hex)
        assert [p.hex() for p in o.split(b)] == pieces_hex, b


def test_adversarial_bpe_vectors():
    for voc in json.load(open(os.path.join(G, "bpe_adversarial.json"))):
        ranks = dict(BYTES)
        ranks.update({k.encode(): v for k, v in voc["extra"].items()})
        o = Oracle(ranks, {}, vu.R50K_PAT)
        for piece, expected in voc["cases"]:
            p = piece.encode()
            assert o.encode_single_piece(p) == expected
            # linear (lib.rs:140-196) and heap (lib.rs:47-138) algorithms agree on every length
            if len(p) > 1 and p not in ranks:
                assert o.encode_single_piece(p, force=1) == o.encode_single_piece(p, force=2) == expected


@pytest.mark.parametrize("enc,kind", [("cl100k_base", corpus.ENGLISH), ("r50k_base", corpus.ENGLISH),
                                      ("p50k_base", corpus.CODE), ("o200k_base", corpus.MIXED)])
def test_token_fixtures(enc, kind):
    pat, ranks, special, _ = vu.load_encoding(enc, allow_real=False)
    o = Oracle(ranks, special, pat)
    gold = np.load(os.path.join(G, f"tokens_{enc}.npz"))
    text = corpus.generate(kind, 31337, 96 << 10)
    assert np.array_equal(o.encode_ordinary_np(text.tobytes()), gold["corpus"])
    edge = json.load(open(os.path.join(G, "edge_texts.json")))["edge"]
    for i, s in enumerate(edge):
        assert o.encode_ordinary(s) == gold[f"edge_{i}"].tolist(), s
        assert o.encode(s, allowed_special=set(special)) == gold[f"edge_special_{i}"].tolist(), s


def test_special_slices_are_separate_haystacks():
    # "a  <|endoftext|>b": the two spaces form ONE piece because `$` sees the slice end
    ranks = dict(BYTES); ranks[b"  "] = 300
    o = Oracle(ranks, {"<|endoftext|>": 1000}, vu.CL100K_PAT)
    assert o.encode("a  <|endoftext|>b", {"<|endoftext|>"}) == [97, 300, 1000, 98]
    assert o.encode("a  b", {"<|endoftext|>"}) == [97, 32, 32, 98]
    assert o.encode("a<|endoftext|>", set())[:3] == [97, ord("<"), ord("|")]


def test_batch_driver_matches_single():
    pat, ranks, special, _ = vu.load_encoding("cl100k_base", allow_real=False)
    o = Oracle(ranks, special, pat)
    text, off = corpus.config4(n_docs=3000, seed=5)
    toks, toff = o.encode_ordinary_batch_np(text, off, n_threads=4)
    assert toff[-1] == len(toks)
    for d in (0, 1, 17, 2999):
        s, e = int(off[d]), int(off[d + 1])
        assert toks[int(toff[d]):int(toff[d + 1])].tolist() == o.encode_ordinary(text[s:e].tobytes())

#!/usr/bin/env python
"""Generate the golden fixtures that pin oracle/ (and through it the CUDA engine) to the REAL
reference engine.  Needs the `tiktoken` wheel (Rust CoreBPE) importable; run in the build
container, outputs are committed, the GPU box never runs this.

  splits_exhaustive.json   sha256 digests of the piece splits of EVERY string up to length L over
                           per-pattern alphabets (one symbol per character class the pattern uses)
  splits_random.json       random Unicode strings (hex) with their exact piece splits
  bpe_adversarial.json     random tiny vocabularies and pieces of 1..150 bytes with the wheel's
                           token ids (all 256 single bytes present: the wheel panics otherwise)
  tokens_<enc>.npz         token ids of corpus samples + edge cases for the synthetic stand-in
                           vocabularies in tests/golden/vocab/

How piece splits are read out of the wheel: with a vocabulary that contains every substring of
the text, each regex piece is a whole-piece hit (src/lib.rs:367-368) and comes back as exactly
one token, so decoding the tokens one by one yields the pieces.
"""
import hashlib, itertools, json, os, random, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tiktoken import _tiktoken            # noqa: E402  (the installed reference engine)
import tiktoken                           # noqa: E402
import vocab_util as vu                   # noqa: E402
from tools import corpus                  # noqa: E402

PATS = {"r50k": vu.R50K_PAT, "cl100k": vu.CL100K_PAT, "o200k": vu.O200K_PAT}
ALPHABETS = {
    "r50k": (["a", "s", "l", "1", " ", "\t", "\n", "!", "'"], 6),
    "cl100k": (["a", "S", "l", "1", " ", "\t", "\n", "!", "'"], 6),
    "o200k": (["a", "A", "あ", "́", "s", "1", " ", "\t", "\n", "!", "'", "/"], 5),
}
POOL = list("aAbzZ sStTdDmMlLvVeErR019 \t\n\r!?.,;:'\"/-_()[]{}<>@#$%^&*+=|\\~`") + [
    "ſ", "", " ", " ", "　", " ", "é", "É", "ǅ", "ʰ", "あ",
    "中", "́", "⃝", "ः", "٠", "²", "Ⅰ", "\U0001F600", "\U0001F3FB", "‍",
    "️", "א", "م", "가", "K", "İ", "ß", "ẞ"]


def digest_update(h, text: bytes, pieces):
    h.update(text + b"\x00" + b"\x01".join(pieces) + b"\x02")


def wheel_splitter(pat, strings_bytes):
    r = {bytes([i]): i for i in range(256)}
    n = 256
    for b in strings_bytes:
        for i in range(len(b)):
            for j in range(i + 1, len(b) + 1):
                s = b[i:j]
                if s not in r:
                    r[s] = n; n += 1
    core = _tiktoken.CoreBPE(r, {}, pat)
    inv = {v: k for k, v in r.items()}
    return lambda s: [inv[t] for t in core.encode_ordinary(s)]


def exhaustive():
    out = {}
    for name, (alph, L) in ALPHABETS.items():
        pat = PATS[name]
        r = {bytes([i]): i for i in range(256)}; n = 256
        for l in range(1, L + 1):
            for tup in itertools.product(alph, repeat=l):
                b = "".join(tup).encode()
                if b not in r:
                    r[b] = n; n += 1
        core = _tiktoken.CoreBPE(r, {}, pat)
        inv = {v: k for k, v in r.items()}
        per_len = {}
        for l in range(1, L + 1):
            h = hashlib.sha256(); cnt = 0
            for tup in itertools.product(alph, repeat=l):
                s = "".join(tup)
                digest_update(h, s.encode(), [inv[t] for t in core.encode_ordinary(s)]); cnt += 1
            per_len[str(l)] = {"count": cnt, "sha256": h.hexdigest()}
        out[name] = {"alphabet": alph, "max_len": L, "per_len": per_len}
        print("exhaustive", name, {k: v["count"] for k, v in per_len.items()}, file=sys.stderr)
    return out


def random_splits(n_per=1500):
    rnd = random.Random(20260922)
    out = {}
    for name, pat in PATS.items():
        cases = []
        for _ in range(n_per):
            L = rnd.choice([1, 2, 3, 5, 8, 13, 21, 34, 55])
            if rnd.random() < 0.5:
                s = "".join(rnd.choice(POOL) for _ in range(L))
            else:
                chars = []
                while len(chars) < L:
                    chars += [rnd.choice(POOL)] * rnd.choice([1, 1, 2, 3, 4, 7])
                s = "".join(chars[:L])
            b = s.encode()
            pieces = wheel_splitter(pat, [b])(s)
            assert b"".join(pieces) == b
            cases.append([b.hex(), [p.hex() for p in pieces]])
        out[name] = cases
        print("random splits", name, len(cases), file=sys.stderr)
    return out


def adversarial_bpe(n_vocab=250):
    rnd = random.Random(7)
    out = []
    for _ in range(n_vocab):
        alpha = bytes(rnd.sample(range(97, 123), rnd.choice([1, 2, 3, 4])))
        ranks = {bytes([i]): i for i in range(256)}
        toks = set()
        for _ in range(rnd.choice([3, 8, 20, 60])):
            toks.add(bytes(rnd.choice(alpha) for _ in range(rnd.choice([2, 2, 3, 3, 4, 5, 6, 8, 17, 24]))))
        rl = list(range(256, 256 + len(toks))); rnd.shuffle(rl)
        extra = dict(zip(sorted(toks), rl))
        ranks.update(extra)
        core = _tiktoken.CoreBPE(ranks, {}, vu.R50K_PAT)
        cases = []
        for _ in range(16):
            n = rnd.choice([1, 2, 3, 5, 9, 16, 17, 33, 64, 99, 100, 101, 150])
            piece = bytes(rnd.choice(alpha) for _ in range(n))
            cases.append([piece.decode(), core.encode_single_piece(piece)])
        out.append({"extra": {k.decode(): v for k, v in extra.items()}, "cases": cases})
    print("adversarial vocabularies", len(out), file=sys.stderr)
    return out


EDGE = ["", "a", " ", "\n", "hello world", "hello  world\n\n  x", "don't stop 'til you're DONE'S", "x" * 17, "y" * 33,
        "0" * 17, " " * 64, "\n" * 40, "a" * 300, "^" * 300, "'s" * 50, "あ" * 40,
        "日本語のテキスト、です。", "\U0001F600" * 9, "today\n ", "today\n \n",
        "today\n  \n", "rer", "'rer", " 0", "\U0001F44D", "请考试我的软件！12345",
        "hello <|endoftext|> world", "a  <|endoftext|>b"]


def token_fixtures():
    for enc, kind in [("cl100k_base", corpus.ENGLISH), ("r50k_base", corpus.ENGLISH), ("p50k_base", corpus.CODE),
                      ("o200k_base", corpus.MIXED)]:
        pat, ranks, special, src = vu.load_encoding(enc, allow_real=False)
        e = tiktoken.Encoding(enc + "_synthetic", pat_str=pat, mergeable_ranks=ranks, special_tokens=special)
        text = corpus.generate(kind, 31337, 96 << 10).tobytes().decode("utf-8")
        arrays = {"corpus": np.asarray(e.encode_ordinary(text), np.uint32)}
        for i, s in enumerate(EDGE):
            arrays[f"edge_{i}"] = np.asarray(e.encode_ordinary(s), np.uint32)
            arrays[f"edge_special_{i}"] = np.asarray(e.encode(s, allowed_special="all"), np.uint32)
        np.savez_compressed(os.path.join(HERE, f"tokens_{enc}.npz"), **arrays)
        print("tokens", enc, len(arrays["corpus"]), file=sys.stderr)


if __name__ == "__main__":
    json.dump(exhaustive(), open(os.path.join(HERE, "splits_exhaustive.json"), "w"), indent=1)
    json.dump(random_splits(), open(os.path.join(HERE, "splits_random.json"), "w"))
    json.dump(adversarial_bpe(), open(os.path.join(HERE, "bpe_adversarial.json"), "w"))
    json.dump({"edge": EDGE}, open(os.path.join(HERE, "edge_texts.json"), "w"))
    token_fixtures()

"""CPU fuzz of the code the kernels run (host build, tiktoken_b200/csrc/hostcheck.cpp) against the oracle:
  pretok  random multi-document batches -> bit-parallel pre-tokeniser (span_boundaries) vs the oracle's
          literal backtracking matcher, three patterns;
  merge   random tiny-alphabet vocabularies (ties, cascades, missing bytes) -> merge_short / merge_short_conv /
          merge_mid_conv vs the oracle's byte_pair_encode.
Usage: python tools/fuzz_cpu.py [seconds] [seed].  Test infrastructure only."""
import ctypes as C
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vocab_util as vu  # noqa: E402
from oracle import Oracle  # noqa: E402
from oracle.oracle import _flatten  # noqa: E402

H = C.CDLL(os.environ.get("B200BPE_HOSTCHECK") or os.path.join(ROOT, "tiktoken_b200", "csrc", "libb200bpe_hostcheck.so"))
H.hc_piece_starts_fast.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
H.hc_tables_new.restype = C.c_void_p
H.hc_tables_new.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
H.hc_tables_free.argtypes = [C.c_void_p]
H.hc_encode_short.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p]
H.hc_encode_mid.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]

# ASCII of every class + scalars of every class / UTF-8 length the three patterns distinguish
POOL = list("aAbzZ sStTdDmMlLvVeErR0123456789 \t\n\r\x0b\x0c!?.,;:'\"/-_()[]{}<>@#$%^&*+=|\\~`") + [
    "ſ", "", " ", " ", "　", "​", "é", "É", "ǅ", "ʰ",
    "あ", "中", "́", "⃝", "ः", "٠", "²", "Ⅰ", "\U0001F600",
    "\U0001F3FB", "‍", "️", "א", "م", "가", "K", "İ", "ß",
    "ẞ", "、", "「", "’", "“", "«", "\U00010400", "\U0001D7D8", "ก",
    "ั"]


# fragments that exercise the alternatives as wholes: contractions in every case (incl. U+017F / U+212A case
# folds), whitespace / newline mixtures, digit groups, punctuation runs with trailing newlines and slashes
FRAGS = ["'s", "'S", "'t", "'re", "'VE", "'m", "'ll", "'LL", "'d", "'\u017f", "'\u212a", "'x", "''", "'", " '",
         " ", "  ", "   ", "\t", " \t ", "\n", "\r\n", "\n\n", " \n", "\n ", " \n ", "\r", "\u3000", "\u00a0\n", "\x0b",
         "1", "12", "123", "1234", "12345678", "\uff11\uff12", "\u0663", "1\u0663" "2",
         "!", "!!", "...", "/", "//", "!/", "!\n", "!\n\n/", "-\r\n", "(", ")", "_", "\u3001", "\u300c", "\u2019",
         "word", "Word", "WORD", "wOrD", "camelCase", "HTTPServer", "\u00e9t\u00e9", "\u00c9T\u00c9", "\u4e2d\u6587", "\u0e01\u0e31",
         "a\u0301", "\u0301", "\u0301\u0301", "A\u0301b", "\u01c5x", "x\u01c5", "\u02b0", "\U0001f600", "\U0001f600\u200d\U0001f3fb"]


def rnd_doc_frags(rnd):
    return "".join(rnd.choice(FRAGS) for _ in range(rnd.choice([0, 1, 2, 3, 5, 8, 13, 30, 60])))


def rnd_doc(rnd):
    if rnd.random() < 0.4:
        return rnd_doc_frags(rnd)
    n = rnd.choice([0, 1, 2, 3, 5, 8, 13, 21, 40, 80, 200])
    if rnd.random() < 0.5:
        return "".join(rnd.choice(POOL) for _ in range(n))
    out = []
    while len(out) < n:
        out += [rnd.choice(POOL)] * rnd.choice([1, 1, 2, 3, 4, 7, 9, 33])
    return "".join(out[:n])


def fuzz_pretok(rnd, oracles):
    pid = rnd.randrange(3)
    o = oracles[pid]
    docs = [rnd_doc(rnd).encode() for _ in range(rnd.choice([1, 3, 50, 200]))]
    blob = b"".join(docs)
    n = len(blob)
    off = np.zeros(len(docs) + 1, np.uint64)
    off[1:] = np.cumsum([len(d) for d in docs])
    a = np.frombuffer(blob, np.uint8) if n else np.zeros(1, np.uint8)
    out = np.zeros(n + 2, np.uint8)
    assert H.hc_piece_starts_fast(pid, a.ctypes.data, n, off.ctypes.data, len(docs), out.ctypes.data, None) == 0
    for i, d in enumerate(docs):
        exp = np.zeros(len(d), np.uint8)
        p = 0
        for piece in o.split(d):
            exp[p] = 1
            p += len(piece)
        got = out[int(off[i]):int(off[i + 1])]
        if not np.array_equal(exp, got):
            print("PRETOK MISMATCH pattern", pid, repr(d.decode()), o.split(d), got.tolist())
            return 1
    return 0


def fuzz_merge(rnd):
    alpha = bytes(rnd.sample(range(97, 123), rnd.choice([2, 2, 3, 4])))
    ranks = {bytes([i]): i for i in range(256)}
    toks = set()
    for _ in range(rnd.randint(3, 60)):
        toks.add(bytes(rnd.choice(alpha) for _ in range(rnd.choice([2, 2, 2, 3, 3, 4, 5, 8, 16, 17, 19, 30, 64]))))
    for t, r in zip(sorted(toks), rnd.sample(range(256, 2000), len(toks))):
        ranks[t] = r
    o = Oracle(ranks, {}, vu.R50K_PAT)
    tl = list(ranks.keys())
    blob, off = _flatten(tl)
    rk = np.asarray([ranks[t] for t in tl], np.uint32)
    rc = C.c_int(0)
    h = H.hc_tables_new(blob.ctypes.data, off.ctypes.data, rk.ctypes.data, len(tl), C.byref(rc))
    assert rc.value == 0
    bad = 0
    for _ in range(40):
        n = rnd.choice([2, 3, 5, 9, 15, 16, 17, 18, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256])
        piece = bytes(rnd.choice(alpha) for _ in range(n))
        exp = o.encode_single_piece(piece)
        out = np.zeros(300, np.uint32)
        if n <= 16:
            k = H.hc_encode_short(h, piece, n, out.ctypes.data)
        else:
            k = H.hc_encode_mid(h, piece, n, next(c for c in (64, 128, 256) if c >= n), out.ctypes.data)
        got = out[:max(k, 0)].tolist()
        if k < 0 or got != exp:
            print("MERGE MISMATCH", piece, sorted(toks), exp, got, k)
            bad = 1
            break
    H.hc_tables_free(h)
    return bad


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rnd = random.Random(seed)
    base = {bytes([i]): i for i in range(256)}
    oracles = [Oracle(base, {}, p) for p in (vu.R50K_PAT, vu.CL100K_PAT, vu.O200K_PAT)]
    t0 = time.time()
    n_p = n_m = bad = 0
    while time.time() - t0 < secs and not bad:
        if rnd.random() < 0.6:
            bad |= fuzz_pretok(rnd, oracles)
            n_p += 1
        else:
            bad |= fuzz_merge(rnd)
            n_m += 1
    print(f"fuzz seed={seed}: {n_p} pretok batches, {n_m} vocabularies, {'MISMATCH' if bad else 'all equal'} "
          f"in {time.time() - t0:.0f}s")
    return bad


if __name__ == "__main__":
    sys.exit(main())

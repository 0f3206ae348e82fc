#!/usr/bin/env python
"""Development GPU check: parity of the CUDA engine against the oracle on corpus samples and
edge cases, plus stage timings.  Run under gpurun; writes gpurun_out/gpu_check.log."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from tools import corpus
from oracle import Oracle
import vocab_util as vu
from tiktoken_b200 import _tiktoken

def log(*a):
    print(*a, flush=True)

def compare(name, eng, orc, text, off, label):
    t0 = time.time()
    buf = eng.encode_ordinary_batch_buffer(text, off)
    t1 = time.time()
    got_t, got_o = np.array(buf.tokens()), np.array(buf.offsets()); buf.close()
    exp_t, exp_o = orc.encode_ordinary_batch_np(text, off, n_threads=os.cpu_count())
    ok = np.array_equal(got_t, exp_t) and np.array_equal(got_o, exp_o)
    log(f"[{name}] {label}: bytes={len(text)} docs={len(off)-1} tokens={len(exp_t)} ok={ok} wall_ms={(t1-t0)*1e3:.1f} {eng.last_timings()}")
    if not ok:
        n = min(len(got_t), len(exp_t))
        bad = np.flatnonzero(got_t[:n] != exp_t[:n])
        log("   first token mismatch at", bad[:5], "lens", len(got_t), len(exp_t))
        if len(bad):
            i = int(bad[0]); d = int(np.searchsorted(exp_o, i, side="right") - 1)
            log("   doc", d, "exp", exp_t[max(0,i-3):i+5].tolist(), "got", got_t[max(0,i-3):i+5].tolist())
            s, e = int(off[d]), int(off[d+1])
            log("   doc text head", bytes(text[s:min(e, s+120)]))
        bo = np.flatnonzero(got_o != exp_o)
        log("   offset mismatches", len(bo), bo[:5])
    return ok

def main():
    allok = True
    for name, kind in [("cl100k_base", corpus.ENGLISH), ("r50k_base", corpus.ENGLISH), ("p50k_base", corpus.CODE), ("o200k_base", corpus.MIXED)]:
        pat, ranks, sp, src = vu.load_encoding(name)
        t0 = time.time(); eng = _tiktoken.CoreBPE(ranks, sp, pat); orc = Oracle(ranks, sp, pat)
        log(f"[{name}] vocab={len(ranks)} src={src} ctor_s={time.time()-t0:.2f} tables={eng.table_bytes()}")
        # edge cases as separate docs
        edge = ["", "a", " ", "\n", "hello world", "hello  world\n\n  x", "don't stop 'til", "x" * 17, "y" * 33, "0" * 17, " " * 64,
                "\n" * 40, "a" * 1000, "^" * 300, "'s" * 50, "あ" * 40, "日本語のテキスト、です。", "\U0001F600" * 9, "", "end"]
        for n in (99, 100, 101, 255, 256, 257, 4095, 4096, 4097, 8191):
            edge.append(("ab" * n)[:n]); edge.append(" " * n); edge.append("xyz " * (n // 4))
        blob = [e.encode() for e in edge]
        text = np.frombuffer(b"".join(blob), np.uint8); off = np.zeros(len(blob) + 1, np.uint64); off[1:] = np.cumsum([len(b) for b in blob])
        allok &= compare(name, eng, orc, text, off, "edge")
        # corpus samples
        t = corpus.generate(kind, 4242, 6 << 20)
        _, off2 = corpus.docs_fixed(t, 65536, at_space=(kind == corpus.ENGLISH))
        allok &= compare(name, eng, orc, t, off2, "corpus-64k-docs")
        allok &= compare(name, eng, orc, t, np.asarray([0, len(t)], np.uint64), "corpus-1-doc")
        rng = np.random.default_rng(1)
        lens = np.clip(np.rint(rng.lognormal(np.log(90), 0.5, size=80000)), 0, 2000).astype(np.int64)
        k = int(np.searchsorted(np.cumsum(lens), len(t)))
        _, off3 = corpus.docs_from_lengths(t, lens[:k], at_space=False)
        allok &= compare(name, eng, orc, t, off3, "corpus-short-docs")
        # single piece API
        for piece in [b"hello", b"x" * 100, b"ab" * 700, b" " * 5000]:
            g = eng.encode_single_piece(piece); e = orc.encode_single_piece(piece)
            if g != e: log(f"[{name}] single piece mismatch len={len(piece)}", g[:8], e[:8]); allok = False
        # bigger timing run (no oracle): 256 MiB
        if name == "cl100k_base":
            big = corpus.generate(kind, 77, 256 << 20)
            _, offb = corpus.docs_fixed(big, 65536, True)
            for it in range(3):
                t0 = time.time(); buf = eng.encode_ordinary_batch_buffer(big, offb); dt = time.time() - t0
                log(f"[{name}] big run {it}: {len(big)/dt/1e9:.2f} GB/s e2e(pageable) tokens={buf.n_tokens} {eng.last_timings()}"); buf.close()
        del eng
    log("ALL OK" if allok else "FAILURES")
    return 0 if allok else 1

if __name__ == "__main__":
    sys.exit(main())

"""Seeded synthetic workloads for tests and bench.py.

Wraps tools/corpusgen.c (built on demand with gcc) and cuts the byte stream into documents.
Every function returns (text: np.uint8[N], doc_off: np.uint64[n_docs+1]).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
ENGLISH, MIXED, CODE = 0, 1, 2


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcorpus.so")
    src = os.path.join(_HERE, "corpusgen.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.corpus_generate.restype = C.c_int
        _LIB.corpus_generate.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    return _LIB


CHUNK = 8 << 20


def generate(kind: int, seed: int, nbytes: int) -> np.ndarray:
    """Deterministic in (kind, seed, nbytes).  english/mixed streams are produced in independent
    8 MiB chunks (chunk i uses seed*1000003+i) on a host thread pool; code is one chunk so that
    its long-piece stressors land at fixed fractions of the document."""
    out = np.empty(nbytes, dtype=np.uint8)
    lib = _lib()
    if kind == CODE or nbytes <= CHUNK:
        rc = lib.corpus_generate(kind, seed, nbytes, out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise ValueError("bad corpus kind")
        return out
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        lo = i * CHUNK
        n = min(CHUNK, nbytes - lo)
        return lib.corpus_generate(kind, seed * 1000003 + i + 1, n, C.c_void_p(out.ctypes.data + lo))

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        rcs = list(ex.map(one, range((nbytes + CHUNK - 1) // CHUNK)))
    if any(rcs):
        raise ValueError("bad corpus kind")
    return out


def generate_range(kind: int, seed: int, nbytes: int, lo: int, hi: int) -> np.ndarray:
    """Bytes [lo, hi) of generate(kind, seed, nbytes) without producing the rest (english / mixed streams are
    independent 8 MiB chunks); lo must be a multiple of the chunk size.  Used to shard ONE corpus over ranks."""
    if kind == CODE or nbytes <= CHUNK:
        return generate(kind, seed, nbytes)[lo:hi]
    assert lo % CHUNK == 0
    out = np.empty(hi - lo, dtype=np.uint8)
    lib = _lib()
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        a = i * CHUNK
        n = min(CHUNK, nbytes - a)                     # the chunk as generate() sizes it
        if a + n <= hi:
            return lib.corpus_generate(kind, seed * 1000003 + i + 1, n, C.c_void_p(out.ctypes.data + (a - lo)))
        tmp = np.empty(n, dtype=np.uint8)              # last, partially wanted chunk
        rc = lib.corpus_generate(kind, seed * 1000003 + i + 1, n, tmp.ctypes.data_as(C.c_void_p))
        out[a - lo:] = tmp[:hi - a]
        return rc

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        rcs = list(ex.map(one, range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK)))
    if any(rcs):
        raise ValueError("bad corpus kind")
    return out


def _cut_points(text: np.ndarray, marks: np.ndarray, at_space: bool) -> np.ndarray:
    """Move each mark back to a legal cut: after a space (english) or onto a UTF-8 lead byte."""
    cuts = []
    n = len(text)
    for m in marks:
        m = int(min(m, n))
        if at_space:
            lo = max(0, m - 256)
            w = np.flatnonzero(text[lo:m] == 0x20)
            m = lo + int(w[-1]) + 1 if len(w) else m
        while 0 < m < n and (text[m] & 0xC0) == 0x80:
            m -= 1
        cuts.append(m)
    return np.asarray(cuts, dtype=np.uint64)


def docs_fixed(text: np.ndarray, doc_bytes: int, at_space: bool = True):
    n = len(text)
    marks = np.arange(doc_bytes, n, doc_bytes, dtype=np.int64)
    cuts = _cut_points(text, marks, at_space)
    off = np.unique(np.concatenate([[0], cuts, [n]]).astype(np.uint64))
    return text, off


def docs_from_lengths(text: np.ndarray, lengths: np.ndarray, at_space: bool):
    """Cut into docs of the requested byte lengths (legalised); keeps empty docs."""
    n = len(text)
    marks = np.minimum(np.cumsum(lengths.astype(np.int64)), n)
    if at_space or True:
        # vectorised legalisation: move back to the previous UTF-8 lead byte only
        m = marks.copy()
        for _ in range(3):
            bad = (m > 0) & (m < n) & ((text[np.minimum(m, n - 1)] & 0xC0) == 0x80)
            m = np.where(bad, m - 1, m)
        marks = np.maximum.accumulate(m)
    off = np.concatenate([[0], marks]).astype(np.uint64)
    if off[-1] != n:
        off = np.concatenate([off, [n]]).astype(np.uint64)
    return text, off


def config1(nbytes: int = 1 << 20, seed: int = 1001):
    """gpt2/r50k plumbing: ONE document of ASCII english-like text."""
    t = generate(ENGLISH, seed, nbytes)
    return t, np.asarray([0, nbytes], dtype=np.uint64)


def config2(nbytes: int = 1 << 30, seed: int = 1002, doc_bytes: int = 65536):
    """cl100k headline: english-like text as ~64 KiB documents cut at spaces."""
    return docs_fixed(generate(ENGLISH, seed, nbytes), doc_bytes, at_space=True)


def config3(nbytes: int = 1 << 30, seed: int = 1003):
    """o200k mixed UTF-8: documents log-uniform in 4 KiB .. 256 KiB."""
    t = generate(MIXED, seed, nbytes)
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.exp(rng.uniform(np.log(4096), np.log(262144), size=nbytes // 4096 + 8)).astype(np.int64)
    k = int(np.searchsorted(np.cumsum(lens), nbytes)) + 1
    return docs_from_lengths(t, lens[:k], at_space=False)


def config4(n_docs: int = 10_000_000, seed: int = 1004):
    """cl100k many short docs: lengths clip(round(LogNormal(ln 90, 0.5)), 0, 2000)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.clip(np.rint(rng.lognormal(np.log(90.0), 0.5, size=n_docs)), 0, 2000).astype(np.int64)
    lens[rng.integers(0, n_docs, size=max(1, n_docs // 1000))] = 0     # some empty docs
    total = int(lens.sum())
    t = generate(ENGLISH, seed, total)
    return docs_from_lengths(t, lens, at_space=False)


def config5(nbytes: int = 64 << 20, seed: int = 1005):
    """p50k: ONE code-like document with long-piece stressors."""
    t = generate(CODE, seed, nbytes)
    return t, np.asarray([0, nbytes], dtype=np.uint64)

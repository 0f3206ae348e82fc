#!/usr/bin/env python
"""Synthesise the stand-in vocabularies used by tests and bench.py.

The public vocabulary files (openaipublic.blob.core.windows.net/...) are not available offline
(SURVEY.md section 0 D5), so each encoding gets a synthetic mergeable_ranks table of matching size,
trained by tools/bpe_train.cpp on a sample of the same seeded generator the workload uses and
pre-tokenised with the REAL pat_str.  Output: tests/golden/vocab/<name>.tiktoken.gz in the
reference's own `base64(token) SP rank` line format (tiktoken/load.py:147-171).

Development tooling (uses the oracle's splitter); run once, outputs are committed.
"""
import base64, ctypes as C, gzip, os, subprocess, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import corpus                      # noqa: E402
from oracle import oracle as orc              # noqa: E402

SPECS = {
    # name: (pattern, corpus kind, seed, sample bytes, mergeable size)
    "r50k_like": (orc.R50K_PAT, corpus.ENGLISH, 7001, 48 << 20, 50256),
    "p50k_like": (orc.R50K_PAT, corpus.CODE, 7005, 48 << 20, 50280),
    "cl100k_like": (orc.CL100K_PAT, corpus.ENGLISH, 7002, 96 << 20, 100256),
    "o200k_like": (orc.O200K_PAT, corpus.MIXED, 7003, 128 << 20, 199998),
}


def train(name):
    pat, kind, seed, nbytes, size = SPECS[name]
    so = os.path.join(ROOT, "tools", "libbpetrain.so")
    src = os.path.join(ROOT, "tools", "bpe_train.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", so, src])
    T = C.CDLL(so)
    T.bpe_train.restype = C.c_int64
    t0 = time.time()
    text = corpus.generate(kind, seed, nbytes)
    if kind == corpus.ENGLISH and name != "r50k_like":
        pass
    L = orc._lib()
    o = orc.Oracle({bytes([i]): i for i in range(256)}, {}, pat)
    st = np.zeros(nbytes + 1, np.uint64); en = np.zeros(nbytes + 1, np.uint64)
    k = L.orc_split(o._h, text.ctypes.data_as(C.c_void_p), nbytes, st.ctypes.data_as(C.c_void_p),
                    en.ctypes.data_as(C.c_void_p), nbytes + 1)
    print(name, "pieces", k, "split s", round(time.time() - t0, 1), file=sys.stderr)
    out = np.zeros(size * 24 + 4096, np.uint8); off = np.zeros(size + 1, np.uint64)
    n = T.bpe_train(text.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), en.ctypes.data_as(C.c_void_p),
                    C.c_uint64(k), C.c_uint32(size), C.c_uint64(2), out.ctypes.data_as(C.c_void_p),
                    C.c_uint64(len(out)), off.ctypes.data_as(C.c_void_p))
    assert n > 0, n
    print(name, "vocab", n, "of", size, "train s", round(time.time() - t0, 1), file=sys.stderr)
    raw = out.tobytes()
    lines = []
    for r in range(n):
        lines.append(base64.b64encode(raw[int(off[r]):int(off[r + 1])]) + b" " + str(r).encode())
    os.makedirs(os.path.join(ROOT, "tests", "golden", "vocab"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "vocab", name + ".tiktoken.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(b"\n".join(lines) + b"\n")
    print(name, "->", path, os.path.getsize(path), file=sys.stderr)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(SPECS)):
        train(nm)

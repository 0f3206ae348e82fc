"""csrc/pack_ext.c `pack`: list[str] -> one UTF-8 blob + offsets, encoded by several threads straight from CPython's
1 / 2 / 4-byte string representations.  Must equal str.encode("utf-8") joined, whatever the mix of kinds and sizes, and
raise UnicodeEncodeError for lone surrogates like the `&str` extraction of src/py.rs:30 does."""
import random

import numpy as np
import pytest

from tiktoken_b200 import _b200pack


def _ref(docs):
    enc = [d.encode("utf-8") for d in docs]
    off = np.zeros(len(docs) + 1, np.uint64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    return b"".join(enc), off.tobytes()


def _rand_str(rnd, kind, n):
    if kind == 0:
        return "".join(chr(rnd.randrange(32, 127)) for _ in range(n))                       # ASCII (memcpy path)
    if kind == 1:
        return "".join(chr(rnd.randrange(32, 256)) for _ in range(n))                       # latin-1, one byte per code point
    if kind == 2:
        return "".join(chr(rnd.choice([rnd.randrange(32, 127), rnd.randrange(0x100, 0xD800), rnd.randrange(0xE000, 0x10000)]))
                       for _ in range(n))                                                   # UCS-2
    return "".join(chr(rnd.choice([rnd.randrange(32, 127), rnd.randrange(0x10000, 0x110000), rnd.randrange(0x800, 0xD800)]))
                   for _ in range(n))                                                       # UCS-4


def test_pack_equals_str_encode_on_every_string_kind():
    rnd = random.Random(11)
    for _ in range(150):
        docs = [_rand_str(rnd, rnd.randrange(4), rnd.choice([0, 1, 2, 5, 100, 5000])) for _ in range(rnd.choice([0, 1, 3, 50]))]
        assert _b200pack.pack(docs) == _ref(docs)
    assert _b200pack.pack([]) == _ref([]) and _b200pack.pack(["", "", ""]) == _ref(["", "", ""])
    assert _b200pack.pack(("a", "b")) == _ref(["a", "b"])                                   # any sequence


def test_pack_with_several_threads_and_uneven_documents():
    rnd = random.Random(12)
    docs = [_rand_str(rnd, rnd.randrange(4), n) for n in (3_000_000, 0, 7, 400_000, 400_000, 1, 2_000_000, 50_000)]   # > 1 Mi code points: threads
    rnd.shuffle(docs)
    assert _b200pack.pack(docs) == _ref(docs)


def test_pack_errors():
    for bad in (["ok", "a\ud800b"], ["x" * 3_000_000, "\udfff"], ["\ud83d"]):
        with pytest.raises(UnicodeEncodeError):
            _b200pack.pack(bad)
    with pytest.raises(TypeError):
        _b200pack.pack(["a", 5])
    with pytest.raises(TypeError):
        _b200pack.pack(["a", b"bytes"])


def test_unpack_shares_the_int_objects_of_the_cache():
    """`unpack`: token arrays -> list[list[int]]; ids inside the cache come out as the cache's own int objects (no allocation
    per token), ids beyond it as fresh ints; the lists equal the array slices either way."""
    rng = np.random.default_rng(5)
    tok = rng.integers(0, 5000, size=20_000, dtype=np.uint32)
    tok[::97] = rng.integers(1 << 20, 1 << 31, size=len(tok[::97]), dtype=np.uint32)        # ids beyond the cache
    off = np.sort(np.concatenate([[0, len(tok), 7, 7], rng.integers(0, len(tok), size=40)])).astype(np.uint64)   # incl. an empty doc
    cache = list(range(5000))
    for c in (None, cache):
        args = (tok.ctypes.data, off.ctypes.data, len(off) - 1) + ((c,) if c is not None else ())
        out = _b200pack.unpack(*args)
        assert out == [tok[int(off[i]):int(off[i + 1])].tolist() for i in range(len(off) - 1)]
    first = next(d for d in out if d)
    assert all(x is cache[x] for x in first if x < 5000)
    with pytest.raises(TypeError):
        _b200pack.unpack(tok.ctypes.data, off.ctypes.data, 2, (1, 2, 3))

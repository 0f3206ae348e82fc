"""The C-ABI library loads without a GPU and exports every symbol include/b200bpe.h declares;
argument / pattern validation works without touching the device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import vocab_util as vu
from conftest import ROOT, have_gpu
from tiktoken_b200 import _lib


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200bpe.h")).read()
    return sorted(set(re.findall(r"\b(b200bpe_[a-z_0-9]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    so = _lib.build()
    L = C.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/b200bpe.h but not exported"
    assert sorted(_lib.EXPORTS) == syms
    assert b"sm_100a" in _lib.lib().b200bpe_version()


def test_unsupported_pattern_is_value_error():
    from tiktoken_b200 import _tiktoken
    with pytest.raises(ValueError):
        _tiktoken.CoreBPE({bytes([i]): i for i in range(256)}, {}, r"\w+|\s+")   # no CPU regex fallback


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_not_silently():
    from tiktoken_b200 import _tiktoken
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _tiktoken.CoreBPE({bytes([i]): i for i in range(256)}, {}, vu.CL100K_PAT)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "tiktoken_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_pack_helper_matches_python_marshalling():
    """csrc/pack_ext.c (list[str] -> packed UTF-8 + offsets, token buffers -> list[list[int]])
    agrees with the pure-Python marshalling, including the UnicodeEncodeError on lone surrogates
    that tiktoken/core.py:77,128 relies on."""
    import numpy as np
    import __graft_entry__  # noqa: F401  (sys.path)
    from tiktoken_b200 import _tiktoken as T
    if T._b200pack is None:
        import __graft_entry__ as g
        g.build()
        import importlib
        importlib.reload(T)
    assert T._b200pack is not None
    docs = ["", "hello", "héllo 日本", "", "\U0001F600 x" * 50, "a" * 100000]
    arr, off = T.CoreBPE._pack(docs)
    enc = [d.encode() for d in docs]
    assert arr.tobytes() == b"".join(enc)
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(e) for e in enc])]).tolist()
    arr0, off0 = T.CoreBPE._pack([])
    assert off0.tolist() == [0]
    with pytest.raises(UnicodeEncodeError):
        T.CoreBPE._pack(["ok", "\ud83d"])
    with pytest.raises(TypeError):
        T._b200pack.pack([b"bytes"])
    toks = (np.arange(10, dtype=np.uint64) * 400000003 % (1 << 32)).astype(np.uint32)
    toff = np.asarray([0, 0, 3, 3, 10], np.uint64)
    got = T._b200pack.unpack(toks.ctypes.data, toff.ctypes.data, 4)
    assert got == [[], toks[:3].tolist(), [], toks[3:].tolist()]

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

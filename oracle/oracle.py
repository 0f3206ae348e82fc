"""ctypes wrapper over liboracle.so (oracle/bpe_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# pat_strs of tiktoken_ext/openai_public.py:12-14, :89, :104-114 (reference v0.14.0)
R50K_PAT = r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}++| ?\p{N}++| ?[^\s\p{L}\p{N}]++|\s++$|\s+(?!\S)|\s"""
CL100K_PAT = r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}++|\p{N}{1,3}+| ?[^\s\p{L}\p{N}]++[\r\n]*+|\s++$|\s*[\r\n]|\s+(?!\S)|\s"""
O200K_PAT = "|".join([
    r"""[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?""",
    r"""[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?""",
    r"""\p{N}{1,3}""",
    r""" ?[^\s\p{L}\p{N}]+[\r\n/]*""",
    r"""\s*[\r\n]+""",
    r"""\s+(?!\S)""",
    r"""\s+""",
])
PAT_R50K, PAT_CL100K, PAT_O200K = 0, 1, 2


def pattern_id(pat_str: str) -> int:
    try:
        return {R50K_PAT: PAT_R50K, CL100K_PAT: PAT_CL100K, O200K_PAT: PAT_O200K}[pat_str]
    except KeyError:
        raise ValueError("oracle only restates the three pat_strs of openai_public.py") from None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bpe_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, u32p, u64p = C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [u8p, u64p, u32p, C.c_uint32, u8p, u64p, u32p, C.c_uint32, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_split.restype = C.c_int64
        L.orc_split.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p, u64p, C.c_uint64]
        L.orc_encode_ordinary.restype = C.c_int64
        L.orc_encode_ordinary.argtypes = [C.c_void_p, u8p, C.c_uint64, u32p]
        L.orc_encode_piece.restype = C.c_int64
        L.orc_encode_piece.argtypes = [C.c_void_p, u8p, C.c_uint64, u32p, C.c_int]
        L.orc_byte_pair_split.restype = C.c_int64
        L.orc_byte_pair_split.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p]
        L.orc_encode.restype = C.c_int64
        L.orc_encode.argtypes = [C.c_void_p, u8p, C.c_uint64, u8p, u32p]
        L.orc_encode_ordinary_batch.restype = C.c_int64
        L.orc_encode_ordinary_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_int, u32p, u64p]
        _LIB = L
    return _LIB


def _flatten(items):
    blob = b"".join(items)
    off = np.zeros(len(items) + 1, dtype=np.uint64)
    if items:
        off[1:] = np.cumsum([len(b) for b in items], dtype=np.uint64)
    return np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(0, np.uint8), off


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Same constructor arguments as _tiktoken.CoreBPE (src/py.rs:16-23)."""

    def __init__(self, mergeable_ranks: dict[bytes, int], special_tokens: dict[str, int], pat_str: str):
        self._L = _lib()
        toks = list(mergeable_ranks.keys())
        blob, off = _flatten(toks)
        rank = np.asarray([mergeable_ranks[t] for t in toks], dtype=np.uint32)
        self._special_names = list(special_tokens.keys())
        sblob, soff = _flatten([s.encode("utf-8") for s in self._special_names])
        srank = np.asarray([special_tokens[s] for s in self._special_names], dtype=np.uint32)
        self._h = self._L.orc_new(_ptr(blob), _ptr(off), _ptr(rank), len(toks),
                                  _ptr(sblob), _ptr(soff), _ptr(srank), len(self._special_names),
                                  pattern_id(pat_str))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_free(self._h)
            self._h = None

    @staticmethod
    def _bytes(text) -> np.ndarray:
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        return np.frombuffer(b, dtype=np.uint8) if b else np.zeros(0, np.uint8)

    def split(self, text) -> list[bytes]:
        a = self._bytes(text)
        n = len(a)
        st = np.zeros(n + 1, np.uint64); en = np.zeros(n + 1, np.uint64)
        k = self._L.orc_split(self._h, _ptr(a), n, _ptr(st), _ptr(en), n + 1)
        raw = a.tobytes()
        return [raw[int(st[i]):int(en[i])] for i in range(k)]

    def encode_ordinary_np(self, text) -> np.ndarray:
        a = self._bytes(text)
        out = np.zeros(len(a) + 1, np.uint32)
        k = self._L.orc_encode_ordinary(self._h, _ptr(a), len(a), _ptr(out))
        return out[:k].copy()

    def encode_ordinary(self, text) -> list[int]:
        return self.encode_ordinary_np(text).tolist()

    def encode(self, text, allowed_special=frozenset()) -> list[int]:
        a = self._bytes(text)
        out = np.zeros(len(a) + 1, np.uint32)
        allowed = np.asarray([1 if s in allowed_special else 0 for s in self._special_names] + [0], np.uint8)
        k = self._L.orc_encode(self._h, _ptr(a), len(a), _ptr(allowed), _ptr(out))
        return out[:k].tolist()

    def encode_single_piece(self, piece: bytes, force: int = 0) -> list[int]:
        a = self._bytes(piece)
        out = np.zeros(len(a) + 1, np.uint32)
        k = self._L.orc_encode_piece(self._h, _ptr(a), len(a), _ptr(out), force)
        return out[:k].tolist()

    def byte_pair_split(self, piece: bytes) -> list[bytes]:
        a = self._bytes(piece)
        b = np.zeros(len(a) + 2, np.uint64)
        k = self._L.orc_byte_pair_split(self._h, _ptr(a), len(a), _ptr(b))
        return [bytes(piece[int(b[i]):int(b[i + 1])]) for i in range(k - 1)]

    def encode_ordinary_batch_np(self, text: np.ndarray, doc_off: np.ndarray, n_threads: int = 1):
        """text: uint8[N], doc_off: uint64[n_docs+1] -> (tokens uint32[T], tok_off uint64[n_docs+1])"""
        text = np.ascontiguousarray(text, np.uint8); doc_off = np.ascontiguousarray(doc_off, np.uint64)
        n_docs = len(doc_off) - 1
        out = np.zeros(int(doc_off[-1]) + 1, np.uint32)
        toff = np.zeros(n_docs + 1, np.uint64)
        k = self._L.orc_encode_ordinary_batch(self._h, _ptr(text), _ptr(doc_off), n_docs, n_threads,
                                              _ptr(out), _ptr(toff))
        return out[:k], toff

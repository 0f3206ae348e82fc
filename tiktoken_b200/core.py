"""Host-side `Encoding` with the public surface of `tiktoken.Encoding` (reference:
tiktoken/core.py) on top of the B200 engine.

What differs from the reference host class, by design:
  * `self._core_bpe` is `tiktoken_b200._tiktoken.CoreBPE` (ctypes -> libb200bpe.so -> sm_100a
    kernels) instead of the Rust extension;
  * `encode_ordinary_batch` / `encode_batch` make ONE native call for the whole batch instead of
    a ThreadPoolExecutor over per-document calls (core.py:164-206); `num_threads` is accepted and
    ignored;
  * array-returning batch variants (`*_to_numpy`) avoid building Python lists at all.
Special-token policy, surrogate fix-up, decode helpers and pickling behave as in the reference.
"""
from __future__ import annotations

import functools
import re as _re
from typing import AbstractSet, Collection, Literal, Sequence

import numpy as np

from . import _tiktoken


def _special_pattern(tokens: frozenset[str]):
    return _cached_special_pattern(tokens)


@functools.lru_cache(maxsize=128)
def _cached_special_pattern(tokens: frozenset[str]):
    return _re.compile("(" + "|".join(_re.escape(t) for t in tokens) + ")")


def _raise_disallowed(token: str):
    raise ValueError(
        f"Encountered text corresponding to disallowed special token {token!r}.\n"
        "If you want this text to be encoded as a special token, "
        f"pass it to `allowed_special`, e.g. `allowed_special={{{token!r}, ...}}`.\n"
        "If you want this text to be encoded as normal text, disable the check for this token "
        f"by passing `disallowed_special=(enc.special_tokens_set - {{{token!r}}})`.\n"
        "To disable this check for all special tokens, pass `disallowed_special=()`.\n"
    )


def _fix_surrogates(text: str) -> str:
    # same fix-up as the reference (core.py:77-80): lone surrogates become U+FFFD
    return text.encode("utf-16", "surrogatepass").decode("utf-16", "replace")


class Encoding:
    def __init__(self, name: str, *, pat_str: str, mergeable_ranks: dict[bytes, int],
                 special_tokens: dict[str, int], explicit_n_vocab: int | None = None, device: int | None = None):
        self.name = name
        self._pat_str = pat_str
        self._mergeable_ranks = mergeable_ranks
        self._special_tokens = special_tokens
        self.max_token_value = max(max(mergeable_ranks.values()), max(special_tokens.values(), default=0))
        if explicit_n_vocab:
            assert len(mergeable_ranks) + len(special_tokens) == explicit_n_vocab
            assert self.max_token_value == explicit_n_vocab - 1
        self._special_token_values = set(special_tokens.values())
        self._core_bpe = _tiktoken.CoreBPE(mergeable_ranks, special_tokens, pat_str, device=device)

    @classmethod
    def from_tiktoken_file(cls, name: str, path_or_bytes, *, pat_str: str, special_tokens: dict[str, int],
                           explicit_n_vocab: int | None = None, device: int | None = None) -> "Encoding":
        """An Encoding straight from a `.tiktoken` vocabulary file (`base64(token) rank` per line, the format
        `tiktoken/load.py:159-171` reads; `.gz` accepted), parsed in C into the flattened arrays the engine
        takes -- the 100-200 k-entry Python dict is only built if something asks for it (pickling by value,
        `encode_single_token`, `token_byte_values`)."""
        from ._tiktoken import _b200pack
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            data = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as f:
                data = f.read()
        if data[:2] == b"\x1f\x8b":
            import gzip
            data = gzip.decompress(data)
        if _b200pack is None:                          # no C helper: the reference's own parse (load.py:159-171)
            import base64
            ranks = {base64.b64decode(tok): int(r) for tok, r in (ln.split() for ln in data.splitlines() if ln)}
            return cls(name, pat_str=pat_str, mergeable_ranks=ranks, special_tokens=special_tokens,
                       explicit_n_vocab=explicit_n_vocab, device=device)
        blob, off, rk = _b200pack.parse_tiktoken(data)
        blob, off, rk = np.frombuffer(blob, np.uint8), np.frombuffer(off, np.uint64), np.frombuffer(rk, np.uint32)
        self = cls.__new__(cls)
        self.name = name
        self._pat_str = pat_str
        self._ranks_dict = None
        self._special_tokens = special_tokens
        self.max_token_value = max(int(rk.max()) if len(rk) else 0, max(special_tokens.values(), default=0))
        if explicit_n_vocab:
            assert len(rk) + len(special_tokens) == explicit_n_vocab
            assert self.max_token_value == explicit_n_vocab - 1
        self._special_token_values = set(special_tokens.values())
        self._core_bpe = _tiktoken.CoreBPE.from_flat(blob, off, rk, special_tokens, pat_str, device=device)
        return self

    @property
    def _mergeable_ranks(self) -> dict[bytes, int]:
        if self._ranks_dict is None:
            self._ranks_dict = self._core_bpe._encoder
        return self._ranks_dict

    @_mergeable_ranks.setter
    def _mergeable_ranks(self, value: dict[bytes, int]) -> None:
        self._ranks_dict = value

    def __repr__(self) -> str:
        return f"<Encoding {self.name!r}>"

    # ---------------------------------------------------------------- special-token policy
    def _policy(self, allowed_special, disallowed_special):
        if allowed_special == "all":
            allowed_special = self.special_tokens_set
        if disallowed_special == "all":
            disallowed_special = self.special_tokens_set - allowed_special
        if disallowed_special and not isinstance(disallowed_special, frozenset):
            disallowed_special = frozenset(disallowed_special)
        return allowed_special, disallowed_special

    @staticmethod
    def _check_disallowed(text: str, disallowed_special) -> None:
        if disallowed_special:
            m = _special_pattern(disallowed_special).search(text)
            if m:
                _raise_disallowed(m.group())

    def _check_disallowed_packed(self, texts, blob: np.ndarray, off: np.ndarray, disallowed_special) -> None:
        """The same check (core.py:120-124) for a batch, on the packed UTF-8: one C pass (memchr + memcmp,
        csrc/pack_ext.c) instead of a Python regex search per document; a match never straddles documents."""
        if not disallowed_special:
            return
        from ._tiktoken import _b200pack
        if _b200pack is None:
            for t in texts:
                self._check_disallowed(t, disallowed_special)
            return
        names = sorted(disallowed_special)
        hit = _b200pack.find_first(blob[:int(off[-1])], off, [n.encode("utf-8") for n in names])
        if hit is not None:
            _raise_disallowed(names[hit[1]])

    # ---------------------------------------------------------------- encoding
    def encode_ordinary(self, text: str) -> list[int]:
        try:
            return self._core_bpe.encode_ordinary(text)
        except UnicodeEncodeError:
            return self._core_bpe.encode_ordinary(_fix_surrogates(text))

    def encode(self, text: str, *, allowed_special: Literal["all"] | AbstractSet[str] = set(),  # noqa: B006
               disallowed_special: Literal["all"] | Collection[str] = "all") -> list[int]:
        allowed_special, disallowed_special = self._policy(allowed_special, disallowed_special)
        self._check_disallowed(text, disallowed_special)
        try:
            return self._core_bpe.encode(text, allowed_special)
        except UnicodeEncodeError:
            return self._core_bpe.encode(_fix_surrogates(text), allowed_special)

    def encode_to_numpy(self, text: str, *, allowed_special: Literal["all"] | AbstractSet[str] = set(),  # noqa: B006
                        disallowed_special: Literal["all"] | Collection[str] = "all") -> np.ndarray:
        allowed_special, disallowed_special = self._policy(allowed_special, disallowed_special)
        self._check_disallowed(text, disallowed_special)
        buffer = self._core_bpe.encode_to_tiktoken_buffer(text, allowed_special)
        return np.frombuffer(buffer, dtype=np.uint32)

    def _pack(self, texts: Sequence[str]):
        try:
            return self._core_bpe._pack(list(texts))
        except UnicodeEncodeError:
            return self._core_bpe._pack([_fix_surrogates(t) for t in texts])

    def encode_ordinary_batch(self, text: list[str], *, num_threads: int = 8) -> list[list[int]]:
        """One native call for the whole batch (num_threads is ignored: the GPU is the pool)."""
        t, off = self._pack(text)
        return self._core_bpe._unpack(self._core_bpe.encode_ordinary_batch_buffer(t, off))

    def encode_batch(self, text: list[str], *, num_threads: int = 8,
                     allowed_special: Literal["all"] | AbstractSet[str] = set(),  # noqa: B006
                     disallowed_special: Literal["all"] | Collection[str] = "all") -> list[list[int]]:
        allowed_special, disallowed_special = self._policy(allowed_special, disallowed_special)
        t, off = self._pack(text)
        self._check_disallowed_packed(text, t, off, disallowed_special)
        return self._core_bpe._unpack(self._core_bpe.encode_batch_buffer(t, off, allowed_special))

    def encode_ordinary_batch_to_numpy(self, text: list[str]):
        """-> (tokens uint32[T], offsets uint64[n_docs+1]); document d is tokens[offsets[d]:offsets[d+1]]."""
        t, off = self._pack(text)
        buf = self._core_bpe.encode_ordinary_batch_buffer(t, off)
        out = (np.array(buf.tokens()), np.array(buf.offsets()))
        buf.close()
        return out

    def encode_ordinary_packed(self, text_bytes: np.ndarray, doc_off: np.ndarray):
        """Already-packed input: uint8[N] UTF-8 + uint64[n_docs+1] -> zero-copy TokenBuffer."""
        return self._core_bpe.encode_ordinary_batch_buffer(np.ascontiguousarray(text_bytes, np.uint8),
                                                           np.ascontiguousarray(doc_off, np.uint64))

    def encode_with_unstable(self, text: str, *, allowed_special=set(), disallowed_special="all"):  # noqa: B006
        allowed_special, disallowed_special = self._policy(allowed_special, disallowed_special)
        self._check_disallowed(text, disallowed_special)
        return self._core_bpe.encode_with_unstable(text, allowed_special)

    def encode_single_token(self, text_or_bytes: str | bytes) -> int:
        if isinstance(text_or_bytes, str):
            text_or_bytes = text_or_bytes.encode("utf-8")
        return self._core_bpe.encode_single_token(text_or_bytes)

    # ---------------------------------------------------------------- decoding
    def decode_bytes(self, tokens: Sequence[int]) -> bytes:
        return self._core_bpe.decode_bytes(tokens)

    def decode(self, tokens: Sequence[int], errors: str = "replace") -> str:
        return self._core_bpe.decode_bytes(tokens).decode("utf-8", errors=errors)

    def decode_single_token_bytes(self, token: int) -> bytes:
        return self._core_bpe.decode_single_token_bytes(token)

    def decode_tokens_bytes(self, tokens: Sequence[int]) -> list[bytes]:
        return [self.decode_single_token_bytes(t) for t in tokens]

    def decode_with_offsets(self, tokens: Sequence[int]) -> tuple[str, list[int]]:
        pieces = self.decode_tokens_bytes(tokens)
        n_chars, offsets = 0, []
        for piece in pieces:
            starts_mid_char = 0x80 <= piece[0] < 0xC0
            offsets.append(max(0, n_chars - (1 if starts_mid_char else 0)))
            n_chars += sum(1 for b in piece if not 0x80 <= b < 0xC0)
        return b"".join(pieces).decode("utf-8", errors="strict"), offsets

    def decode_batch(self, batch: Sequence[Sequence[int]], *, errors: str = "replace", num_threads: int = 8) -> list[str]:
        """One native call: device gather of the token byte strings (reference: thread pool, core.py:337-343)."""
        return [b.decode("utf-8", errors=errors) for b in self._core_bpe.decode_bytes_batch(batch)]

    def decode_bytes_batch(self, batch: Sequence[Sequence[int]], *, num_threads: int = 8) -> list[bytes]:
        return self._core_bpe.decode_bytes_batch(batch)

    def decode_packed(self, tokens: np.ndarray, tok_off: np.ndarray):
        """Array form: tokens uint32[T] + offsets uint64[n_docs+1] -> (bytes uint8[B], byte offsets uint64[n_docs+1])."""
        return self._core_bpe.decode_batch_buffer(tokens, tok_off)

    # ---------------------------------------------------------------- misc
    def token_byte_values(self) -> list[bytes]:
        return self._core_bpe.token_byte_values()

    @property
    def eot_token(self) -> int:
        return self._special_tokens["<|endoftext|>"]

    @functools.cached_property
    def special_tokens_set(self) -> set[str]:
        return set(self._special_tokens.keys())

    def is_special_token(self, token: int) -> bool:
        assert isinstance(token, int)
        return token in self._special_token_values

    @property
    def n_vocab(self) -> int:
        return self.max_token_value + 1

    # ---------------------------------------------------------------- private helpers kept for parity
    def _encode_single_piece(self, text_or_bytes: str | bytes) -> list[int]:
        if isinstance(text_or_bytes, str):
            text_or_bytes = text_or_bytes.encode("utf-8")
        return self._core_bpe.encode_single_piece(text_or_bytes)

    def _encode_only_native_bpe(self, text: str) -> list[int]:
        import regex
        out: list[int] = []
        for piece in regex.findall(regex.compile(self._pat_str), text):
            out.extend(self._core_bpe.encode_single_piece(piece.encode("utf-8")))
        return out

    def _encode_bytes(self, text: bytes) -> list[int]:
        return self._core_bpe._encode_bytes(text)

    def __getstate__(self) -> object:
        from . import registry
        if self is registry.ENCODINGS.get(self.name):
            return self.name                      # registered encodings pickle by reference
        return {"name": self.name, "pat_str": self._pat_str, "mergeable_ranks": self._mergeable_ranks,
                "special_tokens": self._special_tokens}

    def __setstate__(self, value: object) -> None:
        from . import registry
        if isinstance(value, str):
            self.__dict__ = registry.get_encoding(value).__dict__
            return
        self.__init__(**value)                    # rebuilds the device tables

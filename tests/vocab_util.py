"""Vocabulary sources for tests and bench.py.

Real tiktoken vocabularies are used when TIKTOKEN_CACHE_DIR holds them (file names =
sha1(url), tiktoken/load.py:51-53); otherwise the committed synthetic stand-ins from
tests/golden/vocab/ (see tools/make_vocab.py) are used.  `source` in the returned tuple says which.
"""
from __future__ import annotations

import base64
import gzip
import hashlib
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

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

ENDOFTEXT, FIM_PREFIX, FIM_MIDDLE, FIM_SUFFIX, ENDOFPROMPT = (
    "<|endoftext|>", "<|fim_prefix|>", "<|fim_middle|>", "<|fim_suffix|>", "<|endofprompt|>")

# name -> (pat_str, synthetic file, real url, special tokens)   [tiktoken_ext/openai_public.py]
ENCODINGS = {
    "r50k_base": (R50K_PAT, "r50k_like", "https://openaipublic.blob.core.windows.net/encodings/r50k_base.tiktoken",
                  {ENDOFTEXT: 50256}),
    "p50k_base": (R50K_PAT, "p50k_like", "https://openaipublic.blob.core.windows.net/encodings/p50k_base.tiktoken",
                  {ENDOFTEXT: 50256}),
    "cl100k_base": (CL100K_PAT, "cl100k_like", "https://openaipublic.blob.core.windows.net/encodings/cl100k_base.tiktoken",
                    {ENDOFTEXT: 100257, FIM_PREFIX: 100258, FIM_MIDDLE: 100259, FIM_SUFFIX: 100260, ENDOFPROMPT: 100276}),
    "o200k_base": (O200K_PAT, "o200k_like", "https://openaipublic.blob.core.windows.net/encodings/o200k_base.tiktoken",
                   {ENDOFTEXT: 199999, ENDOFPROMPT: 200018}),
}


def parse_tiktoken_bpe(data: bytes) -> dict[bytes, int]:
    out = {}
    for line in data.splitlines():
        if line:
            tok, rank = line.split()
            out[base64.b64decode(tok)] = int(rank)
    return out


def real_vocab_path(url: str) -> str | None:
    d = os.environ.get("TIKTOKEN_CACHE_DIR") or os.environ.get("DATA_GYM_CACHE_DIR")
    if not d:
        return None
    p = os.path.join(d, hashlib.sha1(url.encode()).hexdigest())
    return p if os.path.exists(p) else None


def load_encoding(name: str, allow_real: bool = True):
    """-> (pat_str, mergeable_ranks, special_tokens, source) with source in {'real','synthetic'}."""
    pat, syn, url, special = ENCODINGS[name]
    p = real_vocab_path(url) if allow_real else None
    if p:
        with open(p, "rb") as f:
            return pat, parse_tiktoken_bpe(f.read()), dict(special), "real"
    with gzip.open(os.path.join(_HERE, "golden", "vocab", syn + ".tiktoken.gz"), "rb") as f:
        ranks = parse_tiktoken_bpe(f.read())
    if name == "p50k_base":
        special = {ENDOFTEXT: 50280}      # synthetic table fills ranks 0..50279; keep ids disjoint
    return pat, ranks, dict(special), "synthetic"

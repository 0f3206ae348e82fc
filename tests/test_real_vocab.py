"""The reference's real-vocabulary golden vectors (tests/test_encoding.py:16-78,
tests/test_simple_public.py:10).  They need the public vocabulary files, which are not available
offline; the tests run when TIKTOKEN_CACHE_DIR holds them (SURVEY.md 8(c) lists the file names)."""
import pytest

import vocab_util as vu
from conftest import have_gpu

pytestmark = pytest.mark.gpu


def _enc(name):
    pat, ranks, special, src = vu.load_encoding(name)
    if src != "real":
        pytest.skip(f"real {name} vocabulary not in TIKTOKEN_CACHE_DIR (parity on real vocab unpinned offline)")
    import tiktoken_b200
    return tiktoken_b200.Encoding(name, pat_str=pat, mergeable_ranks=ranks, special_tokens=special)


def test_r50k_vectors():
    e = _enc("r50k_base")
    assert e.encode("hello world") == [31373, 995]
    assert e.encode("hello <|endoftext|>", allowed_special="all") == [31373, 220, 50256]
    assert e.encode("0") == [15] and e.encode("00") == [405] and e.encode("000") == [830]
    assert e.encode("0" * 17) == [8269, 10535, 830]


def test_cl100k_vectors():
    e = _enc("cl100k_base")
    assert e.encode("hello world") == [15339, 1917]
    assert e.encode("hello <|endoftext|>", allowed_special="all") == [15339, 220, 100257]
    assert e.encode("rer") == [38149] and e.encode("'rer") == [2351, 81]
    assert e.encode("today\n ") == [31213, 198, 220]
    assert e.encode("today\n \n") == [31213, 27907] and e.encode("today\n  \n") == [31213, 14211]
    assert e.encode(" \x850") == [220, 126, 227, 15]
    assert e.encode("\U0001F44D") == [9468, 239, 235]

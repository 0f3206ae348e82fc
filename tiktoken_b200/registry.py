"""Encoding registry with the behaviour of tiktoken/registry.py: encodings are looked up by
name among constructors published by plugin modules of the `tiktoken_ext` namespace package
(each exposes ENCODING_CONSTRUCTORS: name -> () -> kwargs for Encoding), built once and cached.
The plugins (e.g. tiktoken_ext.openai_public) are the reference's own, untouched; only the
class they feed is the B200-backed `tiktoken_b200.Encoding`.
"""
from __future__ import annotations

import importlib
import pkgutil
import threading
from typing import Callable

from .core import Encoding

ENCODINGS: dict[str, Encoding] = {}
ENCODING_CONSTRUCTORS: dict[str, Callable[[], dict]] | None = None
_lock = threading.RLock()
_local: dict[str, Callable[[], dict]] = {}


def register_encoding(name: str, constructor: Callable[[], dict]) -> None:
    """Register a local constructor (same contract as a tiktoken_ext plugin entry)."""
    with _lock:
        _local[name] = constructor
        if ENCODING_CONSTRUCTORS is not None:
            ENCODING_CONSTRUCTORS[name] = constructor


def _discover() -> None:
    global ENCODING_CONSTRUCTORS
    with _lock:
        if ENCODING_CONSTRUCTORS is not None:
            return
        found: dict[str, Callable[[], dict]] = {}
        try:
            import tiktoken_ext
            for info in pkgutil.iter_modules(tiktoken_ext.__path__, tiktoken_ext.__name__ + "."):
                mod = importlib.import_module(info.name)
                ctors = getattr(mod, "ENCODING_CONSTRUCTORS", None)
                if ctors is None:
                    raise ValueError(f"tiktoken plugin {info.name} does not define ENCODING_CONSTRUCTORS")
                for name, ctor in ctors.items():
                    if name in found:
                        raise ValueError(f"Duplicate encoding name {name} in tiktoken plugin {info.name}")
                    found[name] = ctor
        except ImportError:
            pass
        found.update(_local)
        ENCODING_CONSTRUCTORS = found


def get_encoding(encoding_name: str) -> Encoding:
    if not isinstance(encoding_name, str):
        raise ValueError(f"Expected a string in get_encoding, got {type(encoding_name)}")
    if encoding_name in ENCODINGS:
        return ENCODINGS[encoding_name]
    with _lock:
        if encoding_name in ENCODINGS:
            return ENCODINGS[encoding_name]
        _discover()
        assert ENCODING_CONSTRUCTORS is not None
        if encoding_name not in ENCODING_CONSTRUCTORS:
            raise ValueError(f"Unknown encoding {encoding_name}.\nPlugins found: {sorted(ENCODING_CONSTRUCTORS)}")
        enc = Encoding(**ENCODING_CONSTRUCTORS[encoding_name]())
        ENCODINGS[encoding_name] = enc
        return enc


def list_encoding_names() -> list[str]:
    with _lock:
        _discover()
        assert ENCODING_CONSTRUCTORS is not None
        return list(ENCODING_CONSTRUCTORS)

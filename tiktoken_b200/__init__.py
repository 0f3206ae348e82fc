"""tiktoken_b200 -- B200-native BPE encoder behind tiktoken's API (encode hot path only).

    from tiktoken_b200 import Encoding          # same constructor as tiktoken.Encoding
    enc = Encoding("my_enc", pat_str=..., mergeable_ranks=..., special_tokens=...)
    enc.encode_ordinary_batch(docs)             # one native call -> sm_100a kernels

`tiktoken_b200._tiktoken.CoreBPE` is the drop-in for the Rust extension.
"""
from .core import Encoding  # noqa: F401
from .registry import get_encoding, list_encoding_names, register_encoding  # noqa: F401

__version__ = "0.1.0"

"""Document sharding across the GPUs of one box.

Documents are independent haystacks, so the encode path shards embarrassingly: rank r encodes a
contiguous range of documents chosen to balance BYTES, with its own replica of the rank tables.
The only exchange is one all-gather of (n_tokens, n_docs) per rank (NCCL over NVLink when the
process group is NCCL), from which every rank derives the global token offset of its shard; no
token payload crosses NVLink.
"""
from __future__ import annotations

import numpy as np


def shard_ranges(doc_off: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous document ranges [lo, hi) per rank with ~equal bytes.  A single huge document
    cannot be split (one haystack), so ranks may receive empty ranges."""
    doc_off = np.asarray(doc_off, dtype=np.uint64)
    n_docs = len(doc_off) - 1
    total = int(doc_off[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        # first document whose START is >= target keeps prefix sums balanced
        d = int(np.searchsorted(doc_off[:-1], target, side="left"))
        cuts.append(min(max(d, cuts[-1]), n_docs))
    cuts.append(n_docs)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def local_view(text: np.ndarray, doc_off: np.ndarray, lo: int, hi: int):
    """Slice the packed batch to documents [lo, hi) and rebase the offsets."""
    doc_off = np.asarray(doc_off, dtype=np.uint64)
    b0, b1 = int(doc_off[lo]), int(doc_off[hi])
    return text[b0:b1], (doc_off[lo:hi + 1] - doc_off[lo]).astype(np.uint64)


def gather_counts(n_tokens: int, n_docs: int, rank: int, world: int, device=None):
    """All-gather (n_tokens, n_docs) of every rank -> (counts[world,2], token_base, doc_base)."""
    import torch
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        counts = np.asarray([[n_tokens, n_docs]], dtype=np.int64)
        return counts, 0, 0
    mine = torch.tensor([n_tokens, n_docs], dtype=torch.int64, device=device or "cpu")
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    counts = torch.stack(allc).cpu().numpy()
    return counts, int(counts[:rank, 0].sum()), int(counts[:rank, 1].sum())


def encode_sharded(encode_packed, text: np.ndarray, doc_off: np.ndarray, rank: int, world: int, device=None):
    """Encode this rank's shard with `encode_packed(text, doc_off) -> (tokens, tok_off)` and locate it
    in the global output.  Returns a dict with the local arrays and the global placement."""
    lo, hi = shard_ranges(doc_off, world)[rank]
    ltext, loff = local_view(text, doc_off, lo, hi)
    tokens, tok_off = encode_packed(ltext, loff)
    counts, token_base, doc_base = gather_counts(int(len(tokens)), hi - lo, rank, world, device)
    return {"doc_range": (lo, hi), "tokens": tokens, "tok_off": tok_off, "token_base": token_base,
            "doc_base": doc_base, "counts": counts, "total_tokens": int(counts[:, 0].sum())}

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


class CountExchange:
    """The same all-gather as `gather_counts`, posted asynchronously on preallocated buffers so that a
    stream of batches does not stop for it: `post()` after batch k, `wait()` (any time later) returns
    (counts[world,2], token_base, doc_base) of that batch.  With an NCCL group the exchange runs on the
    communicator's stream next to the kernels of batch k+1; the placement is only needed when the shard
    is written out."""

    def __init__(self, rank: int, world: int, device=None):
        import torch
        import torch.distributed as dist
        self.rank, self.world = rank, world
        self.active = world > 1 and dist.is_initialized()
        self._pending = []
        if self.active:
            self._dev = device or "cpu"
            self._pool = []                                  # reusable (host, mine, gathered) buffer sets

    def _buffers(self):
        import torch
        if self._pool:
            return self._pool.pop()
        pin = self._dev != "cpu" and torch.cuda.is_available()
        host = torch.zeros(2, dtype=torch.int64, pin_memory=pin)
        mine = torch.zeros(2, dtype=torch.int64, device=self._dev)
        gathered = [torch.zeros(2, dtype=torch.int64, device=self._dev) for _ in range(self.world)]
        return host, mine, gathered

    def post(self, n_tokens: int, n_docs: int) -> None:
        if not self.active:
            self._pending.append((None, None, (int(n_tokens), int(n_docs))))
            return
        import torch.distributed as dist
        host, mine, gathered = self._buffers()
        host[0] = int(n_tokens); host[1] = int(n_docs)
        mine.copy_(host, non_blocking=True)
        work = dist.all_gather(gathered, mine, async_op=True)
        self._pending.append((work, (host, mine, gathered), None))

    def wait(self):
        import torch
        work, bufs, local = self._pending.pop(0)
        if work is None:
            counts = np.asarray([local], dtype=np.int64)
            return counts, 0, 0
        work.wait()
        counts = torch.stack(bufs[2]).cpu().numpy()
        self._pool.append(bufs)
        return counts, int(counts[:self.rank, 0].sum()), int(counts[:self.rank, 1].sum())

    def drain(self):
        out = []
        while self._pending:
            out.append(self.wait())
        return out

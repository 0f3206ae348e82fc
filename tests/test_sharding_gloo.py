"""N>1 host logic on CPU: world_size-2 gloo process group, document sharding + count gather.
The encoder plugged in here is the oracle (checker) -- the test is about the sharding plumbing."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import vocab_util as vu
from tiktoken_b200.sharding import shard_ranges


def test_shard_ranges_balance_and_cover():
    off = np.asarray([0, 10, 10, 50, 120, 121, 300, 300], np.uint64)
    for world in (1, 2, 3, 8):
        r = shard_ranges(off, world)
        assert r[0][0] == 0 and r[-1][1] == len(off) - 1
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    one = shard_ranges(np.asarray([0, 1000], np.uint64), 4)          # one huge doc: replicas get nothing
    assert sum(hi - lo for lo, hi in one) == 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import Oracle
    from tools import corpus
    from tiktoken_b200.sharding import encode_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pat, ranks, special, _ = vu.load_encoding("cl100k_base", allow_real=False)
    o = Oracle(ranks, special, pat)
    text, off = corpus.config4(n_docs=4000, seed=77)
    res = encode_sharded(lambda t, d: o.encode_ordinary_batch_np(t, d, 1), text, off, rank, world)
    # the asynchronous form of the same exchange: three batches posted back to back, collected in order
    from tiktoken_b200.sharding import CountExchange, gather_counts
    x = CountExchange(rank, world)
    for k in range(3):
        x.post(100 * (rank + 1) + k, 7 + rank)
    async_ok = True
    for k, (counts, tbase, dbase) in enumerate(x.drain()):
        exp = gather_counts(100 * (rank + 1) + k, 7 + rank, rank, world)
        async_ok &= np.array_equal(counts, exp[0]) and (tbase, dbase) == (exp[1], exp[2])
    assert async_ok
    q.put((rank, res["doc_range"], res["token_base"], res["doc_base"], res["total_tokens"],
           res["tokens"].tobytes(), res["tok_off"].tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_count_exchange_single_process():
    from tiktoken_b200.sharding import CountExchange
    x = CountExchange(0, 1)
    x.post(5, 2); x.post(9, 3)
    (c0, t0, d0), (c1, t1, d1) = x.drain()
    assert c0.tolist() == [[5, 2]] and c1.tolist() == [[9, 3]] and (t0, d0, t1, d1) == (0, 0, 0, 0)


def test_world2_gloo_matches_single_process():
    from oracle import Oracle
    from tools import corpus
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pat, ranks, special, _ = vu.load_encoding("cl100k_base", allow_real=False)
    text, off = corpus.config4(n_docs=4000, seed=77)
    exp_t, exp_o = Oracle(ranks, special, pat).encode_ordinary_batch_np(text, off, 2)
    total = got[0][4]
    assert total == len(exp_t) == got[1][4]
    out = np.zeros(total, np.uint32)
    for rank, (lo, hi), tbase, dbase, _, tb, ob in got:
        t = np.frombuffer(tb, np.uint32)
        o = np.frombuffer(ob, np.uint64)
        assert dbase == lo and tbase == int(exp_o[lo])
        out[tbase:tbase + len(t)] = t
        assert np.array_equal(o + tbase, exp_o[lo:hi + 1])
    assert np.array_equal(out, exp_t)

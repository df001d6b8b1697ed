"""CPU, world_size = 2 over gloo: the N > 1 host path -- shard ranges, query broadcast, all_gather layout and merge
contract -- with the oracle as the per-shard engine (the CUDA engine needs a GPU; the collective plumbing does not)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, gauss


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import bind
    from research_new_hnsw_b200.sharded import ShardedSearcher, merge_topk_numpy, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d, k = 3001, 24, 10
    X = gauss(7, n, d)
    lo, hi = shard_range(n, rank, world)
    orc = bind.Oracle()
    shard = orc.bf_new(bind.L2, d, hi - lo)
    shard.add(X[lo:hi], np.arange(lo, hi, dtype=np.uint64))       # global labels

    def local_search(Q, kk):
        r = shard.search(Q.numpy(), kk)
        return torch.from_numpy(r["labels"].view(np.int64)), torch.from_numpy(r["dists"])

    def merge(gl, gd, kk):
        l, dd = merge_topk_numpy(gl.numpy().view(np.uint64), gd.numpy(), kk)
        return torch.from_numpy(l.view(np.int64)), torch.from_numpy(dd)

    s = ShardedSearcher(local_search, merge)
    Q = torch.from_numpy(gauss(8, 64, d)) if rank == 0 else torch.zeros(64, d)
    s.broadcast_queries(Q)
    labels, dists = s.search(Q, k)
    # every rank must hold the same merged result, equal to a single unsharded exact search
    full = orc.bf_new(bind.L2, d, n)
    full.add(X)
    ref = full.search(gauss(8, 64, d), k)
    ok = np.array_equal(labels.numpy().view(np.uint64), ref["labels"]) and np.array_equal(dists.numpy(), ref["dists"])
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    open(os.path.join(tmpdir, "rank%d" % rank), "w").write(str(int(flag.item())))
    dist.destroy_process_group()


def test_two_rank_shard_search_gloo(tmp_path, orc):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "rank0").read() == "1" and open(tmp_path / "rank1").read() == "1"


def test_shard_range_partitions_exactly():
    from research_new_hnsw_b200.sharded import shard_range
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_reference_semantics():
    from research_new_hnsw_b200.sharded import merge_topk_numpy
    inf, pad = np.float32(np.inf), np.uint64(0xFFFFFFFFFFFFFFFF)
    L = np.array([[[5, 9, pad]], [[7, 2, 11]]], np.uint64)            # [2 shards][1 query][3]
    D = np.array([[[0.5, 1.0, inf]], [[0.5, 0.75, 2.0]]], np.float32)
    l, d = merge_topk_numpy(L, D, 3)
    assert l.tolist() == [[5, 7, 2]] and d.tolist() == [[0.5, 0.5, 0.75]]   # ties -> smaller label first

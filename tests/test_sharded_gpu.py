"""GPU: the single-process sharded index of the C ABI (b200hnsw_sharded_*).

Runs on ONE GPU too (two shards on device 0 exercise routing, the packed blocks and the merge kernel); with >= 2 GPUs the
shards sit on different devices and the search kernels store their rows into the root device's buffer over NVLink.
Bar: the merged rows are EXACTLY the k best (dist, label) pairs of the rows each shard returns on its own (numpy merge),
and recall against the exact scan matches an unsharded index of the same data."""
import numpy as np
import pytest

from oracle import bind

pytestmark = pytest.mark.gpu


def _merge(rows, k):
    L = np.concatenate([r["labels"] for r in rows], 1)
    D = np.concatenate([r["dists"] for r in rows], 1)
    order = np.lexsort((L, D), axis=1)[:, :k]
    return np.take_along_axis(L, order, 1), np.take_along_axis(D, order, 1)


@pytest.mark.parametrize("n_shards", [2, 3])
def test_sharded_search_equals_merge_of_shards(lib, tmp_path, n_shards, monkeypatch):
    ndev = lib.device_count()
    devices = [s % ndev for s in range(n_shards)]
    n, d, M, efc, k = 30_000, 64, 16, 100, 10
    X = bind.lowrank_data(n, d, seed=3, latent=12, noise=0.15)
    Q = bind.lowrank_data(700, d, seed=4, latent=12, noise=0.15)
    labels = np.arange(n, dtype=np.uint64) * 3 + 1
    g = lib.ShardedHierarchicalNSW(lib.L2Space(d), n, devices, M, efc)
    g.addPoints(X[:10_000], labels[:10_000])
    g.addPoints(X[10_000:], labels[10_000:])
    g.flush()
    assert g.cur_element_count == n
    for s in range(n_shards):                                   # a label lives on shard label % n_shards
        assert g.shard(s).cur_element_count == int((labels % n_shards == s).sum())
    r = g.searchKnnBatch(Q, k, ef=64)
    parts = [g.shard(s).searchKnnBatch(Q, k, ef=64) for s in range(n_shards)]
    el, ed = _merge(parts, k)
    assert np.array_equal(r["labels"], el) and np.array_equal(r["dists"], ed)
    assert (r["counts"] == k).all()
    bf = lib.BruteforceSearch(lib.L2Space(d), n)
    bf.addPoints(X, labels)
    gt = bf.searchKnnBatch(Q, k)["labels"]
    rec = np.mean([len(set(a) & set(b)) for a, b in zip(r["labels"].tolist(), gt.tolist())]) / k
    one = lib.HierarchicalNSW(lib.L2Space(d), n, M, efc)
    one.addPoints(X, labels)
    rec1 = np.mean([len(set(a) & set(b)) for a, b in zip(one.searchKnnBatch(Q, k, ef=64)["labels"].tolist(), gt.tolist())]) / k
    assert rec >= rec1 - 0.005, (rec, rec1)                     # every shard searched with the full ef: never worse
    # per-shard files are ordinary saveIndex files; loading them back gives the same answers
    paths = [str(tmp_path / ("shard%d.bin" % s)) for s in range(n_shards)]
    g.saveIndex(paths)
    g2 = lib.ShardedHierarchicalNSW(lib.L2Space(d), paths, devices)
    r2 = g2.searchKnnBatch(Q, k, ef=64)
    assert np.array_equal(r2["labels"], r["labels"]) and np.array_equal(r2["dists"], r["dists"])
    # the copy path (no peer stores) must agree with the direct path
    monkeypatch.setenv("B200HNSW_SHARD_NO_P2P", "1")
    g3 = lib.ShardedHierarchicalNSW(lib.L2Space(d), paths, devices)
    r3 = g3.searchKnnBatch(Q, k, ef=64)
    assert np.array_equal(r3["labels"], r["labels"]) and np.array_equal(r3["dists"], r["dists"])
    # re-adding a label updates it on its shard
    g.addPoints(X[:5] + 1.0, labels[:5])
    assert g.cur_element_count == n


def test_sharded_small_k_larger_than_a_shard(lib):
    d = 8
    g = lib.ShardedHierarchicalNSW(lib.L2Space(d), 10, [0, 0], 4, 20)
    X = np.eye(d, dtype=np.float32)[:5]
    g.addPoints(X)                                              # labels 0..4: shard 0 holds 0,2,4; shard 1 holds 1,3
    r = g.searchKnnBatch(X[:2], 4, ef=10)
    assert r["labels"][0, 0] == 0 and r["labels"][1, 0] == 1
    assert (r["counts"] == 4).all()
    r = g.searchKnnBatch(X[:1], 8, ef=10)                       # more than stored: padded rows
    assert r["counts"][0] == 5 and (r["labels"][0, 5:] == np.uint64(2**64 - 1)).all()

"""GPU parity: BruteforceSearch through the C ABI.  Bar: ids bit-exact (ties broken by id), distances within 1e-5
relative -- the exact-scan kernel sums in the reference's SSE lane order, so they are in fact bit-identical."""
import hashlib

import numpy as np
import pytest

from conftest import gauss
from oracle import bind

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric,n,d,k", [(bind.L2, 5000, 128, 10), (bind.IP, 3000, 96, 100), (bind.L2, 2000, 30, 7),
                                          (bind.IP, 1500, 17, 3), (bind.L2, 700, 3, 5), (bind.IP, 4097, 768, 100)])
def test_exact_vs_oracle(lib, orc, ref, metric, n, d, k):
    X = gauss(21, n, d)
    if metric == bind.IP:
        X /= np.linalg.norm(X, axis=1, keepdims=True)
    X[n // 2] = X[n // 3]  # exact duplicate rows -> equal distances, order decided by label
    Q = gauss(22, 130, d)
    labels = (np.arange(n, dtype=np.uint64) * 7 + 3)
    space = lib.L2Space(d) if metric == bind.L2 else lib.InnerProductSpace(d)
    g = lib.BruteforceSearch(space, n)
    g.addPoints(X, labels)
    c = orc.bf_new(metric, d, n)
    c.add(X, labels)
    rg, rc = g.searchKnnBatch(Q, k), c.search(Q, k)
    assert np.array_equal(rg["labels"], rc["labels"])
    assert np.array_equal(rg["dists"], rc["dists"])
    assert np.array_equal(rg["counts"], rc["counts"])
    if ref is not None:
        r = ref.bf_new(metric, d, n)
        r.add(X, labels)
        rr = r.search(Q, k)
        assert np.array_equal(rg["labels"], rr["labels"]) and np.array_equal(rg["dists"], rr["dists"])


def test_add_overwrite_remove_save_load(lib, orc, tmp_path):
    d, n = 20, 600
    X = gauss(31, n, d)
    Q = gauss(32, 40, d)
    g = lib.BruteforceSearch(lib.L2Space(d), n)
    c = orc.bf_new(bind.L2, d, n)
    for eng in (g, c):
        add = eng.addPoints if eng is g else eng.add
        add(X[:500])
        add(X[500:510], np.arange(10, 20, dtype=np.uint64))       # existing labels: rows overwritten
        (eng.removePoint if eng is g else eng.remove)(3)           # last row swapped in
        (eng.removePoint if eng is g else eng.remove)(499)
    assert g.cur_element_count == c.count() == 498
    rg, rc = g.searchKnnBatch(Q, 12), c.search(Q, 12)
    assert np.array_equal(rg["labels"], rc["labels"]) and np.array_equal(rg["dists"], rc["dists"])
    pg, pc = str(tmp_path / "g.bin"), str(tmp_path / "c.bin")
    g.saveIndex(pg)
    c.save(pc)
    sha = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest()
    assert sha(pg) == sha(pc)
    g2 = lib.BruteforceSearch(lib.L2Space(d), pg)
    r2 = g2.searchKnnBatch(Q, 12)
    assert np.array_equal(r2["labels"], rc["labels"])
    with pytest.raises(lib.B200Error, match="exceeds the specified limit"):
        g.addPoints(gauss(33, 200, d), np.arange(10_000, 10_200, dtype=np.uint64))


def test_k_larger_than_count(lib):
    g = lib.BruteforceSearch(lib.L2Space(4), 10)
    X = np.eye(4, dtype=np.float32)[:3]
    g.addPoints(X)
    r = g.searchKnnBatch(X[:1], 6)
    assert r["counts"][0] == 3 and r["labels"][0, :3].tolist() == [0, 1, 2]
    assert np.isinf(r["dists"][0, 3:]).all()


@pytest.mark.parametrize("metric,n,d,k,nq", [(bind.L2, 40000, 128, 10, 300), (bind.IP, 30000, 768, 100, 130),
                                             (bind.L2, 20000, 100, 5, 257), (bind.IP, 33000, 30, 20, 64)])
@pytest.mark.parametrize("cta_group", [1, 2])
def test_tensor_core_path_is_exact(lib, orc, monkeypatch, metric, n, d, k, nq, cta_group):
    """tcgen05 GEMM candidate generation + exact re-rank (csrc/bf_tensor.cu): same bar as the scan -- ids bit-exact,
    distances bit-identical -- because the re-rank uses the reference's summation order.  cta_group = 2: the CTA-pair
    form of the GEMM (tcgen05 cta_group::2; 20000 rows = an odd number of 128-row panels, so one pair has a phantom
    panel)."""
    monkeypatch.setenv("B200HNSW_BF_PATH", "tensor")
    monkeypatch.setenv("B200HNSW_BF_CG", str(cta_group))
    X = bind.lowrank_data(n, d, seed=51, latent=24, noise=0.2, normalize=(metric == bind.IP))
    X[n // 2] = X[n // 3]
    Q = bind.lowrank_data(nq, d, seed=52, latent=24, noise=0.2, normalize=(metric == bind.IP))
    labels = (np.arange(n, dtype=np.uint64) * 3 + 1)
    space = lib.L2Space(d) if metric == bind.L2 else lib.InnerProductSpace(d)
    g = lib.BruteforceSearch(space, n)
    g.addPoints(X, labels)
    rg = g.searchKnnBatch(Q, k)
    assert g.stats()["hops_base"] == 1, "tensor path did not run"
    c = orc.bf_new(metric, d, n)
    c.add(X, labels)
    rc = c.search(Q[:40], k)
    assert np.array_equal(rg["labels"][:40], rc["labels"])
    assert np.array_equal(rg["dists"][:40], rc["dists"])
    monkeypatch.setenv("B200HNSW_BF_PATH", "scan")
    rs = g.searchKnnBatch(Q, k)                                  # all queries against the exact scan kernel
    assert g.stats()["hops_base"] == 0
    assert np.array_equal(rg["labels"], rs["labels"]) and np.array_equal(rg["dists"], rs["dists"])
    assert np.array_equal(rg["counts"], rs["counts"])


@pytest.mark.parametrize("metric,n,d,k", [(bind.L2, 5000, 128, 10), (bind.IP, 3001, 96, 100), (bind.L2, 2000, 30, 7),
                                          (bind.IP, 1500, 17, 3), (bind.L2, 700, 3, 5), (bind.IP, 4097, 768, 100),
                                          (bind.L2, 20000, 100, 1), (bind.IP, 50, 768, 100)])
def test_streaming_path_is_exact(lib, orc, monkeypatch, metric, n, d, k):
    """Small query batches stream the rows once (csrc/bf_stream.cu: bulk-copy ring, lane = row, the reference's SSE
    summation order): ids and distances bit-identical to the oracle for every pass shape (1, 2, 4, 8, 16 queries per
    pass, several passes), ragged last row tile, dims with sequential tails, duplicate rows, k > count."""
    monkeypatch.setenv("B200HNSW_BF_PATH", "stream")
    X = gauss(61, n, d)
    if metric == bind.IP:
        X /= np.linalg.norm(X, axis=1, keepdims=True)
    X[n // 2] = X[n // 3]
    labels = (np.arange(n, dtype=np.uint64) * 5 + 2)
    space = lib.L2Space(d) if metric == bind.L2 else lib.InnerProductSpace(d)
    g = lib.BruteforceSearch(space, n)
    g.addPoints(X, labels)
    c = orc.bf_new(metric, d, n)
    c.add(X, labels)
    Qall = gauss(62, 41, d)
    Qall[0] = X[n // 3]                                        # query equal to the duplicated rows: an exact tie
    for nq in (1, 2, 3, 5, 8, 9, 16, 17, 41):
        Q = Qall[:nq]
        rg, rc = g.searchKnnBatch(Q, k), c.search(Q, k)
        assert g.stats()["hops_base"] == 2, "streaming path did not run"
        assert np.array_equal(rg["labels"], rc["labels"]), (nq,)
        assert np.array_equal(rg["dists"], rc["dists"]), (nq,)
        assert np.array_equal(rg["counts"], rc["counts"])


def test_partial_merge_tree_many_slices(lib, orc, monkeypatch):
    """Per-CTA partial lists are reduced by merge_launch.cuh: 64 slices x k=100 (tiled scan) and 148 x 100 (streaming
    scan) take the one-CTA-per-query selection kernel, 148 x 200 = 29 600 candidates the multi-level tree."""
    monkeypatch.setenv("B200HNSW_BF_PATH", "scan")
    n, d, k = 300_000, 16, 100
    X = gauss(71, n, d)
    Q = gauss(72, 70, d)
    g = lib.BruteforceSearch(lib.L2Space(d), n)
    g.addPoints(X)
    c = orc.bf_new(bind.L2, d, n)
    c.add(X)
    rg, rc = g.searchKnnBatch(Q, k), c.search(Q[:8], k)
    assert g.stats()["hops_base"] == 0
    assert np.array_equal(rg["labels"][:8], rc["labels"]) and np.array_equal(rg["dists"][:8], rc["dists"])
    monkeypatch.setenv("B200HNSW_BF_PATH", "stream")
    rs = g.searchKnnBatch(Q, k)
    assert np.array_equal(rg["labels"], rs["labels"]) and np.array_equal(rg["dists"], rs["dists"])
    k2 = 200                                                     # too many candidates for the selection kernel
    rt, rc2 = g.searchKnnBatch(Q[:5], k2), c.search(Q[:5], k2)
    assert g.stats()["hops_base"] == 2
    assert np.array_equal(rt["labels"], rc2["labels"]) and np.array_equal(rt["dists"], rc2["dists"])


@pytest.mark.parametrize("path,n,d,k,nq", [("scan", 6000, 48, 10, 70), ("stream", 20000, 100, 7, 5), ("tensor", 40000, 128, 20, 300)])
def test_filter_functor_is_a_row_mask_in_every_path(lib, orc, monkeypatch, path, n, d, k, nq):
    """searchKnn(query, k, isIdAllowed) (bruteforce.h:106-135, filter at :114 and :121): rows whose label the functor
    rejects are skipped.  Equal to an unfiltered search over an index that holds only the allowed rows -- ids and
    distances bit-identical -- on the tiled scan, the streaming scan and the tcgen05 path."""
    monkeypatch.setenv("B200HNSW_BF_PATH", path)
    X = bind.lowrank_data(n, d, seed=71, latent=16, noise=0.2)
    Q = bind.lowrank_data(nq, d, seed=72, latent=16, noise=0.2)
    labels = np.arange(n, dtype=np.uint64) * 5 + 3
    allowed = lambda l: l % 3 != 0
    g = lib.BruteforceSearch(lib.L2Space(d), n)
    g.addPoints(X, labels)
    r = g.searchKnnFiltered(Q, k, allowed)
    assert g.stats()["hops_base"] == {"scan": 0, "tensor": 1, "stream": 2}[path]
    keep = np.array([allowed(int(l)) for l in labels])
    c = orc.bf_new(bind.L2, d, int(keep.sum()))
    c.add(X[keep], labels[keep])
    rc = c.search(Q, k)
    assert np.array_equal(r["labels"], rc["labels"]) and np.array_equal(r["dists"], rc["dists"])
    assert (r["counts"] == k).all()
    r0 = g.searchKnnBatch(Q, k)                                  # the mask does not outlive its call
    c0 = orc.bf_new(bind.L2, d, n)
    c0.add(X, labels)
    assert np.array_equal(r0["labels"], c0.search(Q, k)["labels"])
    few = g.searchKnnFiltered(Q[:3], k, lambda l: l < 3 + 5 * 4)  # fewer allowed rows than k: padded result
    assert (few["counts"] == 4).all() and (few["labels"][:, 4:] == np.uint64(2**64 - 1)).all()

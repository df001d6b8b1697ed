"""CPU: pin the oracle (own restatement) to the reference's known answers, golden outputs and, when the compiled
reference is present (oracle/_ref), to the reference itself on fresh seeded inputs."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, gauss
from oracle import bind


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def test_distance_known_answers(orc, golden):
    _, g = golden
    off = 0
    for i, d in enumerate(g["dist_dims"]):
        a = g["dist_a"][off:off + d]
        b = g["dist_b"][off:off + d]
        off += d
        assert np.float32(orc.dist(bind.L2, a, b)) == g["dist_l2"][i], d   # bit-exact
        assert np.float32(orc.dist(bind.IP, a, b)) == g["dist_ip"][i], d


@pytest.mark.parametrize("name", ["l2_n2000_d16_M8", "ip_n1500_d24_M6", "l2_n1200_d13_M5"])
def test_search_matches_reference_dump(orc, golden, name):
    meta, g = golden
    m = meta[name]
    path = os.path.join(GOLDEN, name + ".bin")
    assert _sha(path) == m["sha256"]
    idx = orc.hnsw_load(m["metric"], m["d"], path)
    info = idx.info()
    assert (info["maxlevel"], info["enterpoint"]) == (m["maxlevel"], m["enterpoint"])
    Q = g[name + "/Q"]
    for ef in m["efs"]:
        r = idx.search(Q, 10, ef)
        assert np.array_equal(r["labels"], g["%s/ef%d/labels" % (name, ef)])
        assert np.array_equal(r["dists"], g["%s/ef%d/dists" % (name, ef)])      # bit-exact distances
        assert np.array_equal(r["D"], g["%s/ef%d/D" % (name, ef)])              # same work, eval for eval
        assert np.array_equal(r["Hup"], g["%s/ef%d/Hup" % (name, ef)])


@pytest.mark.parametrize("name", ["l2_n2000_d16_M8", "ip_n1500_d24_M6", "l2_n1200_d13_M5"])
def test_save_is_byte_identical(orc, golden, name, tmp_path):
    meta, _ = golden
    m = meta[name]
    idx = orc.hnsw_load(m["metric"], m["d"], os.path.join(GOLDEN, name + ".bin"))
    out = str(tmp_path / "resaved.bin")
    idx.save(out)
    assert _sha(out) == m["sha256"]


@pytest.mark.parametrize("name", ["l2_n2000_d16_M8", "ip_n1500_d24_M6", "l2_n1200_d13_M5"])
def test_bruteforce_matches_reference_dump(orc, golden, name):
    meta, g = golden
    m = meta[name]
    idx = orc.hnsw_load(m["metric"], m["d"], os.path.join(GOLDEN, name + ".bin"))
    n = idx.info()["cur_element_count"]
    # the vectors live inside the index file: pull them back out through the restated layout
    raw = np.fromfile(os.path.join(GOLDEN, name + ".bin"), dtype=np.uint8)[96:96 + n * idx.info()["size_data_per_element"]]
    rec = raw.reshape(n, -1)
    off = 4 + 4 * 2 * m["M"]
    X = rec[:, off:off + 4 * m["d"]].copy().view(np.float32)
    bf = orc.bf_new(m["metric"], m["d"], n)
    bf.add(X)
    r = bf.search(g[name + "/Q"], 10)
    assert np.array_equal(r["labels"], g[name + "/bf/labels"])
    assert np.array_equal(r["dists"], g[name + "/bf/dists"])


def test_load_errors(orc, tmp_path):
    with pytest.raises(RuntimeError, match="Cannot open file"):
        orc.hnsw_load(bind.L2, 16, str(tmp_path / "missing.bin"))
    data = open(os.path.join(GOLDEN, "l2_n2000_d16_M8.bin"), "rb").read()
    bad = tmp_path / "trunc.bin"
    bad.write_bytes(data[:-7])
    with pytest.raises(RuntimeError, match="corrupted or unsupported"):
        orc.hnsw_load(bind.L2, 16, str(bad))


# ---- against the compiled reference itself (skipped where oracle/_ref is absent) -------------------------------
def test_build_reproduces_survey_known_answer(orc, ref, golden, tmp_path):
    """test.cpp / index_builder at N=10000, d=128, M=16, efc=200: sha256 + entry/max_level."""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    meta, g = golden
    X = ref.gen_gaussian(123, 10000, 128)
    assert np.array_equal(X[:4], g["gauss123_head"])
    idx = orc.hnsw_new(bind.L2, 128, 10000, 16, 200)
    idx.add(X)
    out = str(tmp_path / "orc10k.bin")
    idx.save(out)
    k = meta["survey_10k"]
    assert os.path.getsize(out) == k["bytes"] == 6606132
    assert _sha(out) == k["sha256"]
    info = idx.info()
    assert (info["enterpoint"], info["maxlevel"]) == (4373, 3)


@pytest.mark.parametrize("name", ["l2_n2000_d16_M8", "ip_n1500_d24_M6", "l2_n1200_d13_M5"])
def test_update_and_replace_deleted_match_reference_fixture(orc, golden, name, tmp_path):
    """updatePoint (addPoint with an existing label, hnswalg.h:1157-1174 -> 995-1139) and replace_deleted (:954-992):
    the saved file after the reference applied the seeded changes (tests/golden/update_golden.json, written by
    make_golden.py) must be reproduced byte for byte by the restatement starting from the committed index file."""
    sys.path.insert(0, GOLDEN)
    import make_golden
    meta, _ = golden
    m = meta[name]
    u = make_golden.update_inputs(dict(n=m["n"], d=m["d"]))
    fx = json.load(open(os.path.join(GOLDEN, "update_golden.json")))[name]
    src = os.path.join(GOLDEN, name + ".bin")
    out = str(tmp_path / "changed.bin")
    idx = orc.hnsw_load(m["metric"], m["d"], src)
    idx.add(u["Xn"], u["upd"])
    idx.save(out)
    assert _sha(out) == fx["update_sha256"]
    idx = orc.hnsw_load(m["metric"], m["d"], src, allow_replace_deleted=True)
    for l in u["dead"].tolist():
        idx.mark_delete(l)
    idx.add_replace_deleted(u["Xr"], u["new_labels"])
    idx.save(out)
    assert _sha(out) == fx["replace_deleted_sha256"]
    plain = orc.hnsw_load(m["metric"], m["d"], src)
    with pytest.raises(RuntimeError, match="disabled in constructor"):
        plain.add_replace_deleted(u["Xr"][:1], u["new_labels"][:1])


@pytest.mark.parametrize("metric,d,M,efc", [(bind.L2, 20, 6, 40), (bind.IP, 33, 9, 64), (bind.L2, 7, 4, 30)])
def test_differential_vs_reference(orc, ref, tmp_path, metric, d, M, efc):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    X = gauss(5, 1500, d)
    Q = gauss(6, 100, d)
    a = ref.hnsw_new(metric, d, 1500, M, efc, counting=True)
    a.add(X)
    b = orc.hnsw_new(metric, d, 1500, M, efc)
    b.add(X)
    pa, pb = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    a.save(pa)
    b.save(pb)
    assert _sha(pa) == _sha(pb)
    for ef in (10, 37, 150):
        ra = a.search(Q, 10, ef, counters=True)
        rb = b.search(Q, 10, ef)
        assert np.array_equal(ra["labels"], rb["labels"])
        assert np.array_equal(ra["dists"], rb["dists"])
        assert np.array_equal(ra["D"], rb["D"])
        assert np.array_equal(ra["Hup"], rb["Hup"])
    fa, fb = ref.bf_new(metric, d, 1500), orc.bf_new(metric, d, 1500)
    fa.add(X)
    fb.add(X)
    ra, rb = fa.search(Q, 25), fb.search(Q, 25)
    assert np.array_equal(ra["labels"], rb["labels"]) and np.array_equal(ra["dists"], rb["dists"])
    # updates and replace_deleted against the live reference (same calls on both, files byte-identical)
    rng = np.random.default_rng(11)
    upd = rng.choice(1500, 150, replace=False).astype(np.uint64)
    Xn = gauss(12, 150, d)
    a.add(Xn, upd)
    b.add(Xn, upd)
    a.save(pa)
    b.save(pb)
    assert _sha(pa) == _sha(pb)
    a2 = ref.hnsw_new(metric, d, 1500, M, efc, allow_replace_deleted=True)
    b2 = orc.hnsw_new(metric, d, 1500, M, efc, allow_replace_deleted=True)
    a2.add(X)
    b2.add(X)
    for l in rng.choice(1500, 100, replace=False).tolist():
        a2.mark_delete(l)
        b2.mark_delete(l)
    nl = np.arange(50_000, 50_100, dtype=np.uint64)
    a2.add_replace_deleted(gauss(13, 100, d), nl)
    b2.add_replace_deleted(gauss(13, 100, d), nl)
    a2.save(pa)
    b2.save(pb)
    assert _sha(pa) == _sha(pb)
    fa.remove(3)
    fb.remove(3)
    fa.save(str(tmp_path / "fa.bin"))
    fb.save(str(tmp_path / "fb.bin"))
    assert _sha(str(tmp_path / "fa.bin")) == _sha(str(tmp_path / "fb.bin"))

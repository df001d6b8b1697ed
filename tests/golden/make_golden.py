"""Generates tests/golden/* from the UNMODIFIED reference (oracle/_ref, compiled from /root/reference).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference ships no tests or fixtures of its own, so these vectors are outputs of the
reference itself run here on seeded inputs, plus SURVEY.md's known answers (sha256 of the N=10000 index,
entry/max-level printed by test.cpp).  Everything is small enough to commit.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bind  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def update_inputs(c):
    """Seeded inputs of the updatePoint / replace_deleted fixtures (shared with tests/test_oracle.py)."""
    rng = np.random.default_rng(789 + c["n"])
    nu = c["n"] // 10
    return dict(nu=nu, upd=rng.choice(c["n"], nu, replace=False).astype(np.uint64),
                Xn=rng.standard_normal((nu, c["d"]), dtype=np.float32),
                dead=rng.choice(c["n"], nu, replace=False),
                new_labels=np.arange(100000, 100000 + nu, dtype=np.uint64),
                Xr=rng.standard_normal((nu, c["d"]), dtype=np.float32))


UPDATE_CASES = [dict(name="l2_n2000_d16_M8", metric=bind.L2, n=2000, d=16, M=8, efc=100),
                dict(name="ip_n1500_d24_M6", metric=bind.IP, n=1500, d=24, M=6, efc=60),
                dict(name="l2_n1200_d13_M5", metric=bind.L2, n=1200, d=13, M=5, efc=40)]


def update_golden(ref):
    """sha256 of the reference's saveIndex file after (a) re-adding n/10 existing labels with new vectors
    (addPoint -> updatePoint, hnswalg.h:1157-1174, 995-1139) and (b) marking n/10 elements deleted and adding n/10 new
    labels with replace_deleted = true (hnswalg.h:954-992) -> tests/golden/update_golden.json."""
    out = {}
    for c in UPDATE_CASES:
        X = ref.gen_gaussian(123, c["n"], c["d"])
        if c["metric"] == bind.IP:
            X /= np.linalg.norm(X, axis=1, keepdims=True)
        u = update_inputs(c)
        tmp = os.path.join(OUT, "_tmp_update.bin")
        idx = ref.hnsw_new(c["metric"], c["d"], c["n"], c["M"], c["efc"], 100)
        idx.add(X)
        idx.add(u["Xn"], u["upd"])
        idx.save(tmp)
        sha_u = hashlib.sha256(open(tmp, "rb").read()).hexdigest()
        idx = ref.hnsw_new(c["metric"], c["d"], c["n"], c["M"], c["efc"], 100, allow_replace_deleted=True)
        idx.add(X)
        for l in u["dead"].tolist():
            idx.mark_delete(l)
        idx.add_replace_deleted(u["Xr"], u["new_labels"])
        idx.save(tmp)
        sha_r = hashlib.sha256(open(tmp, "rb").read()).hexdigest()
        os.remove(tmp)
        out[c["name"]] = dict(update_sha256=sha_u, replace_deleted_sha256=sha_r)
    json.dump(out, open(os.path.join(OUT, "update_golden.json"), "w"), indent=1, sort_keys=True)
    return out


def main():
    bind.build()
    ref = bind.Ref("sse")
    assert ref.simd_level() == "sse"
    if "--updates-only" in sys.argv:
        print(update_golden(ref))
        return
    update_golden(ref)
    meta = {}
    arrays = {}

    # ---- distance known answers: every branch of the dispatch ladder (space_l2.h:214-238, space_ip.h:348-383)
    dims = [1, 2, 3, 4, 5, 7, 8, 12, 15, 16, 17, 20, 23, 31, 32, 33, 48, 63, 64, 96, 100, 128, 130, 768]
    rng = np.random.default_rng(7)
    dist_a, dist_b, dist_l2, dist_ip = [], [], [], []
    for d in dims:
        a = rng.standard_normal(d).astype(np.float32)
        b = rng.standard_normal(d).astype(np.float32)
        dist_a.append(a)
        dist_b.append(b)
        dist_l2.append(ref.dist(bind.L2, a, b))
        dist_ip.append(ref.dist(bind.IP, a, b))
    arrays["dist_dims"] = np.array(dims, np.int32)
    arrays["dist_a"] = np.concatenate(dist_a)
    arrays["dist_b"] = np.concatenate(dist_b)
    arrays["dist_l2"] = np.array(dist_l2, np.float32)
    arrays["dist_ip"] = np.array(dist_ip, np.float32)

    # ---- small graphs built serially by the reference (the build.cpp:137-145 pattern) + search dumps
    cases = [
        dict(name="l2_n2000_d16_M8", metric=bind.L2, n=2000, d=16, M=8, efc=100, efs=[10, 50, 200]),
        dict(name="ip_n1500_d24_M6", metric=bind.IP, n=1500, d=24, M=6, efc=60, efs=[10, 40]),
        dict(name="l2_n1200_d13_M5", metric=bind.L2, n=1200, d=13, M=5, efc=40, efs=[10, 64]),
    ]
    for c in cases:
        X = ref.gen_gaussian(123, c["n"], c["d"])
        Q = ref.gen_gaussian(456, 200, c["d"])
        if c["metric"] == bind.IP:
            X /= np.linalg.norm(X, axis=1, keepdims=True)
        idx = ref.hnsw_new(c["metric"], c["d"], c["n"], c["M"], c["efc"], 100)
        idx.add(X)
        path = os.path.join(OUT, c["name"] + ".bin")
        idx.save(path)
        info = idx.info()
        meta[c["name"]] = dict(metric=c["metric"], n=c["n"], d=c["d"], M=c["M"], efc=c["efc"], efs=c["efs"],
                               sha256=hashlib.sha256(open(path, "rb").read()).hexdigest(),
                               maxlevel=info["maxlevel"], enterpoint=info["enterpoint"])
        arrays[c["name"] + "/Q"] = Q
        cnt = ref.hnsw_load(c["metric"], c["d"], path, counting=True)
        for ef in c["efs"]:
            r = cnt.search(Q, 10, ef, counters=True)
            arrays["%s/ef%d/labels" % (c["name"], ef)] = r["labels"]
            arrays["%s/ef%d/dists" % (c["name"], ef)] = r["dists"]
            arrays["%s/ef%d/D" % (c["name"], ef)] = r["D"]
            arrays["%s/ef%d/Hup" % (c["name"], ef)] = r["Hup"]
        bf = ref.bf_new(c["metric"], c["d"], c["n"])
        bf.add(X)
        r = bf.search(Q, 10)
        arrays[c["name"] + "/bf/labels"] = r["labels"]
        arrays[c["name"] + "/bf/dists"] = r["dists"]

    # ---- known answers from SURVEY.md section 4 (test.cpp / index_builder at N=10000, d=128, M=16, efc=200)
    X = ref.gen_gaussian(123, 10000, 128)
    idx = ref.hnsw_new(bind.L2, 128, 10000, 16, 200, 100)
    idx.add(X)
    tmp = "/tmp/golden_10k.bin"
    idx.save(tmp)
    info = idx.info()
    meta["survey_10k"] = dict(sha256=hashlib.sha256(open(tmp, "rb").read()).hexdigest(), bytes=os.path.getsize(tmp),
                              entry=info["enterpoint"], max_level=info["maxlevel"],
                              data_head=[float(v) for v in X[0, :4]])
    assert meta["survey_10k"]["sha256"] == "01816aef966a07ee317b2e36f4baeaf9ffbb865649dde5166ecbaf4c193a7abb"
    assert (info["enterpoint"], info["maxlevel"]) == (4373, 3)
    os.remove(tmp)
    arrays["gauss123_head"] = X[:4].copy()  # pins the libstdc++ generator stream

    np.savez_compressed(os.path.join(OUT, "golden.npz"), **arrays)
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

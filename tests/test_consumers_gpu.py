"""GPU: the reference's own programs, compiled UNCHANGED against the drop-in hnswlib header
(research_new_hnsw_b200/hnswlib/Makefile, binaries prebuilt in the build container), run on the GPU engine."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import bind

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "research_new_hnsw_b200", "hnswlib", "_consumers")


def _need(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.skip("%s not prebuilt (needs /root/reference at build time)" % name)
    return p


def test_reference_test_cpp_prints_the_known_answer():
    """test.cpp:3-26 -- the same line the reference build prints."""
    r = subprocess.run([_need("test")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Exporting adjacency: nodes=10000, entry=4373, max_level=3" in r.stdout


def test_shim_api_program():
    """tests/cpp/shim_api.cpp over the drop-in header: filter functors on both index types, updatePoint, markDelete /
    replace_deleted, stop_condition.h, parallel addPoint / searchKnn (the parts of the vendored API no reference
    consumer touches)."""
    r = subprocess.run([_need("shim_api")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "shim_api OK" in r.stdout, r.stdout + r.stderr


def test_reference_index_builder_end_to_end(orc, ref, tmp_path):
    """index_builder/build.cpp:110-154 unchanged: addPoint x N -> saveIndex -> export_adjacency through raw link-list
    pointers and public fields.  The file has the reference's exact size, loads in the reference, and searches well."""
    out = str(tmp_path / "g.bin")
    n, d, M, efc = 10000, 128, 16, 200
    r = subprocess.run([_need("index_builder"), str(n), str(d), str(tmp_path / "db"), out, str(M), str(efc)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "HNSW index saved" in r.stderr and "export_adjacency: written 10000 nodes" in r.stderr
    ref_size = 6606132                                    # SURVEY.md section 4 (depends only on levels, not on links)
    assert os.path.getsize(out) == ref_size
    back = ref.hnsw_load(bind.L2, d, out) if ref is not None else orc.hnsw_load(bind.L2, d, out)
    info = back.info()
    assert (info["cur_element_count"], info["enterpoint"], info["maxlevel"]) == (n, 4373, 3)
    # .adj header: entry, max_level, node_count (build.cpp:40-46)
    hdr = np.fromfile(out + ".adj", dtype=np.uint32, count=3)
    assert hdr.tolist() == [4373, 3, n]
    if ref is not None:
        X = ref.gen_gaussian(123, n, d)                  # the data index_builder generated (build.cpp:124-138)
        Q = ref.gen_gaussian(456, 200, d)
        cpu = ref.hnsw_new(bind.L2, d, n, M, efc)
        cpu.add(X)
        bf = orc.bf_new(bind.L2, d, n)
        bf.add(X)
        gt = bf.search(Q, 10)["labels"]
        rec = lambda lab: np.mean([len(set(a) & set(b)) for a, b in zip(lab.tolist(), gt.tolist())]) / 10
        r_gpu_graph = rec(back.search(Q, 10, 200)["labels"])
        r_cpu_graph = rec(cpu.search(Q, 10, 200)["labels"])
        assert r_gpu_graph >= r_cpu_graph - 0.01, (r_gpu_graph, r_cpu_graph)


def test_reference_hnsw_service_over_http(orc, tmp_path):
    """hnsw_service/main.cpp:49-96 (normal mode) unchanged: loadIndex + POST /search -> setEf + searchKnn, results
    furthest first (main.cpp:71-75).  RLIMIT_AS (main.cpp:19-22) is lifted by the preload shim."""
    import json
    import socket
    import time
    import urllib.request
    exe = _need("hnsw_service")
    n, d = 5000, 128
    X = np.random.default_rng(3).standard_normal((n, d), dtype=np.float32)
    cpu = orc.hnsw_new(bind.L2, d, n, 16, 100)
    cpu.add(X)
    g = str(tmp_path / "svc.bin")
    cpu.save(g)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, LD_PRELOAD=os.path.join(ROOT, "research_new_hnsw_b200", "libb200_norlimit.so"))
    p = subprocess.Popen([exe, "--graph", g, "--port", str(port), "--dim", str(d)], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(200):
            try:
                info = json.load(urllib.request.urlopen("http://127.0.0.1:%d/info" % port, timeout=1))
                break
            except Exception:
                assert p.poll() is None, p.stderr.read()
                time.sleep(0.1)
        assert info["nodes"] == n and info["dim"] == d
        Q = np.random.default_rng(4).standard_normal((20, d), dtype=np.float32)
        want = cpu.search(Q, 5, 64)
        for i in range(len(Q)):
            body = json.dumps({"query": Q[i].tolist(), "k": 5, "ef": 64}).encode()
            req = urllib.request.Request("http://127.0.0.1:%d/search" % port, data=body)
            res = json.load(urllib.request.urlopen(req, timeout=10))["results"]
            assert [r["id"] for r in res] == want["labels"][i][::-1].tolist()      # furthest first
            assert np.allclose([r["distance"] for r in res], want["dists"][i][::-1], rtol=1e-5)
    finally:
        p.kill()

"""Ad-hoc perf probe (not part of the bench contract): CPU-reference-built graph -> GPU search sweep."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_new_hnsw_b200 as pkg
from oracle import bind

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
d, M, efc, nq, k = 128, 32, 200, 10000, 10
threads = os.cpu_count()
print("host threads", threads, "ref level", bind.best_ref_level(), flush=True)
X = bind.lowrank_data(n, d, seed=1)
Q = bind.lowrank_data(nq, d, seed=2)
ref = bind.Ref(bind.best_ref_level())
path = "/tmp/probe_%d.bin" % n
if not os.path.exists(path):
    idx = ref.hnsw_new(bind.L2, d, n, M, efc)
    sec = idx.add(X, threads=threads)
    print("cpu build %.1fs = %.0f pts/s on %d threads" % (sec, n / sec, threads), flush=True)
    idx.save(path)
cpu = ref.hnsw_load(bind.L2, d, path)
t = time.time(); gpu = pkg.HierarchicalNSW(pkg.L2Space(d), path); print("gpu load %.2fs" % (time.time() - t), flush=True)
bf = pkg.BruteforceSearch(pkg.L2Space(d), n); bf.addPoints(X)
t = time.time(); gt = bf.searchKnnBatch(Q, k)["labels"]; print("gpu bf gt %.2fs kernel %.1f ms" % (time.time() - t, bf.stats()["last_kernel_ms"]), flush=True)
for ef in (16, 32, 64, 128, 256):
    for _ in range(2):
        r = gpu.searchKnnBatch(Q, k, ef=ef, work=True)
    st = gpu.stats()
    rec = np.mean([len(set(a) & set(b)) for a, b in zip(r["labels"].tolist(), gt.tolist())]) / k
    c = cpu.search(Q, k, ef, threads=threads)
    crec = np.mean([len(set(a) & set(b)) for a, b in zip(c["labels"].tolist(), gt.tolist())]) / k
    same = np.mean([set(a) == set(b) for a, b in zip(r["labels"].tolist(), c["labels"].tolist())])
    B = st["dist_evals"] * d * 4 + st["hops_base"] * (4 + 8 * M) + st["hops_upper"] * (4 + 4 * M) + nq * (4 * d + 12 * k)
    ms = st["last_kernel_ms"]
    print("ef %3d  gpu %.3f ms  %.2f MQPS  recall %.4f | D/q %.0f H0/q %.1f resets %d  %.0f GB/s | cpu %.0f QPS recall %.4f | same-sets %.4f"
          % (ef, ms, nq / ms / 1e3, rec, st["dist_evals"] / nq, st["hops_base"] / nq, st["visited_resets"], B / ms / 1e6,
             nq / c["seconds"], crec, same), flush=True)

"""Ad-hoc probe: GPU build quality/time vs batch ratio, against the CPU-built graph."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_new_hnsw_b200 as pkg
from oracle import bind

def recall(l, gt): return float(np.mean([len(set(a) & set(b)) for a, b in zip(l.tolist(), gt.tolist())]) / gt.shape[1])

def run(name, metric, X, Q, M, efc, ratios, efs):
    n, d = X.shape
    space = pkg.L2Space(d) if metric == 0 else pkg.InnerProductSpace(d)
    bf = pkg.BruteforceSearch(space, n); bf.addPoints(X); gt = bf.searchKnnBatch(Q, 10)["labels"]; del bf
    ref = bind.Ref(bind.best_ref_level())
    c = ref.hnsw_new(metric, d, n, M, efc); sec = c.add(X, threads=os.cpu_count())
    print("%s n=%d d=%d M=%d efc=%d | cpu build %.1fs (%.0f pts/s) recall %s" % (name, n, d, M, efc, sec, n / sec,
          ["%.4f" % recall(c.search(Q, 10, ef, threads=16)["labels"], gt) for ef in efs]), flush=True)
    for ratio in ratios:
        os.environ["B200HNSW_BUILD_RATIO"] = str(ratio)
        g = pkg.HierarchicalNSW(space, n, M, efc)
        t = time.time(); g.addPoints(X); g.flush(); sec = time.time() - t
        st = g.stats()
        print("   ratio %3d: gpu build %.2fs (%.0f pts/s, kernels %.0f ms, launches %d, D/pt %.0f resets/pt %.2f) recall %s" % (
            ratio, sec, n / sec, st["last_kernel_ms"], st["kernel_launches"], st["dist_evals"] / n, st["visited_resets"] / n,
            ["%.4f" % recall(g.searchKnnBatch(Q, 10, ef=ef)["labels"], gt) for ef in efs]), flush=True)
        del g

rng = np.random.default_rng(0)
which = sys.argv[1] if len(sys.argv) > 1 else "small"
if which == "small":
    X = rng.standard_normal((20000, 48), dtype=np.float32); X /= np.linalg.norm(X, axis=1, keepdims=True)
    run("ip-gauss", 1, X, rng.standard_normal((1000, 48), dtype=np.float32), 12, 80, [4, 8, 16, 32, 64], [32, 64, 128])
    X = rng.standard_normal((50000, 128), dtype=np.float32)
    run("l2-gauss", 0, X, rng.standard_normal((1000, 128), dtype=np.float32), 16, 200, [4, 8, 16, 32], [64, 128, 256])
else:
    n = int(which)
    X = bind.lowrank_data(n, 128, seed=1); Q = bind.lowrank_data(2000, 128, seed=2)
    for ramp in os.environ.get("PROBE_RAMPS", "32").split(","):
        os.environ["B200HNSW_BUILD_RAMP_RATIO"] = ramp
        print("== ramp ratio", ramp, flush=True)
        run("c2-lowrank", 0, X, Q, 32, 200, [int(r) for r in os.environ.get("PROBE_RATIOS", "32").split(",")], [16, 28, 32, 64])

#!/usr/bin/env python
"""GPU build of the C2 data set only (1M x 128, M=32, efc=200), twice; B200HNSW_LIB selects the library (A/B)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
X = lowrank_data(1_000_000, 128, seed=1)
for rep in range(2):
    t_c = time.time(); g = pkg.HierarchicalNSW(pkg.L2Space(128), len(X), 32, 200); t_c = time.time() - t_c
    t = time.time(); g.addPoints(X); t_add = time.time() - t; g.flush(); sec = time.time() - t
    print("constructor %.3f s" % t_c, flush=True)
    st = g.stats()
    print("%s rep %d: %.2f s (%.0f pts/s), addPoints %.2f s, flush events %.0f ms, D/pt %.1f, dropped reverse edges %d" % (
        os.path.basename(os.environ.get("B200HNSW_LIB", "head")), rep, sec, len(X) / sec, t_add, st["last_kernel_ms"],
        st["dist_evals"] / len(X), st["dropped_reverse_edges"]), flush=True)
    if rep == 1 and os.environ.get("PROBE_RECALL"):
        # recall@10 of the GPU-built graph at the bench's ef (reference-built graph: 0.950-0.952, same queries)
        Qr = lowrank_data(10000, 128, seed=2)
        bf = pkg.BruteforceSearch(pkg.L2Space(128), len(X)); bf.addPoints(X); gt = bf.searchKnnBatch(Qr, 10)["labels"]; del bf
        for ef in (28, 64):
            lab = g.searchKnnBatch(Qr, 10, ef=ef)["labels"]
            print("recall@10 of the built graph at ef=%d: %.4f" % (ef, np.mean([len(set(a) & set(b)) for a, b in zip(lab.tolist(), gt.tolist())]) / 10), flush=True)
    if rep == 1 and os.environ.get("PROBE_SEARCH_AFTER"):
        # the same amount of traversal as one full build batch, through the search kernel (for comparison)
        Qb = X[-16384:]
        for ef in (200,):
            for _ in range(3):
                r = g.searchKnnBatch(Qb, 10, ef=ef, work=True)
            s2 = g.stats()
            print("search kernel on the built graph: 16384 queries ef=%d: %.2f ms, D/q %.0f, H0/q %.1f, resets %d" % (
                ef, s2["last_kernel_ms"], s2["dist_evals"] / len(Qb), s2["hops_base"] / len(Qb), s2["visited_resets"]), flush=True)
    del g

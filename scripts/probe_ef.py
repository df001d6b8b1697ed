"""Ad-hoc: search time vs ef and visited-table size on the cached C2 graph (run bench.py once before)."""
import os, sys, glob
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
path = sorted(glob.glob("/tmp/b200hnsw_cache/l2_n1000000_*_r0.bin"))[0]
Q = lowrank_data(10000, 128, seed=2)
for ef in (32, 64, 128, 256):
    for hb in (0, 10, 11, 12, 13):
        if hb: os.environ["B200HNSW_HASH_BITS"] = str(hb)
        else: os.environ.pop("B200HNSW_HASH_BITS", None)
        idx = pkg.HierarchicalNSW(pkg.L2Space(128), path) if ef == 32 and hb == 0 else idx
        for _ in range(3): r = idx.searchKnnBatch(Q, 10, ef=ef, work=True)
        st = idx.stats()
        print("ef %3d hash_bits %2d: %.3f ms  D/q %.0f resets/q %.2f" % (ef, hb, st["last_kernel_ms"], st["dist_evals"] / 1e4, st["visited_resets"] / 1e4), flush=True)

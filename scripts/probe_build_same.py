"""Build a graph and print the sha256 of its saveIndex file (B200HNSW_LIB selects the library): two library builds whose
construction kernels make the same decisions must print the same hash.  usage: probe_build_same.py [n d M efc metric]"""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
n, d, M, efc = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (60000, 128, 16, 100)))
metric = sys.argv[5] if len(sys.argv) > 5 else "l2"
X = lowrank_data(n, d, seed=5)
space = pkg.L2Space(d) if metric == "l2" else pkg.InnerProductSpace(d)
g = pkg.HierarchicalNSW(space, n, M, efc)
g.addPoints(X[: n // 2]); g.flush(); g.addPoints(X[n // 2:]); g.flush()
path = "/tmp/probe_same_%d.bin" % os.getpid()
g.saveIndex(path)
print("%s n=%d d=%d M=%d efc=%d %s: sha256 %s, D/pt %.1f" % (os.path.basename(os.environ.get("B200HNSW_LIB", "head")), n, d, M, efc,
      metric, hashlib.sha256(open(path, "rb").read()).hexdigest()[:16], g.stats()["dist_evals"] / (n - n // 2)), flush=True)
os.remove(path)

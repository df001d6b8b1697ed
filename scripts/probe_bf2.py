import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
n, d, k, nq = int(sys.argv[1]), 768, 100, int(sys.argv[2])
X = lowrank_data(n, d, seed=1, latent=64, noise=0.1, normalize=True)
Q = lowrank_data(nq, d, seed=2, latent=64, noise=0.1, normalize=True)
g = pkg.BruteforceSearch(pkg.InnerProductSpace(d), n); g.addPoints(X)
os.environ["B200HNSW_BF_PATH"] = "tensor"
for _ in range(3): r = g.searchKnnBatch(Q, k)
print("ok", g.stats()["last_kernel_ms"])

"""C4 stage timing probe: B200HNSW_BF_PROFILE=1 python scripts/probe_bf_c4.py [nq ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
n, d, k = 1_000_000, 768, 100
nqs = [int(x) for x in sys.argv[1:]] or [10000]
X = lowrank_data(n, d, seed=11, latent=64, noise=0.1, normalize=True)
Q = lowrank_data(max(nqs), d, seed=12, latent=64, noise=0.1, normalize=True)
g = pkg.BruteforceSearch(pkg.InnerProductSpace(d), n)
g.addPoints(X)
dev = torch.device("cuda", 0)
dQ = torch.from_numpy(Q).to(dev)
for nq in nqs:
    ol = torch.empty((nq, k), dtype=torch.int64, device=dev); od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for _ in range(3):
        g.searchKnnDevice(dQ.data_ptr(), nq, k, ol.data_ptr(), od.data_ptr(), 0, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.searchKnnDevice(dQ.data_ptr(), nq, k, ol.data_ptr(), od.data_ptr(), 0, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("nq %d: %.3f ms/batch, %.1f kQPS, %.0f TFLOP/s algorithmic" % (nq, ms, nq / ms, 2.0 * nq * n * d / ms / 1e9), flush=True)
os.environ["B200HNSW_BF_PATH"] = "scan"
r0 = g.searchKnnBatch(Q[:256], k)
os.environ["B200HNSW_BF_PATH"] = "tensor"
r1 = g.searchKnnBatch(Q[:256], k)
print("tensor == scan on 256 queries:", bool(np.array_equal(r0["labels"], r1["labels"]) and np.array_equal(r0["dists"], r1["dists"])))

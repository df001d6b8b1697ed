"""C4 small-batch probe: every dispatch path at 1..128 queries (1M x 768 inner product, k=100)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data
n, d, k = 1_000_000, 768, 100
X = lowrank_data(n, d, seed=11, latent=64, noise=0.1, normalize=True)
Q = lowrank_data(256, d, seed=12, latent=64, noise=0.1, normalize=True)
g = pkg.BruteforceSearch(pkg.InnerProductSpace(d), n)
g.addPoints(X)
dev = torch.device("cuda", 0)
dQ = torch.from_numpy(Q).to(dev)
ol = torch.empty((256, k), dtype=torch.int64, device=dev); od = torch.empty((256, k), dtype=torch.float32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for nq in (1, 2, 4, 8, 16, 32, 64, 128):
    for path in ("auto", "stream", "tensor"):
        if path == "stream" and nq > 64: continue
        if path == "auto": os.environ.pop("B200HNSW_BF_PATH", None)
        else: os.environ["B200HNSW_BF_PATH"] = path
        for _ in range(3): g.searchKnnDevice(dQ.data_ptr(), nq, k, ol.data_ptr(), od.data_ptr(), 0, 0)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): g.searchKnnDevice(dQ.data_ptr(), nq, k, ol.data_ptr(), od.data_ptr(), 0, 0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("nq %3d %-6s %.4f ms  %.1f QPS  one-pass-fp32 frac %.3f" % (nq, path, ms, nq / ms * 1e3, 4.0 * n * d / (ms * 1e-3) / 6550.1e9), flush=True)

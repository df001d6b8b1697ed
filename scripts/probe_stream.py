#!/usr/bin/env python
"""nq-query streaming brute force on 1M x 768 (for ncu): python scripts/probe_stream.py [nq] [n] [dim]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import research_new_hnsw_b200 as pkg
import torch
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 768
rng = np.random.default_rng(1)
X = rng.standard_normal((n, dim), dtype=np.float32)
X /= np.linalg.norm(X, axis=1, keepdims=True)
g = pkg.BruteforceSearch(pkg.InnerProductSpace(dim), n)
g.addPoints(X)
dev = torch.device("cuda", 0)
dQ = torch.from_numpy(X[:64].copy()).to(dev)
ol = torch.empty((64, 100), dtype=torch.int64, device=dev); od = torch.empty((64, 100), dtype=torch.float32, device=dev)
os.environ["B200HNSW_BF_PATH"] = "stream"
st = torch.cuda.current_stream().cuda_stream
for _ in range(3): g.searchKnnDevice(dQ.data_ptr(), nq, 100, ol.data_ptr(), od.data_ptr(), 0, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g.searchKnnDevice(dQ.data_ptr(), nq, 100, ol.data_ptr(), od.data_ptr(), 0, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("nq %d n %d dim %d: %.3f ms, %.0f GB/s one pass" % (nq, n, dim, ms, 4.0 * n * dim / ms / 1e6))

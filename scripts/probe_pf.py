"""A/B probe of the search kernel's run-time policies (B200HNSW_PF / B200HNSW_PDL / B200HNSW_TEAM ... are read once per
process, so every variant is its own process).  Graph + ground truth are cached in /tmp for the lifetime of the box.
usage: probe_pf.py [ef ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data

n, d, M, efc, nq, k = int(os.environ.get("PROBE_N", 1000000)), 128, 32, 200, 10000, 10
efs = [int(x) for x in sys.argv[1:]] or [28]
path, gtp = "/tmp/probe_c2_%d.bin" % n, "/tmp/probe_c2_%d_gt.npy" % n
batches = [lowrank_data(nq, d, seed=2 + 7 * b) for b in range(4)]
storage = 1 if os.environ.get("PROBE_BF16") else 0
if not os.path.exists(path):
    X = lowrank_data(n, d, seed=1)
    t = time.time()
    g = pkg.HierarchicalNSW(pkg.L2Space(d), n, M, efc)
    g.addPoints(X)
    g.flush()
    g.saveIndex(path)
    print("gpu build+save %.1fs" % (time.time() - t), flush=True)
    bf = pkg.BruteforceSearch(pkg.L2Space(d), n)
    bf.addPoints(X)
    np.save(gtp, bf.searchKnnBatch(batches[0][:1000], k)["labels"])
    del bf, g, X
gt = np.load(gtp)
idx = pkg.HierarchicalNSW(pkg.L2Space(d), path, storage=storage)
dev = torch.device("cuda", 0)
dq = [torch.from_numpy(b).to(dev) for b in batches]
ol = torch.empty((nq, k), dtype=torch.int64, device=dev)
od = torch.empty((nq, k), dtype=torch.float32, device=dev)
w = torch.zeros((nq, 4), dtype=torch.int32, device=dev)
if not os.environ.get("PROBE_NULLSTREAM"):
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
st = torch.cuda.current_stream().cuda_stream
tag = " ".join("%s=%s" % (v, os.environ[v]) for v in sorted(os.environ) if v.startswith("B200HNSW_") or v.startswith("PROBE_"))
for ef in efs:
    idx.searchKnnDevice(dq[0].data_ptr(), nq, k, ef, ol.data_ptr(), od.data_ptr(), 0, w.data_ptr(), st)
    torch.cuda.synchronize()
    lab = ol.cpu().numpy().view(np.uint64)[:1000]
    rec = np.mean([len(set(a) & set(b)) for a, b in zip(lab.tolist(), gt.tolist())]) / k
    wk = w.cpu().numpy().astype(np.int64)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for s in range(5):
            idx.searchKnnDevice(dq[s % 4].data_ptr(), nq, k, ef, ol.data_ptr(), od.data_ptr(), 0, 0, st)
        torch.cuda.synchronize()
        e0.record()
        evs = []
        for s in range(20):
            if os.environ.get("PROBE_WAIT") and len(evs) >= 2:   # a wait on an event that completed long ago
                torch.cuda.current_stream().wait_event(evs[-2])
            idx.searchKnnDevice(dq[s % 4].data_ptr(), nq, k, ef, ol.data_ptr(), od.data_ptr(), 0, 0, st)
            if os.environ.get("PROBE_EVENTS"):                    # an event record between consecutive launches
                ev = torch.cuda.Event(); ev.record(); evs.append(ev)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    B = wk[:, 0].sum() * d * (2 if storage else 4) + wk[:, 1].sum() * (4 + 8 * M) + wk[:, 2].sum() * (4 + 4 * M) + nq * (4 * d + 12 * k)
    print("[%s] ef %3d  %.4f ms/step  %.2f MQPS  recall %.4f  D/q %.0f H0/q %.1f resets %d  alg %.0f GB/s"
          % (tag, ef, best, nq / best / 1e3, rec, wk[:, 0].mean(), wk[:, 1].mean(), wk[:, 3].sum(), B / best / 1e6), flush=True)

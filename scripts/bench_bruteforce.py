#!/usr/bin/env python
"""C4 (BASELINE.json configs[3]): BruteforceSearch exact k=100 on 1M x 768 inner product -- tcgen05 GEMM candidates +
exact re-rank, next to the exact-scan kernel and the unmodified reference on the host cores.  Prints one JSON line
(same spirit as bench.py; the headline metric of the repo is bench.py's)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import research_new_hnsw_b200 as pkg
from research_new_hnsw_b200.synth import lowrank_data

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--nq", type=int, default=10_000); ap.add_argument("--k", type=int, default=100)
ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
import torch
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if WORLD > 1:
    # SURVEY.md 8(e), brute force: rows sharded across GPUs (strong scaling: the SAME 1M rows, 1/N per GPU), every GPU
    # scans its rows for the whole query batch, one packed all_gather of the per-shard top-k + GPU merge.
    import torch.distributed as dist
    from research_new_hnsw_b200.sharded import PackedShardExchange, shard_range
    rank, lrank = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    dev = torch.device("cuda", lrank)
    X = lowrank_data(a.n, a.dim, seed=1, latent=64, noise=0.1, normalize=True)
    Qs = [lowrank_data(a.nq, a.dim, seed=2 + 7 * b, latent=64, noise=0.1, normalize=True) for b in range(2)]
    lo, hi = shard_range(a.n, rank, WORLD)
    g = pkg.BruteforceSearch(pkg.InnerProductSpace(a.dim), hi - lo, device=lrank)
    g.addPoints(X[lo:hi], np.arange(lo, hi, dtype=np.uint64))
    dQ = [torch.from_numpy(q).to(dev) for q in Qs]
    ex = PackedShardExchange(a.nq, a.k, dev)
    stream = torch.cuda.current_stream().cuda_stream
    def step(s):
        pl, pd = ex.local_ptrs()
        g.searchKnnDevice(dQ[s % 2].data_ptr(), a.nq, a.k, pl, pd, 0, stream)
        return ex.exchange_and_merge(stream)
    for s in range(a.warmup): step(s)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps): ol, od = step(s)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / a.steps
    if rank == 0:
        full = pkg.BruteforceSearch(pkg.InnerProductSpace(a.dim), a.n, device=lrank)
        full.addPoints(X)
        ref = full.searchKnnBatch(Qs[(a.steps - 1) % 2][:256], a.k)
        same = bool(np.array_equal(ref["labels"], ol[:256].cpu().numpy().view(np.uint64))
                    and np.array_equal(ref["dists"], od[:256].cpu().numpy()))
        print(json.dumps({"metric": "BruteforceSearch exact k=%d QPS, %dx%d inner product, rows sharded over %d GPUs"
                          % (a.k, a.n, a.dim, WORLD), "value": a.nq / (ms * 1e-3), "unit": "queries/s", "n_gpus": WORLD,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "strong", "data": "synthetic",
                          "config": {"workload": "C4 rows split %d ways, %d queries per batch, one packed all_gather + merge"
                                     % (WORLD, a.nq), "merged_equals_single_gpu_256_queries": same}}))
    dist.barrier()
    sys.exit(0)
X = lowrank_data(a.n, a.dim, seed=1, latent=64, noise=0.1, normalize=True)
Qs = [lowrank_data(a.nq, a.dim, seed=2 + 7 * b, latent=64, noise=0.1, normalize=True) for b in range(2)]
g = pkg.BruteforceSearch(pkg.InnerProductSpace(a.dim), a.n)
g.addPoints(X)
dev = torch.device("cuda", 0)
dQ = [torch.from_numpy(q).to(dev) for q in Qs]
ol = torch.empty((a.nq, a.k), dtype=torch.int64, device=dev); od = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
os.environ["B200HNSW_BF_PATH"] = "tensor"
for s in range(a.warmup): g.searchKnnDevice(dQ[s % 2].data_ptr(), a.nq, a.k, ol.data_ptr(), od.data_ptr(), 0, stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(a.steps): g.searchKnnDevice(dQ[s % 2].data_ptr(), a.nq, a.k, ol.data_ptr(), od.data_ptr(), 0, stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
hq = [torch.from_numpy(q).pin_memory() for q in Qs]
for s in range(3): g.searchKnnBatch(hq[s % 2].numpy(), a.k)
t0 = time.perf_counter()
for s in range(a.steps): r = g.searchKnnBatch(hq[s % 2].numpy(), a.k)
e2e = (time.perf_counter() - t0) / a.steps
# parity on a sample: exact scan kernel + reference as shipped (SSE)
os.environ["B200HNSW_BF_PATH"] = "scan"
rs = g.searchKnnBatch(Qs[(a.steps - 1) % 2][:512], a.k)
scan_ok = bool(np.array_equal(rs["labels"], r["labels"][:512]) and np.array_equal(rs["dists"], r["dists"][:512]))
cpu = None
try:
    from oracle import bind
    ref = bind.Ref("sse"); b = ref.bf_new(bind.IP, a.dim, a.n); b.add(X)
    T = os.cpu_count() or 1
    rr = b.search(Qs[(a.steps - 1) % 2][:64], a.k, threads=T)
    cpu = {"value": 64 / rr["seconds"], "unit": "queries/s", "cores": T, "kind": "reference",
           "sample": "64 queries, -O3 SSE build as shipped, %d threads" % T,
           "ids_bit_exact": bool(np.array_equal(rr["labels"], r["labels"][:64])),
           "dists_bit_equal": bool(np.array_equal(rr["dists"], r["dists"][:64]))}
except Exception as e:
    cpu = {"value": None, "sample": "unavailable: %s" % e}
# small-batch regime (nq = 1 and 128 next to the headline): one pass over the rows is the floor
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6550.0)
small = []
for nq_s in (1, 4, 8, 16, 32, 64, 128, 1024):
    for path in ("auto", "stream", "tensor"):
        if path == "stream" and nq_s > 128: continue
        if path == "tensor" and nq_s < 16: continue
        if path == "auto": os.environ.pop("B200HNSW_BF_PATH", None)
        else: os.environ["B200HNSW_BF_PATH"] = path
        for _ in range(2): g.searchKnnDevice(dQ[0].data_ptr(), nq_s, a.k, ol.data_ptr(), od.data_ptr(), 0, stream)
        torch.cuda.synchronize()
        e0.record()
        for it in range(5): g.searchKnnDevice(dQ[it % 2].data_ptr(), nq_s, a.k, ol.data_ptr(), od.data_ptr(), 0, stream)
        e1.record(); torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1) / 5
        small.append({"nq": nq_s, "path": path, "ms": round(ms_s, 4), "qps": round(nq_s / (ms_s * 1e-3), 1),
                      "hbm_frac_one_pass": round(4.0 * a.n * a.dim / (ms_s * 1e-3) / 1e9 / hbm, 3)})
os.environ.pop("B200HNSW_BF_PATH", None)
# exactness of the small-batch paths against the tiled exact scan
rq = g.searchKnnBatch(Qs[0][:16], a.k)
os.environ["B200HNSW_BF_PATH"] = "scan"
rq2 = g.searchKnnBatch(Qs[0][:16], a.k)
os.environ.pop("B200HNSW_BF_PATH", None)
small_ok = bool(np.array_equal(rq["labels"], rq2["labels"]) and np.array_equal(rq["dists"], rq2["dists"]))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
flops = 2.0 * a.nq * a.n * a.dim
print(json.dumps({"metric": "BruteforceSearch exact k=%d QPS, %dx%d inner product" % (a.k, a.n, a.dim), "value": a.nq / (ms * 1e-3),
    "unit": "queries/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
    "dtype": "bf16 candidates + f32 exact re-rank", "data": "synthetic",
    "config": {"workload": "C4: %dx%d unit-norm rank-64+noise rows, IP, k=%d, %d queries per batch" % (a.n, a.dim, a.k, a.nq),
               "exact_scan_parity_512_queries": scan_ok, "small_batches": small, "small_batch_parity_16_queries": small_ok},
    "e2e": {"value": a.nq / e2e, "unit": "queries/s", "h2d_bytes_per_step": a.nq * a.dim * 4, "d2h_bytes_per_step": a.nq * a.k * 12 + a.nq * 4},
    "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                 "frac": flops / (ms * 1e-3) / 1e12 / peak, "traffic": None,
                 "note": "algorithmic 2*nq*N*d over the whole pipeline (sampled bound pass + candidate GEMM + re-rank)"},
    "cpu_baseline": cpu}))

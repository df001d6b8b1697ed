#!/usr/bin/env python
"""bench.py -- headline benchmark: QPS @ recall@10 >= 0.95 of batched searchKnn (BASELINE.json metric).

Workload (BASELINE.json configs[1], "C2"): synthetic SIFT-shaped 1M x 128 fp32, L2, M=32, ef_construction=200,
batches of 10 000 queries, k=10; ef = the smallest value of the sweep whose recall@10 (exact ground truth, first
1000 queries) is >= 0.95.  A "step" is one batch of 10 000 queries through the hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm  (CUDA kernels through the C ABI)
  python bench.py --impl reference ...                           reference arm (unmodified hnswlib on host cores)

N > 1 (torchrun, one rank per GPU): the data set is sharded, one 1M-point sub-index per GPU (weak scaling: N x 1M
points in total), every rank searches the same query batch, per-shard top-k are exchanged with an NCCL all_gather and
merged on the GPU.  The unit of work that is fixed per GPU is one (query, shard) search, so at N > 1
`value` is in "shard-searches/s" (N x nq per step) and `merged_qps` = value / N is the rate of merged answers over the
whole N x 1M data set.  The reference arm at N searches the same N shards on the host cores (its own graphs, its own
recall-driven ef) and reports the same unit.  The default N = 1 run also measures C4 (BruteforceSearch 1M x 768, tensor
roofline) and C5 (GPU graph build, HBM roofline) into `config.c4` / `config.build`.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EF_SWEEP = [10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 48, 56, 64, 96, 128, 192, 256]
CACHE = os.environ.get("B200HNSW_CACHE", "/tmp/b200hnsw_cache")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--points", dest="n", type=int, default=1_000_000, help="points per GPU (shard size)")
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--M", type=int, default=32)
    ap.add_argument("--efc", type=int, default=200)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ef", type=int, default=0, help="override the recall-driven ef selection")
    ap.add_argument("--recall", type=float, default=0.95)
    ap.add_argument("--batches", type=int, default=4, help="distinct query batches cycled through the steps")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--metric", default="l2", choices=["l2", "ip"],
                    help="l2 = C2 (SIFT-shaped); ip = C3-style shards (Deep-shaped: unit-norm rows, inner product)")
    ap.add_argument("--parallel", default="shard", choices=["shard", "replica"],
                    help="N > 1: 'shard' = one sub-index per GPU + all_gather/merge (north_star); 'replica' = the same "
                         "index on every GPU, each GPU serving its own query batches (no exchange)")
    ap.add_argument("--storage", default="f32", choices=["f32", "bf16"],
                    help="device copy of the vectors: f32 (default, reference-exact traversal) or bf16 traversal + f32 re-rank")
    return ap.parse_args()


def recall_at_k(labels, gt):
    return float(np.mean([len(set(a) & set(b)) for a, b in zip(labels.tolist(), gt.tolist())]) / gt.shape[1])


def workload_name(a):
    if a.metric == "ip":
        return ("C3-shaped shard: synthetic Deep-shaped (rank-16 + 0.1 noise, unit-norm) %dx%d fp32 inner product, M=%d "
                "ef_construction=%d, batched searchKnn %d queries k=%d" % (a.n, a.dim, a.M, a.efc, a.nq, a.k))
    return ("C2: synthetic SIFT-shaped (rank-16 + 0.1 noise) %dx%d fp32 L2, M=%d ef_construction=%d, batched searchKnn "
            "%d queries k=%d" % (a.n, a.dim, a.M, a.efc, a.nq, a.k))


def shard_data(a, rank):
    from research_new_hnsw_b200.synth import lowrank_data
    return lowrank_data(a.n, a.dim, seed=1 + 1000 * rank, normalize=(a.metric == "ip"))


def query_batches(a):
    from research_new_hnsw_b200.synth import lowrank_data
    return [lowrank_data(a.nq, a.dim, seed=2 + 7 * b, normalize=(a.metric == "ip")) for b in range(max(1, a.batches))]


def metric_name(a):
    if a.metric == "ip":
        return "QPS @ recall@10>=0.95, %dx%d inner-product shard per GPU" % (a.n, a.dim)
    return "QPS @ recall@10>=0.95, 1Mx128 L2"


def space_of(pkg, a):
    return pkg.InnerProductSpace(a.dim) if a.metric == "ip" else pkg.L2Space(a.dim)


def ref_metric(a):
    from oracle import bind
    return bind.IP if a.metric == "ip" else bind.L2


def graph_path(a, rank):
    os.makedirs(CACHE, exist_ok=True)
    return os.path.join(CACHE, "%s_n%d_d%d_M%d_efc%d_r%d.bin" % (a.metric, a.n, a.dim, a.M, a.efc, rank))


def build_graph_with_reference(a, rank, X, threads):
    """The graph both arms search is the reference's own (multi-threaded addPoint + saveIndex); cached on disk so the
    two arms of one driver run share it."""
    from oracle import bind
    path = graph_path(a, rank)
    if os.path.exists(path) and os.path.getsize(path) > 96 + a.n * (a.dim * 4 + 8 * a.M + 12):
        return path, 0.0
    ref = bind.Ref(bind.best_ref_level())
    idx = ref.hnsw_new(ref_metric(a), a.dim, a.n, a.M, a.efc)
    labels = np.arange(a.n, dtype=np.uint64) + np.uint64(rank * a.n)
    sec = idx.add(X, labels, threads=threads)
    tmp = path + ".tmp%d" % os.getpid()
    idx.save(tmp)
    os.replace(tmp, path)
    return path, sec


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (nvidia-smi's clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def algorithmic_bytes(a, work, ef=0):
    """B_q = D*d*s + H0*(4+4*maxM0) + Hup*(4+4*maxM) + 4d + 12k, summed over the batch (counted).
    s = 4 (f32 rows); with bf16 storage the traversal reads 2-byte components and the final ef entries are re-read
    as f32 rows (the kernel's D counts both)."""
    D, H0, Hup = int(work[:, 0].sum()), int(work[:, 1].sum()), int(work[:, 2].sum())
    nq = work.shape[0]
    if a.storage == "bf16":
        rer = min(ef, D // max(nq, 1)) * nq
        vec_bytes = (D - rer) * a.dim * 2 + rer * a.dim * 4
    else:
        vec_bytes = D * a.dim * 4
    return vec_bytes + H0 * (4 + 8 * a.M) + Hup * (4 + 4 * a.M) + nq * (4 * a.dim + 12 * a.k), D, H0, Hup


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_leg(a, path, batches, ef, budget_s, threads):
    """Unmodified reference (oracle/_ref) on the host cores: the hnsw_service pattern, T threads looping searchKnn."""
    from oracle import bind
    level = bind.best_ref_level()
    ref = bind.Ref(level)
    idx = ref.hnsw_load(ref_metric(a), a.dim, path)
    idx.search(batches[0][:2000], a.k, ef, threads=threads)  # warm-up: per-thread VisitedList allocation
    done, sec, i = 0, 0.0, 0
    while sec < budget_s and i < 64:
        r = idx.search(batches[i % len(batches)], a.k, ef, threads=threads)
        sec += r["seconds"]
        done += a.nq
        i += 1
    return dict(value=done / sec, unit="queries/s", cores=threads, kind="reference",
                sample="%d batches of %d queries, ef=%d, reference built -O3 %s, %d threads looping searchKnn"
                       % (i, a.nq, ef, level, threads)), idx


def merge_rows_numpy(L, D, k):
    """k smallest (dist, label) pairs per row of [nq][shards*k] candidates (the merge kernel's contract), vectorised"""
    order = np.lexsort((L, D), axis=1)[:, :k]
    return np.take_along_axis(L, order, 1), np.take_along_axis(D, order, 1)


def reference_exact_topk(a, X, labels, Q, threads):
    """exact top-k of one shard by the UNMODIFIED reference's BruteforceSearch (as shipped, SSE)"""
    from oracle import bind
    ref = bind.Ref("sse")
    bf = ref.bf_new(ref_metric(a), a.dim, X.shape[0])
    bf.add(X, labels)
    r = bf.search(Q, a.k, threads=threads)
    return r["labels"], r["dists"]


def pick_ef(a, search_fn, gt1000, Qs):
    if a.ef:
        r = search_fn(Qs, a.ef)
        return a.ef, recall_at_k(r, gt1000), {a.ef: recall_at_k(r, gt1000)}
    table = {}
    for ef in EF_SWEEP:
        if ef < a.k:
            continue
        rec = recall_at_k(search_fn(Qs, ef), gt1000)
        table[ef] = round(rec, 4)
        if rec >= a.recall:
            return ef, rec, table
    return EF_SWEEP[-1], rec, table


def c4_leg(pkg, local_rank, steps, warmup):
    """BASELINE.json configs[3] (C4): BruteforceSearch exact k=100 on 1M x 768 inner product -- tcgen05 GEMM candidates
    + exact re-rank (csrc/bf_tensor.cu), 10 000 queries per batch; tensor roofline F = 2*nq*N*d against the measured
    bf16 peak; ids / distances compared with the unmodified reference (as shipped, SSE) on a query sample."""
    import torch
    from research_new_hnsw_b200.synth import lowrank_data
    n, d, k, nq = 1_000_000, 768, 100, 10_000
    X = lowrank_data(n, d, seed=11, latent=64, noise=0.1, normalize=True)
    Qs = [lowrank_data(nq, d, seed=12 + b, latent=64, noise=0.1, normalize=True) for b in range(2)]
    g = pkg.BruteforceSearch(pkg.InnerProductSpace(d), n, device=local_rank)
    g.addPoints(X)
    dev = torch.device("cuda", local_rank)
    dQ = [torch.from_numpy(q).to(dev) for q in Qs]
    ol = torch.empty((nq, k), dtype=torch.int64, device=dev)
    od = torch.empty((nq, k), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for s in range(warmup):
        g.searchKnnDevice(dQ[s % 2].data_ptr(), nq, k, ol.data_ptr(), od.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        g.searchKnnDevice(dQ[s % 2].data_ptr(), nq, k, ol.data_ptr(), od.data_ptr(), 0, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    paths = {0: "scan", 1: "tensor", 2: "stream"}
    hq = [torch.from_numpy(q).pin_memory() for q in Qs]
    for s in range(2):
        g.searchKnnBatch(hq[s % 2].numpy(), k)
    t0 = time.perf_counter()
    for s in range(steps):
        r = g.searchKnnBatch(hq[s % 2].numpy(), k)
    e2e = (time.perf_counter() - t0) / steps
    path_taken = paths.get(g.stats()["hops_base"], "?")  # the dispatcher's choice (recorded by the host-pointer call)
    Qlast = Qs[(steps - 1) % 2]
    small = []
    for nq_s in (1, 8, 128):  # the bandwidth-bound regime: one pass over the fp32 rows is the floor
        for _ in range(2):
            g.searchKnnDevice(dQ[0].data_ptr(), nq_s, k, ol.data_ptr(), od.data_ptr(), 0, stream)
        torch.cuda.synchronize()
        e0.record()
        for it in range(5):
            g.searchKnnDevice(dQ[it % 2].data_ptr(), nq_s, k, ol.data_ptr(), od.data_ptr(), 0, stream)
        e1.record()
        torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1) / 5
        g.searchKnnBatch(Qs[0][:nq_s], k)
        small.append({"nq": nq_s, "ms": round(ms_s, 4), "qps": round(nq_s / (ms_s * 1e-3), 1),
                      "path": paths.get(g.stats()["hops_base"], "?"),
                      "hbm_frac_one_pass": round(4.0 * n * d / (ms_s * 1e-3) / 1e9 / peaks()[0], 3)})
    cpu, exact = None, None
    try:
        from oracle import bind
        T = os.cpu_count() or 1
        b = bind.Ref("sse").bf_new(bind.IP, d, n)
        b.add(X)
        rr = b.search(Qlast[:64], k, threads=T)
        exact = {"queries": 64, "ids_bit_exact": bool(np.array_equal(rr["labels"], r["labels"][:64])),
                 "dists_bit_identical": bool(np.array_equal(rr["dists"], r["dists"][:64]))}
        del b
        lvl = bind.best_ref_level()
        b = bind.Ref(lvl).bf_new(bind.IP, d, n)
        b.add(X)
        rb = b.search(Qlast[:128], k, threads=T)
        cpu = {"value": 128 / rb["seconds"], "unit": "queries/s", "cores": T, "kind": "reference",
               "sample": "128 queries, BruteforceSearch::searchKnn over %d threads, -O3 %s build of the unmodified headers"
                         % (T, lvl)}
    except Exception as e:  # the checker binary is missing: report it, never substitute
        cpu = {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_t = float(json.load(open(pj))["bf16_tflops"]) if os.path.exists(pj) else 1590.0
    # cuBLAS under the power cap (seconds-long loop); the candidate GEMM runs power-capped as well (ncu: 1.34 GHz)
    peak_s = float(json.load(open(pj)).get("bf16_tflops_sustained", 0.0)) if os.path.exists(pj) else 1400.0
    flops = 2.0 * nq * n * d
    return {"workload": "C4: BruteforceSearch exact k=%d, %dx%d unit-norm rank-64+noise rows, inner product, %d queries "
                        "per batch" % (k, n, d, nq),
            "qps": nq / (ms * 1e-3), "ms_per_step": ms, "path": path_taken,
            "e2e": {"value": nq / e2e, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4,
                    "d2h_bytes_per_step": nq * k * 12 + nq * 4, "api": "b200bf_search_batch (host pointers, pinned)"},
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak_t, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / peak_t, "traffic": None,
                         "peak_sustained": peak_s or None,
                         "frac_of_sustained": (flops / (ms * 1e-3) / 1e12 / peak_s) if peak_s else None,
                         "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops)" if os.path.exists(pj) else "fallback",
                         "note": "algorithmic 2*nq*N*d over the WHOLE pipeline (candidate GEMM + exact fp32 re-rank)"},
            "small_batches": small, "parity_vs_reference": exact, "cpu_baseline": cpu}


def run_reference(a, rank, world):
    """The unmodified reference on the host cores, same workload and unit as our arm: at N = 1 one 1M-point index; at
    N > 1 the same N shards (seeds of shard_data) built by its own multi-threaded addPoint, every query searched on
    every shard (T threads looping searchKnn, the hnsw_service pattern) and the per-shard top-k merged on the CPU."""
    if rank != 0:
        return
    from oracle import bind
    if not bind.ref_available("sse"):
        # the compiled reference (oracle/_ref, built from /root/reference in the build container) did not travel
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhnswref_sse.so is missing on this box"}),
              flush=True)
        return
    threads = os.cpu_count() or 1
    batches = query_batches(a)
    Qs = batches[0][:1000]
    ref = bind.Ref(bind.best_ref_level())
    budget = float(os.environ.get("B200HNSW_REF_BUILD_BUDGET_S", "480"))
    shards, gts, build_s, build_note = [], [], 0.0, []
    for r in range(world):
        X = shard_data(a, r)
        labels = np.arange(a.n, dtype=np.uint64) + np.uint64(r * a.n)
        path, sec = build_graph_with_reference(a, r, X, threads)
        build_s += sec
        build_note.append(sec)
        bf = ref.bf_new(ref_metric(a), a.dim, a.n)
        bf.add(X, labels)
        g = bf.search(Qs, a.k, threads=threads)
        gts.append((g["labels"], g["dists"]))
        del bf, X
        shards.append(ref.hnsw_load(ref_metric(a), a.dim, path))
        if r + 1 < world and build_s > 0 and build_s / (r + 1) * world > budget:
            break  # a box too slow to build all N shards in time: measure what was built and say so
    S = len(shards)
    gt, _ = merge_rows_numpy(np.concatenate([g[0] for g in gts], 1), np.concatenate([g[1] for g in gts], 1), a.k)

    def search(Q, ef):
        rs = [idx.search(Q, a.k, ef, threads=threads) for idx in shards]
        sec = sum(r["seconds"] for r in rs)
        if S == 1:
            return rs[0]["labels"], sec
        t0 = time.perf_counter()
        L, _ = merge_rows_numpy(np.concatenate([r["labels"] for r in rs], 1), np.concatenate([r["dists"] for r in rs], 1), a.k)
        return L, sec + time.perf_counter() - t0

    ef, rec, table = pick_ef(a, lambda Q, e: search(Q, e)[0], gt, Qs)
    ef_table = []
    if S == 1:
        for e in (16, 32, 48, 64, 96, 128, 192, 256):
            L, sec = search(batches[0][:2000], e)
            ef_table.append({"ef": e, "recall_at_10": round(recall_at_k(L[:1000], gt), 4), "qps": 2000 / sec})
    for _ in range(a.warmup):
        search(batches[0][:2000], ef)
    sec = 0.0
    # each step = a bounded sample of the batch so the whole run stays within minutes even on few cores
    sample = min(a.nq, 10_000 if S == 1 else max(1000, 10_000 // S))
    for s in range(a.steps):
        sec += search(batches[s % len(batches)][:sample], ef)[1]
    rate = S * a.steps * sample / sec
    unit = "queries/s" if world == 1 else "shard-searches/s"
    line = {"impl": "reference", "metric": metric_name(a), "value": rate, "unit": unit,
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sec / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "merged_qps": rate / S, "total_points": S * a.n, "same_workload": S == world,
            "config": {"workload": workload_name(a), "ef": ef, "recall_at_10": round(rec, 4), "recall_sweep": table,
                       "ef_table": ef_table, "shards": S,
                       "graph": "built by the reference (addPoint, %d threads)%s" %
                                (threads, "" if build_s == 0 else " in %.1f s = %.0f points/s" % (build_s, S * a.n / build_s)),
                       "step": "%d queries of the batch%s" % (sample, "" if S == 1 else
                               ", each searched on all %d shards (%d threads per shard pass) + numpy top-k merge" % (S, threads))},
            "cpu_baseline": {"value": rate, "unit": unit, "cores": threads, "kind": "reference",
                             "sample": "%d queries per step, ef=%d, -O3 %s build of the unmodified headers"
                                       % (sample, ef, bind.best_ref_level())},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_b200(a, rank, local_rank, world):
    import torch
    import research_new_hnsw_b200 as pkg

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    threads = max(1, (os.cpu_count() or 1) // world)
    replica = world > 1 and a.parallel == "replica"
    sw = 1 if replica else world                    # ranks that exchange results (1: no data-path collective)
    drank = 0 if replica else rank
    X = shard_data(a, drank)
    if replica:  # every replica serves its own queries
        from research_new_hnsw_b200.synth import lowrank_data
        batches = [lowrank_data(a.nq, a.dim, seed=2 + 7 * b + 1000 * rank, normalize=(a.metric == "ip"))
                   for b in range(max(1, a.batches))]
    else:
        batches = query_batches(a)
    shard_labels = np.arange(a.n, dtype=np.uint64) + np.uint64(drank * a.n)
    if world == 1:
        # the graph both arms search: built by the reference on the host cores, loaded from its saveIndex file
        path, build_s = build_graph_with_reference(a, rank, X, threads)
        t0 = time.time()
        idx = pkg.HierarchicalNSW(space_of(pkg, a), path, device=local_rank, storage=1 if a.storage == "bf16" else 0)
        load_s = time.time() - t0
        graph_note = "reference-built saveIndex file%s, loaded in %.1f s" % (
            "" if build_s == 0 else " (%.1f s, %.0f points/s on %d threads)" % (build_s, a.n / build_s, threads), load_s)
    else:
        # one sub-index per GPU, built on that GPU (batched addPoint, csrc/build.cu)
        path = None
        t0 = time.time()
        idx = pkg.HierarchicalNSW(space_of(pkg, a), a.n, a.M, a.efc, device=local_rank,
                                  storage=1 if a.storage == "bf16" else 0)
        idx.addPoints(X, shard_labels)
        idx.flush()
        build_s = time.time() - t0
        graph_note = "per-GPU shard graph built on the GPU in %.1f s (%.0f points/s)" % (build_s, a.n / build_s)

    # exact ground truth for the first 1000 queries from the exact-scan kernel (global over all shards when N > 1)
    Qs = batches[0][:1000]
    bf = pkg.BruteforceSearch(space_of(pkg, a), a.n, device=local_rank)
    bf.addPoints(X, shard_labels)
    g = bf.searchKnnBatch(Qs, a.k)
    del bf
    stream = torch.cuda.current_stream().cuda_stream

    from research_new_hnsw_b200.sharded import ShardedSearcher, cuda_merge
    sharded = ShardedSearcher(None, cuda_merge(lambda: stream), device=dev)
    if replica:
        sharded.world = 1

    def merged(labels_t, dists_t, nq):
        """all_gather per-shard rows + GPU k-way merge; identity at world == 1."""
        sharded.local_search = lambda Q, k: (labels_t, dists_t)
        return sharded.search(None, a.k)

    gt_l = torch.from_numpy(g["labels"].view(np.int64)).to(dev)
    gt_d = torch.from_numpy(g["dists"]).to(dev)
    gt_l, _ = merged(gt_l, gt_d, len(Qs))
    gt = gt_l.cpu().numpy().view(np.uint64)
    # The ground truth must not rest on our own kernels alone: the UNMODIFIED reference's BruteforceSearch computes the
    # exact top-k of this rank's shard for a query sample, the per-shard rows are gathered and merged with numpy.
    gt_check = "oracle/_ref missing: ground truth is the exact-scan kernel's only"
    try:
        rl, rd = reference_exact_topk(a, X, shard_labels, Qs[:200], threads)
        if sw > 1:
            tl = torch.from_numpy(rl.view(np.int64)).to(dev)
            td = torch.from_numpy(rd).to(dev)
            al = torch.empty((sw,) + tuple(tl.shape), dtype=tl.dtype, device=dev)
            ad = torch.empty((sw,) + tuple(td.shape), dtype=td.dtype, device=dev)
            dist.all_gather_into_tensor(al, tl)
            dist.all_gather_into_tensor(ad, td)
            rl, rd = merge_rows_numpy(al.permute(1, 0, 2).reshape(200, -1).cpu().numpy().view(np.uint64),
                                      ad.permute(1, 0, 2).reshape(200, -1).cpu().numpy(), a.k)
        if not np.array_equal(rl, gt[:200]):
            raise AssertionError("exact-scan ground truth differs from the reference BruteforceSearch")
        gt_check = "exact-scan kernel%s; first 200 queries identical to the reference BruteforceSearch%s" % (
            " + merge kernel" if sw > 1 else "", " per shard + numpy merge" if sw > 1 else "")
    except (OSError, FileNotFoundError):
        pass

    from research_new_hnsw_b200.sharded import PackedShardExchange
    packed = PackedShardExchange(a.nq, a.k, dev) if sw > 1 else None

    def dev_search(dQ, nq, ef, work=None):
        if packed is not None and nq == a.nq:
            # full batches at N > 1: results go straight into this rank's block, one all_gather, one merge launch
            pl_, pd_ = packed.local_ptrs()
            idx.searchKnnDevice(dQ.data_ptr(), nq, a.k, ef, pl_, pd_, 0, work.data_ptr() if work is not None else 0, stream)
            return packed.exchange_and_merge(stream)
        ol = torch.empty((nq, a.k), dtype=torch.int64, device=dev)
        od = torch.empty((nq, a.k), dtype=torch.float32, device=dev)
        idx.searchKnnDevice(dQ.data_ptr(), nq, a.k, ef, ol.data_ptr(), od.data_ptr(), 0,
                            work.data_ptr() if work is not None else 0, stream)
        return merged(ol, od, nq)

    dQs = torch.from_numpy(Qs).to(dev)

    def sweep_fn(Q, ef):
        ol, _ = dev_search(dQs, len(Qs), ef)
        torch.cuda.synchronize()
        return ol.cpu().numpy().view(np.uint64)

    ef, rec, table = pick_ef(a, sweep_fn, gt, Qs)
    if world > 1:  # every rank must use the same ef
        t = torch.tensor([ef], device=dev)
        dist.broadcast(t, 0)
        ef = int(t.item())

    dbatches = [torch.from_numpy(b).to(dev) for b in batches]
    merge_check = None
    if sw > 1:
        # N > 1 data path proven inline: the packed all_gather + merge kernel must return exactly the k best
        # (dist, label) pairs of the per-shard rows, which are gathered separately and merged with numpy
        ml, md = dev_search(dbatches[0], a.nq, ef)
        torch.cuda.synchronize()
        ml, md = ml.cpu().numpy().view(np.uint64).copy(), md.cpu().numpy().copy()
        ol_ = torch.empty((a.nq, a.k), dtype=torch.int64, device=dev)
        od_ = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
        idx.searchKnnDevice(dbatches[0].data_ptr(), a.nq, a.k, ef, ol_.data_ptr(), od_.data_ptr(), 0, 0, stream)
        al = torch.empty((sw, a.nq, a.k), dtype=torch.int64, device=dev)
        ad = torch.empty((sw, a.nq, a.k), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(al, ol_)
        dist.all_gather_into_tensor(ad, od_)
        el, ed = merge_rows_numpy(al.permute(1, 0, 2).reshape(a.nq, -1).cpu().numpy().view(np.uint64),
                                  ad.permute(1, 0, 2).reshape(a.nq, -1).cpu().numpy(), a.k)
        if not (np.array_equal(el, ml) and np.array_equal(ed, md)):
            raise AssertionError("merged rows differ from the numpy merge of the gathered per-shard rows")
        owners = np.bincount((ml[ml < np.uint64(sw * a.n)] // np.uint64(a.n)).astype(np.int64), minlength=sw)[:sw]
        merge_check = {"rows_checked": int(a.nq), "equal_to_numpy_merge_of_gathered_shard_rows": True,
                       "result_share_per_shard": [round(float(x) / ml.size, 4) for x in owners]}
    # counted work of every batch (outside the timed region; identical launches)
    works = []
    for dQ in dbatches:
        w = torch.zeros((a.nq, 4), dtype=torch.int32, device=dev)
        dev_search(dQ, a.nq, ef, w)
        torch.cuda.synchronize()
        works.append(w.cpu().numpy().astype(np.int64))
    from research_new_hnsw_b200.sharded import PipelinedShardSearch
    pipe, exchange = None, None
    if sw > 1 and not os.environ.get("B200HNSW_BENCH_NO_PIPELINE"):
        # exchange of the per-shard rows: peer-to-peer pushes by the copy engines + stream memory flags (csrc/exchange.cu);
        # B200HNSW_EXCHANGE=nccl (or a failed IPC set-up on any rank) -> one packed NCCL all_gather per batch
        exchange = os.environ.get("B200HNSW_EXCHANGE", "p2p")
        depth = int(os.environ.get("B200HNSW_PIPE_DEPTH", "2"))
        if exchange == "p2p":
            ok = torch.ones(1, device=dev)
            try:
                pipe = PipelinedShardSearch(idx, a.nq, a.k, dev, depth=depth, exchange="p2p")
            except Exception as e:  # noqa: BLE001 -- no peer access / IPC on this box
                sys.stderr.write("rank %d: p2p exchange unavailable (%s)\n" % (rank, e))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                pipe, exchange = None, "nccl"
        if pipe is None:
            pipe = PipelinedShardSearch(idx, a.nq, a.k, dev, depth=depth)
        # the pipelined exchange that is timed below must return the rows the inline check above has just verified
        pl_, pd_, pev = pipe.submit(dbatches[0].data_ptr(), ef)
        pev.synchronize()
        if not (np.array_equal(pl_.cpu().numpy().view(np.uint64), ml) and np.array_equal(pd_.cpu().numpy(), md)):
            raise AssertionError("pipelined %s exchange returned different rows" % exchange)
        merge_check["timed_exchange"] = exchange
        merge_check["timed_exchange_equal_to_checked_rows"] = True
    for s in range(a.warmup):
        if pipe is not None:
            pipe.submit(dbatches[s % len(dbatches)].data_ptr(), ef)
        else:
            dev_search(dbatches[s % len(dbatches)], a.nq, ef)
    if pipe is not None:
        pipe.drain()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    submit_ms = call_ms = None
    if pipe is not None:
        # N > 1: batch i's all_gather + merge overlaps batch i+1's search kernel (side stream, event-ordered);
        # the timed region ends only after the last exchange has finished.
        # The host keeps at most `inflight` batches enqueued (it waits for the exchange of batch s-inflight): with an
        # unbounded queue the NCCL kernel of batch s is dispatched behind search kernels queued long before it.
        inflight = int(os.environ.get("B200HNSW_BENCH_INFLIGHT", "2"))
        evs = []
        t_submit = time.perf_counter()
        t_calls = 0.0
        for s in range(a.steps):
            if inflight > 0 and s >= inflight:
                evs[s - inflight].synchronize()
            t_c = time.perf_counter()
            evs.append(pipe.submit(dbatches[s % len(dbatches)].data_ptr(), ef)[2])
            t_calls += time.perf_counter() - t_c
        submit_ms = 1e3 * (time.perf_counter() - t_submit) / a.steps
        call_ms = 1e3 * t_calls / a.steps
        pipe.drain()
    else:
        for s in range(a.steps):
            dev_search(dbatches[s % len(dbatches)], a.nq, ef)
    ev1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = sampler.result()
    ms_total = ev0.elapsed_time(ev1)
    # search-kernel-only timing (same launches, no collective) for the roofline
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ol = torch.empty((a.nq, a.k), dtype=torch.int64, device=dev)
    od = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    k0.record()
    for s in range(a.steps):
        idx.searchKnnDevice(dbatches[s % len(dbatches)].data_ptr(), a.nq, a.k, ef, ol.data_ptr(), od.data_ptr(), 0, 0,
                            stream)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / a.steps
    rank_diag = None
    if dist:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        # per-rank diagnostics: search-kernel ms and host submit ms per step (explains step time vs kernel time)
        mine = torch.tensor([kernel_ms, submit_ms or 0.0, call_ms or 0.0], device=dev, dtype=torch.float64)
        allr = torch.empty((world, 3), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        # host_submit_ms: host loop per step incl. its wait for the batch `inflight` steps back; host_call_ms: the part
        # spent inside the submit call itself (launches, copies, stream memory operations)
        rank_diag = {"kernel_ms": [round(x, 4) for x in allr[:, 0].tolist()],
                     "host_submit_ms": [round(x, 4) for x in allr[:, 1].tolist()],
                     "host_call_ms": [round(x, 4) for x in allr[:, 2].tolist()]}

    # ---- C2's ef sweep (32..256): device-resident QPS and recall per ef, outside the headline timed region
    ef_table = []
    if world == 1:
        for e in (16, 32, 48, 64, 96, 128, 192, 256):
            rec_e = recall_at_k(sweep_fn(Qs, e), gt)
            for s in range(2):
                dev_search(dbatches[s % len(dbatches)], a.nq, e)
            torch.cuda.synchronize()
            best = None
            for rep in range(2):  # best of two short runs: a 3-launch timing is exposed to a single host hiccup
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record()
                for s in range(4):
                    dev_search(dbatches[s % len(dbatches)], a.nq, e)
                t1e.record()
                torch.cuda.synchronize()
                ms_e = t0e.elapsed_time(t1e) / 4
                best = ms_e if best is None else min(best, ms_e)
            ef_table.append({"ef": e, "recall_at_10": round(rec_e, 4), "qps": a.nq / (best * 1e-3)})

    # ---- end to end through the host-pointer C ABI: pinned host buffers, H2D + kernel + D2H inside the timed region
    hq = [torch.from_numpy(b).pin_memory() for b in batches]
    # page-locked result buffers, as a serving host would keep them
    pl = torch.empty((a.nq, a.k), dtype=torch.int64).pin_memory()
    pd = torch.empty((a.nq, a.k), dtype=torch.float32).pin_memory()
    pc = torch.empty((a.nq,), dtype=torch.int32).pin_memory()
    hout = {"labels": pl.numpy().view(np.uint64), "dists": pd.numpy(), "counts": pc.numpy().view(np.uint32)}
    for s in range(max(3, a.warmup)):
        idx.searchKnnBatch(hq[s % len(hq)].numpy(), a.k, ef=ef, out=hout)
    if dist:
        dist.barrier()
    dq_buf = torch.empty((a.nq, a.dim), dtype=torch.float32, device=dev)

    def e2e_step(s):
        if sw == 1:  # the reference-facing host-pointer C ABI: H2D + kernel + D2H inside the call
            idx.searchKnnBatch(hq[s % len(hq)].numpy(), a.k, ef=ef, out=hout)
        else:  # sharded public API: pinned H2D -> per-shard search -> all_gather + merge -> D2H of the merged rows
            dq_buf.copy_(hq[s % len(hq)], non_blocking=True)
            ol_, od_ = dev_search(dq_buf, a.nq, ef)
            pl.copy_(ol_, non_blocking=True)
            pd.copy_(od_, non_blocking=True)
            torch.cuda.synchronize()

    if pipe is not None:
        # sharded serving loop: two batches in flight; H2D, search, exchange and D2H of consecutive batches overlap
        # (PipelinedShardSearch.submit_host).  Every step still copies its queries in and its merged rows out.
        hl2 = [torch.empty((a.nq, a.k), dtype=torch.int64).pin_memory() for _ in range(2)]
        hd2 = [torch.empty((a.nq, a.k), dtype=torch.float32).pin_memory() for _ in range(2)]
        pend = [None, None]

        def e2e_loop(steps):
            for s in range(steps):
                j = s % 2
                if pend[j] is not None:
                    pend[j].synchronize()             # the host consumes batch s-2 before its buffers are reused
                pend[j] = pipe.submit_host(hq[s % len(hq)], ef, hl2[j], hd2[j])
            for e in pend:
                if e is not None:
                    e.synchronize()
            torch.cuda.synchronize()
    elif sw == 1 and not os.environ.get("B200HNSW_BENCH_SYNC_E2E"):
        # serving loop over the asynchronous host-pointer C ABI (b200hnsw_search_batch_submit / _wait): two batches in
        # flight, each with its own page-locked result buffers; every step still moves its queries in and its rows out
        # (the kernel reads / writes the page-locked buffers over PCIe itself).  The host consumes batch s-2 before it
        # submits batch s.  B200HNSW_BENCH_SYNC_E2E=1 measures the blocking call instead.
        outs = []
        for _ in range(2):
            l_ = torch.empty((a.nq, a.k), dtype=torch.int64).pin_memory()
            d_ = torch.empty((a.nq, a.k), dtype=torch.float32).pin_memory()
            c_ = torch.empty((a.nq,), dtype=torch.int32).pin_memory()
            outs.append({"labels": l_.numpy().view(np.uint64), "dists": d_.numpy(), "counts": c_.numpy().view(np.uint32),
                         "_keep": (l_, d_, c_)})
        hq_np = [h.numpy() for h in hq]

        def e2e_loop(steps):
            tickets = [None, None]
            for s in range(steps):
                j = s % 2
                if tickets[j] is not None:
                    idx.searchKnnBatchWait(tickets[j])
                tickets[j] = idx.searchKnnBatchSubmit(hq_np[s % len(hq_np)], a.k, outs[j], ef=ef)
            for t in tickets:
                if t is not None:
                    idx.searchKnnBatchWait(t)
    else:
        def e2e_loop(steps):
            for s in range(steps):
                e2e_step(s)

    if pipe is not None:  # the host-facing path returns the rows that were verified above
        ev_ = pipe.submit_host(hq[0], ef, hl2[0], hd2[0])
        ev_.synchronize()
        if not (np.array_equal(hl2[0].numpy().view(np.uint64), ml) and np.array_equal(hd2[0].numpy(), md)):
            raise AssertionError("submit_host returned different rows")
        merge_check["host_path_equal_to_checked_rows"] = True
    e2e_loop(3)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(a.steps)
    e2e_s = time.perf_counter() - t0
    e2e_blocking = None
    if sw == 1:  # the blocking form of the same call, for the record: one batch at a time, nothing overlaps its tail
        for s in range(3):
            e2e_step(s)
        t0b = time.perf_counter()
        for s in range(a.steps):
            e2e_step(s)
        tb = time.perf_counter() - t0b
        e2e_blocking = {"value": a.nq * a.steps / tb, "unit": "queries/s", "ms_per_step": 1e3 * tb / a.steps,
                        "api": "b200hnsw_search_batch (host pointers, pinned, blocking call)"}
    if dist:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = a.nq * a.dim * 4
    d2h = a.nq * a.k * 12 + a.nq * 4

    # ---- C5: batched GPU graph build of the same points (wall clock incl. H2D; not part of the timed search region)
    build_info = None
    if rank == 0 and world == 1 and not os.environ.get("B200HNSW_BENCH_SKIP_BUILD"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gi = pkg.HierarchicalNSW(space_of(pkg, a), a.n, a.M, a.efc, device=local_rank)
        gi.addPoints(X, shard_labels)
        gi.flush()
        gsec = time.perf_counter() - t0
        gst = gi.stats()
        rg = gi.searchKnnBatch(Qs, a.k, ef=ef)["labels"]
        # roofline of the build: every distance evaluation the kernels counted (construction searches,
        # heuristic pruning, reverse-link repair) x 4d bytes + the neighbour lists the construction searches read
        b_alg = (gst["dist_evals"] * a.dim * 4 + gst["hops_base"] * (4 + 8 * a.M) + gst["hops_upper"] * (4 + 4 * a.M)
                 + a.n * (a.dim * 4 + 8))
        peak_b, peak_b_src = peaks()
        build_info = {"gpu_points_per_s": a.n / gsec, "gpu_seconds": gsec, "gpu_kernel_ms": gst["last_kernel_ms"],
                      "kernel_launches": gst["kernel_launches"],
                      "roofline": {"bound": "hbm", "achieved": b_alg / (gst["last_kernel_ms"] * 1e-3) / 1e9, "peak": peak_b,
                                   "unit": "GB/s", "frac": b_alg / (gst["last_kernel_ms"] * 1e-3) / 1e9 / peak_b,
                                   "frac_of_wall": b_alg / gsec / 1e9 / peak_b, "traffic": None,
                                   "algorithmic_bytes": b_alg, "peak_source": peak_b_src,
                                   "note": "counted: dist_evals*4d + list reads + one write of every row; time = CUDA "
                                           "events around all build launches (gpu_kernel_ms); frac_of_wall uses the "
                                           "wall clock incl. H2D of the rows"},
                      "dist_evals_per_point": gst["dist_evals"] / a.n,
                      "recall_at_10_gpu_built_graph": round(recall_at_k(rg, gt), 4),
                      "recall_at_10_reference_built_graph": round(rec, 4), "ef": ef,
                      "reference_points_per_s": (a.n / build_s) if build_s else None, "reference_threads": threads}
        del gi
    del X
    c4 = None
    if rank == 0 and world == 1 and a.metric == "l2" and not os.environ.get("B200HNSW_BENCH_SKIP_C4"):
        del idx  # C4 needs 3 GB of rows + 1.5 GB bf16 copy; nothing else below touches the HNSW index
        idx = None
        try:
            c4 = c4_leg(pkg, local_rank, max(3, min(a.steps, 10)), 3)
        except Exception as e:
            c4 = {"error": str(e)}

    if rank == 0:
        bytes_sum = [algorithmic_bytes(a, w, ef) for w in works]
        per_launch = float(np.mean([b[0] for b in bytes_sum]))
        peak, peak_src = peaks()
        achieved = per_launch / (kernel_ms * 1e-3) / 1e9
        # dram__bytes_read+write of ONE launch from an `ncu --set full` capture (profiles/): only quoted when the
        # capture was taken on this very configuration (same ef, 1 GPU, f32 rows); never measured inside this run
        traffic, traffic_src = None, "none for this configuration (ncu capture is keyed by ef / storage / n_gpus)"
        tp = os.path.join(ROOT, "profiles", "search_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                for ent in tj.get("captures", []):
                    if ent.get("ef") == ef and ent.get("n_gpus", 1) == world and ent.get("storage", "f32") == a.storage \
                            and ent.get("metric", "l2") == a.metric and ent.get("n", 1_000_000) == a.n:
                        traffic, traffic_src = ent["dram_bytes_per_launch"], ent.get("source", tp)
            except Exception:
                pass
        resets = int(sum(w[:, 3].sum() for w in works))
        cpu_leg = None  # measured on rank 0 at N = 1 only
        if world == 1:
            try:
                cpu_leg, _ = cpu_reference_leg(a, path, batches, ef, a.cpu_seconds, os.cpu_count() or 1)
            except Exception as e:  # the checker binary is missing: report it, never substitute
                cpu_leg = {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference",
                           "sample": "unavailable: %s" % e}
        line = {
            "metric": metric_name(a), "value": world * a.nq * a.steps / (ms_total * 1e-3),
            "unit": "queries/s" if (world == 1 or replica) else "shard-searches/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup,
            "merged_qps": (world if replica else 1) * a.nq * a.steps / (ms_total * 1e-3),
            "total_points": a.n * (1 if replica else world),
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if a.storage == "f32" else "f32 accumulate over bf16 rows, f32 re-rank", "data": "synthetic",
            "config": {"workload": workload_name(a), "ef": ef, "recall_at_10": round(rec, 4), "recall_sweep": table,
                       "storage": a.storage, "ef_table": ef_table,
                       "parallelism": "1 GPU" if world == 1 else
                       ("replica%d: the same %d-point index on every GPU, each GPU serves its own query batches, no "
                        "data-path collective" % (world, a.n)) if replica else
                       "shard%d: one %d-point sub-index per GPU (%d points in total), queries replicated, per-shard rows exchanged by %s + GPU merge per batch%s; value counts (query, shard) searches, merged_qps = value/%d is the "
                       "rate of merged answers over the whole data set; ef is the smallest whose MERGED recall reaches the "
                       "target, so it falls as N grows (each shard owes only its share of the global top-k)"
                       % (world, a.n, world * a.n,
                          "peer-to-peer copy-engine pushes over NVLink with stream memory flags (no collective kernel)" if exchange == "p2p"
                          else "ONE packed NCCL all_gather",
                          ", exchange of batch i overlapped with the search of batch i+1" if pipe else "", world),
                       "ground_truth": gt_check, "merge_check": merge_check,
                       "l2_policy": "inputs larger than L2 (index %.0f MB vs 126 MB L2); %d distinct query batches cycled"
                                    % ((a.n * (a.dim * 4 + 8 * a.M)) / 1e6, len(batches)),
                       "graph": graph_note, "build": build_info, "c4": c4,
                       "visited_table_rebuilds_per_batch": resets / len(works)},
            "clocks": clocks, "per_rank": rank_diag,
            "e2e": {"value": world * a.nq * a.steps / e2e_s,
                    "unit": "queries/s" if (world == 1 or replica) else "shard-searches/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / a.steps, "blocking_call": e2e_blocking,
                    "api": ("b200hnsw_search_batch (host pointers, pinned, blocking call)" if os.environ.get("B200HNSW_BENCH_SYNC_E2E")
                            else "b200hnsw_search_batch_submit / _wait (host pointers, pinned; two batches in flight)") if sw == 1 else
                           ("PipelinedShardSearch.submit_host: page-locked queries read and merged rows stored by the kernels "
                            "themselves (zero-copy), b200hnsw_search_batch_device, %s, merge kernel; two batches in flight" % ("b200hnsw_exchange_step (P2P pushes)" if exchange == "p2p"
                                                                     else "packed NCCL all_gather")) if pipe is not None else
                           "ShardedSearcher: pinned H2D, b200hnsw_search_batch_device, NCCL all_gather, merge kernel, D2H"},
            "gpu_launches": a.steps * (1 if sw == 1 else 2),  # search kernel (+ merge kernel at N > 1)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "hnsw_search_kernel<team %d, %s>" % (64 if a.nq >= 2368 else 128, a.metric), "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": per_launch, "peak_source": peak_src,
                         "per_query": {"D": bytes_sum[0][1] / a.nq, "H0": bytes_sum[0][2] / a.nq,
                                       "Hup": bytes_sum[0][3] / a.nq}},
            "cpu_baseline": cpu_leg,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
    else:
        run_b200(a, rank, local_rank, world)


if __name__ == "__main__":
    main()

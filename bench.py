#!/usr/bin/env python
"""bench.py -- headline benchmark: QPS @ recall@10 >= 0.95 of batched searchKnn (BASELINE.json metric).

Workload (BASELINE.json configs[1], "C2"): synthetic SIFT-shaped 1M x 128 fp32, L2, M=32, ef_construction=200,
batches of 10 000 queries, k=10; ef = the smallest value of the sweep whose recall@10 (exact ground truth, first
1000 queries) is >= 0.95.  A "step" is one batch of 10 000 queries through the hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm  (CUDA kernels through the C ABI)
  python bench.py --impl reference ...                           reference arm (unmodified hnswlib on host cores)

N > 1 (torchrun, one rank per GPU): the data set is sharded, one 1M-point sub-index per GPU (weak scaling), every
rank searches the same query batch, per-shard top-k are exchanged with an NCCL all_gather and merged on the GPU
(SURVEY.md 8(e)).  `value` counts shard-level searches (N x nq per step); merged queries/s is value / N.
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


def run_reference(a, rank, world):
    if rank != 0:
        return
    from oracle import bind
    if not bind.ref_available("sse"):
        # the compiled reference (oracle/_ref, built from /root/reference in the build container) did not travel
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhnswref_sse.so is missing on this box"}),
              flush=True)
        return
    threads = os.cpu_count() or 1
    X = shard_data(a, 0)
    batches = query_batches(a)
    path, build_s = build_graph_with_reference(a, 0, X, threads)
    ref = bind.Ref(bind.best_ref_level())
    bf = ref.bf_new(ref_metric(a), a.dim, a.n)
    bf.add(X)
    Qs = batches[0][:1000]
    gt = bf.search(Qs, a.k, threads=threads)["labels"]
    del bf
    idx = ref.hnsw_load(ref_metric(a), a.dim, path)
    ef, rec, table = pick_ef(a, lambda Q, e: idx.search(Q, a.k, e, threads=threads)["labels"], gt, Qs)
    ef_table = []
    for e in (16, 32, 48, 64, 96, 128, 192, 256):
        r_e = idx.search(batches[0][:2000], a.k, e, threads=threads)
        ef_table.append({"ef": e, "recall_at_10": round(recall_at_k(r_e["labels"][:1000], gt), 4),
                         "qps": 2000 / r_e["seconds"]})
    for _ in range(a.warmup):
        idx.search(batches[0][:2000], a.k, ef, threads=threads)
    sec = 0.0
    # each step = a bounded sample of the batch so the whole run stays within minutes even on few cores
    sample = min(a.nq, 10_000)
    for s in range(a.steps):
        sec += idx.search(batches[s % len(batches)][:sample], a.k, ef, threads=threads)["seconds"]
    qps = a.steps * sample / sec
    line = {"impl": "reference", "metric": metric_name(a), "value": qps, "unit": "queries/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sec / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "ef": ef, "recall_at_10": round(rec, 4), "recall_sweep": table,
                       "ef_table": ef_table,
                       "graph": "built by the reference (addPoint, %d threads)%s" %
                                (threads, "" if build_s == 0 else " in %.1f s = %.0f points/s" % (build_s, a.n / build_s)),
                       "step": "%d queries of the batch" % sample},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "reference",
                             "sample": "%d queries per step, ef=%d, -O3 %s build of the unmodified headers"
                                       % (sample, ef, bind.best_ref_level())},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    # counted work of every batch (outside the timed region; identical launches)
    works = []
    for dQ in dbatches:
        w = torch.zeros((a.nq, 4), dtype=torch.int32, device=dev)
        dev_search(dQ, a.nq, ef, w)
        torch.cuda.synchronize()
        works.append(w.cpu().numpy().astype(np.int64))
    from research_new_hnsw_b200.sharded import PipelinedShardSearch
    pipe = None
    if sw > 1 and not os.environ.get("B200HNSW_BENCH_NO_PIPELINE"):
        pipe = PipelinedShardSearch(idx, a.nq, a.k, dev, depth=int(os.environ.get("B200HNSW_PIPE_DEPTH", "2")))
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
    submit_ms = None
    if pipe is not None:
        # N > 1: batch i's all_gather + merge overlaps batch i+1's search kernel (side stream, event-ordered);
        # the timed region ends only after the last exchange has finished.
        # The host keeps at most `inflight` batches enqueued (it waits for the exchange of batch s-inflight): with an
        # unbounded queue the NCCL kernel of batch s is dispatched behind search kernels queued long before it.
        inflight = int(os.environ.get("B200HNSW_BENCH_INFLIGHT", "2"))
        evs = []
        t_submit = time.perf_counter()
        for s in range(a.steps):
            if inflight > 0 and s >= inflight:
                evs[s - inflight].synchronize()
            evs.append(pipe.submit(dbatches[s % len(dbatches)].data_ptr(), ef)[2])
        submit_ms = 1e3 * (time.perf_counter() - t_submit) / a.steps
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
        mine = torch.tensor([kernel_ms, submit_ms or 0.0], device=dev, dtype=torch.float64)
        allr = torch.empty((world, 2), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        rank_diag = {"kernel_ms": [round(x, 4) for x in allr[:, 0].tolist()],
                     "host_submit_ms": [round(x, 4) for x in allr[:, 1].tolist()]}

    # ---- C2's ef sweep (32..256): device-resident QPS and recall per ef, outside the headline timed region
    ef_table = []
    if world == 1:
        for e in (16, 32, 48, 64, 96, 128, 192, 256):
            rec_e = recall_at_k(sweep_fn(Qs, e), gt)
            dev_search(dbatches[0], a.nq, e)
            torch.cuda.synchronize()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for s in range(3):
                dev_search(dbatches[s % len(dbatches)], a.nq, e)
            t1e.record()
            torch.cuda.synchronize()
            ef_table.append({"ef": e, "recall_at_10": round(rec_e, 4), "qps": 3 * a.nq / (t0e.elapsed_time(t1e) * 1e-3)})

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
    else:
        def e2e_loop(steps):
            for s in range(steps):
                e2e_step(s)

    e2e_loop(3)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(a.steps)
    e2e_s = time.perf_counter() - t0
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
        build_info = {"gpu_points_per_s": a.n / gsec, "gpu_seconds": gsec, "gpu_kernel_ms": gst["last_kernel_ms"],
                      "dist_evals_per_point": gst["dist_evals"] / a.n,
                      "recall_at_10_gpu_built_graph": round(recall_at_k(rg, gt), 4),
                      "recall_at_10_reference_built_graph": round(rec, 4), "ef": ef,
                      "reference_points_per_s": (a.n / build_s) if build_s else None, "reference_threads": threads}
        del gi
    del X

    if rank == 0:
        bytes_sum = [algorithmic_bytes(a, w, ef) for w in works]
        per_launch = float(np.mean([b[0] for b in bytes_sum]))
        peak, peak_src = peaks()
        achieved = per_launch / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "search_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
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
            "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if a.storage == "f32" else "f32 accumulate over bf16 rows, f32 re-rank", "data": "synthetic",
            "config": {"workload": workload_name(a), "ef": ef, "recall_at_10": round(rec, 4), "recall_sweep": table,
                       "storage": a.storage, "ef_table": ef_table,
                       "parallelism": "1 GPU" if world == 1 else
                       ("replica%d: the same %d-point index on every GPU, each GPU serves its own query batches, no "
                        "data-path collective" % (world, a.n)) if replica else
                       "shard%d: one %d-point sub-index per GPU, queries replicated, ONE packed NCCL all_gather + GPU merge "
                       "per batch%s; value counts shard-level searches (merged queries/s = value/%d)"
                       % (world, a.n, ", exchange of batch i overlapped with the search of batch i+1" if pipe else "", world),
                       "l2_policy": "inputs larger than L2 (index %.0f MB vs 126 MB L2); %d distinct query batches cycled"
                                    % ((a.n * (a.dim * 4 + 8 * a.M)) / 1e6, len(batches)),
                       "graph": graph_note, "build": build_info,
                       "visited_table_rebuilds_per_batch": resets / len(works)},
            "clocks": clocks, "per_rank": rank_diag,
            "e2e": {"value": world * a.nq * a.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / a.steps,
                    "api": "b200hnsw_search_batch (host pointers, pinned)" if sw == 1 else
                           "PipelinedShardSearch.submit_host: pinned H2D, b200hnsw_search_batch_device, packed NCCL "
                           "all_gather, merge kernel, pinned D2H; two batches in flight" if pipe is not None else
                           "ShardedSearcher: pinned H2D, b200hnsw_search_batch_device, NCCL all_gather, merge kernel, D2H"},
            "gpu_launches": a.steps * (1 if sw == 1 else 2),  # search kernel (+ merge kernel at N > 1)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "hnsw_search_kernel<team %d, %s>" % (64 if a.nq >= 2368 else 128, a.metric), "kernel_ms": kernel_ms,
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

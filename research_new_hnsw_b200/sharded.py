"""Multi-GPU plumbing: one sub-index per GPU, queries replicated, per-shard top-k exchanged with an
all_gather and merged.  One process per GPU; `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU
tests) is the plumbing, the merge itself is the CUDA kernel behind ``b200hnsw_merge_topk_device``.

The exchange is a few hundred KB to a few MB per step (nq * k * 12 bytes per rank) -- latency-bound, so it is a plain
collective followed by the merge on the same stream rather than a fused kernel.
"""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous label range [lo, hi) owned by `rank`; labels stay global so no remap after the merge."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk_numpy(labels_all, dists_all, k):
    """Reference semantics of the merge kernel: per query the k smallest (dist, label) pairs over all shards, closest
    first.  labels_all/dists_all are [shards][nq][k]; padding rows carry (inf, UINT64_MAX) and sort last."""
    s, nq, kk = labels_all.shape
    L = np.transpose(labels_all, (1, 0, 2)).reshape(nq, s * kk)
    D = np.transpose(dists_all, (1, 0, 2)).reshape(nq, s * kk)
    out_l = np.empty((nq, k), np.uint64)
    out_d = np.empty((nq, k), np.float32)
    for i in range(nq):
        order = np.lexsort((L[i], D[i]))[:k]
        out_l[i], out_d[i] = L[i][order], D[i][order]
    return out_l, out_d


class ShardedSearcher:
    """search(Q, k) over `world` shards.

    local_search(Q, k) -> (labels[nq,k] int64-viewed-uint64, dists[nq,k] float32) torch tensors on `device`
    merge(labels_all[world,nq,k], dists_all[world,nq,k], k) -> (labels[nq,k], dists[nq,k]) torch tensors
    """

    def __init__(self, local_search, merge, group=None, device="cpu"):
        import torch.distributed as dist
        self.dist, self.group, self.device = dist, group, device
        self.local_search, self.merge = local_search, merge
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def broadcast_queries(self, Q, src=0):
        """Rank `src` owns the query batch; everybody else receives it (north_star: 'queries broadcast')."""
        if self.world > 1:
            self.dist.broadcast(Q, src, group=self.group)
        return Q

    def search(self, Q, k):
        import torch
        labels, dists = self.local_search(Q, k)
        if self.world == 1:
            return labels, dists
        gl = torch.empty((self.world,) + tuple(labels.shape), dtype=labels.dtype, device=labels.device)
        gd = torch.empty((self.world,) + tuple(dists.shape), dtype=dists.dtype, device=dists.device)
        if labels.is_cuda:
            self.dist.all_gather_into_tensor(gl, labels.contiguous(), group=self.group)
            self.dist.all_gather_into_tensor(gd, dists.contiguous(), group=self.group)
        else:  # gloo
            self.dist.all_gather(list(gl.unbind(0)), labels.contiguous(), group=self.group)
            self.dist.all_gather(list(gd.unbind(0)), dists.contiguous(), group=self.group)
        return self.merge(gl, gd, k)


def cuda_merge(stream_getter=None):
    """merge callable for ShardedSearcher backed by the CUDA merge kernel."""
    import torch
    from . import capi

    def merge(gl, gd, k):
        world, nq, kk = gl.shape
        ol = torch.empty((nq, k), dtype=gl.dtype, device=gl.device)
        od = torch.empty((nq, k), dtype=gd.dtype, device=gd.device)
        stream = stream_getter() if stream_getter else torch.cuda.current_stream().cuda_stream
        assert kk == k
        capi.merge_topk_device(gl.data_ptr(), gd.data_ptr(), world, nq, k, ol.data_ptr(), od.data_ptr(), stream)
        return ol, od

    return merge


class PackedShardExchange:
    """Throughput path of the sharded search: every rank owns a block [labels | dists] of its results, the search
    kernel writes straight into it, ONE all_gather_into_tensor moves all blocks, the merge kernel reads them in place.
    All buffers are allocated once (no per-step allocations, one collective, one merge launch per step)."""

    def __init__(self, nq, k, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nq, self.k = nq, k
        self.block = (nq * k * 12 + 7) // 8 * 8
        self.mine = torch.empty(self.block, dtype=torch.uint8, device=device)
        self.all = torch.empty(self.world * self.block, dtype=torch.uint8, device=device)
        self.out_l = torch.empty((nq, k), dtype=torch.int64, device=device)
        self.out_d = torch.empty((nq, k), dtype=torch.float32, device=device)

    def local_ptrs(self):
        """device addresses the shard's search must write its labels / dists to"""
        base = self.mine.data_ptr()
        return base, base + self.nq * self.k * 8

    def exchange_and_merge(self, stream, out_ptrs=None):
        """out_ptrs = (labels, dists) device-accessible addresses for the merged rows (default: this object's tensors)"""
        from . import capi
        if self.world == 1:
            src = self.mine
        else:
            self.dist.all_gather_into_tensor(self.all, self.mine, group=self.group)
            src = self.all
        ol, od = out_ptrs if out_ptrs else (self.out_l.data_ptr(), self.out_d.data_ptr())
        capi.merge_topk_packed_device(src.data_ptr(), self.block, self.world, self.nq, self.k, ol, od, stream)
        return self.out_l, self.out_d


class P2PShardExchange:
    """Same role as PackedShardExchange without a collective: the rank's block is pushed into every peer's receive area by
    the copy engines over NVLink, arrival is signalled and awaited with stream memory operations (capi.P2PExchange,
    csrc/exchange.cu) -- no SM is taken from the search kernel of the next batch, and nothing spins while the slowest rank
    catches up.  One object serves all steps; `views(depth)` hands out the per-slot objects PipelinedShardSearch cycles
    (each with its own merged-output tensors)."""

    def __init__(self, nq, k, device, group=None):
        import torch
        import torch.distributed as dist
        from . import capi
        self.torch, self.nq, self.k = torch, nq, k
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.block = (nq * k * 12 + 7) // 8 * 8
        self.step_no = 0

        def gather(desc):
            mine = torch.frombuffer(bytearray(desc), dtype=torch.uint8).to(device)
            allt = torch.empty(self.world * len(desc), dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allt, mine, group=group)
            return bytes(allt.cpu().numpy().tobytes())

        self.x = capi.P2PExchange(device.index if device.index is not None else torch.cuda.current_device(), self.world,
                                  self.rank, self.block, gather)
        self.device = device

    class _View:
        def __init__(self, parent):
            t = parent.torch
            self.p = parent
            self.out_l = t.empty((parent.nq, parent.k), dtype=t.int64, device=parent.device)
            self.out_d = t.empty((parent.nq, parent.k), dtype=t.float32, device=parent.device)
            self.step = 0
            self.allb = 0

        def local_ptrs(self):
            """starts the next step: device addresses the shard's search must write its labels / dists to"""
            self.p.step_no += 1
            self.step = self.p.step_no
            mine, self.allb = self.p.x.slot(self.step)
            return mine, mine + self.p.nq * self.p.k * 8

        def exchange_and_merge(self, stream, out_ptrs=None):
            from . import capi
            self.p.x.step(self.step, stream)
            ol, od = out_ptrs if out_ptrs else (self.out_l.data_ptr(), self.out_d.data_ptr())
            capi.merge_topk_packed_device(self.allb, self.p.block, self.p.world, self.p.nq, self.p.k, ol, od, stream)
            return self.out_l, self.out_d

    def views(self, depth):
        return [P2PShardExchange._View(self) for _ in range(depth)]


class PipelinedShardSearch:
    """Back-to-back batches at N > 1: the exchange (all_gather + merge) of batch i runs on a side stream while the
    search kernel of batch i+1 already runs on the caller's stream, so collective latency and rank skew are hidden
    behind compute instead of serialising with it.  ``depth`` result blocks are cycled; a block is reused only after
    its exchange finished (event wait, no host sync)."""

    def __init__(self, index, nq, k, device, depth=2, group=None, exchange="nccl"):
        import torch
        self.torch, self.index, self.nq, self.k = torch, index, nq, k
        if exchange == "p2p":  # copy-engine pushes + stream memory flags instead of the NCCL all_gather
            self.p2p = P2PShardExchange(nq, k, device, group)
            self.slots = self.p2p.views(depth)
        else:
            self.slots = [PackedShardExchange(nq, k, device, group) for _ in range(depth)]
        self.done = [None] * depth
        self.side = torch.cuda.Stream(device=device, priority=-1)
        self.i = 0

    def submit(self, d_queries, ef, d_work=0, out_ptrs=None):
        """enqueue one batch; returns (labels, dists, event) -- the tensors are valid once ``event`` completed.
        ``d_queries`` / ``out_ptrs`` may be addresses of page-locked host memory (device-accessible): the search kernel
        then reads the queries and the merge kernel stores the merged rows over PCIe themselves."""
        torch = self.torch
        main = torch.cuda.current_stream()
        j = self.i % len(self.slots)
        self.i += 1
        slot = self.slots[j]
        if self.done[j] is not None and not self.done[j].query():
            # (only when the exchange that last used this slot is still running: a wait between two search launches
            # would also take away their programmatic overlap)
            main.wait_event(self.done[j])
        pl, pd = slot.local_ptrs()
        self.index.searchKnnDevice(d_queries, self.nq, self.k, ef, pl, pd, 0, d_work, main.cuda_stream)
        searched = torch.cuda.Event()
        searched.record(main)
        self.last_searched = searched
        self.side.wait_event(searched)
        with torch.cuda.stream(self.side):
            out_l, out_d = slot.exchange_and_merge(self.side.cuda_stream, out_ptrs)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[j] = ev
        return out_l, out_d, ev

    def drain(self):
        """make the caller's stream wait for every exchange in flight"""
        self.torch.cuda.current_stream().wait_stream(self.side)

    def submit_host(self, h_queries, ef, h_labels, h_dists):
        """Host-facing form of submit(): page-locked query batch in, page-locked result rows out.  The H2D copy, the
        search, the exchange and the D2H copy of consecutive batches run on four streams ordered by events, so a
        serving loop that keeps ``depth`` batches in flight overlaps all of them.  Returns the event after which
        ``h_labels`` / ``h_dists`` hold the merged rows of this batch."""
        torch = self.torch
        import os
        if (h_queries.is_pinned() and h_labels.is_pinned() and h_dists.is_pinned()
                and os.environ.get("B200HNSW_ZEROCOPY", "1") != "0"):
            # page-locked buffers are device-accessible: no staging copies, no copy streams
            return self.submit(h_queries.data_ptr(), ef, out_ptrs=(h_labels.data_ptr(), h_dists.data_ptr()))[2]
        if not hasattr(self, "_h2d"):
            dev = self.slots[0].out_l.device
            self._h2d = torch.cuda.Stream(device=dev)
            self._d2h = torch.cuda.Stream(device=dev)
            self._dq = [torch.empty(tuple(h_queries.shape), dtype=torch.float32, device=dev) for _ in self.slots]
            self._srch = [None] * len(self.slots)
            self._down = [None] * len(self.slots)
        j = self.i % len(self.slots)
        main = torch.cuda.current_stream()
        if self._srch[j] is not None:
            self._h2d.wait_event(self._srch[j])      # the search that last read this query buffer
        with torch.cuda.stream(self._h2d):
            self._dq[j].copy_(h_queries, non_blocking=True)
            up = torch.cuda.Event()
            up.record(self._h2d)
        main.wait_event(up)
        if self._down[j] is not None:
            self.side.wait_event(self._down[j])       # the D2H copy that last read this slot's merged rows
        out_l, out_d, ev = self.submit(self._dq[j].data_ptr(), ef)
        self._srch[j] = self.last_searched
        self._d2h.wait_event(ev)
        with torch.cuda.stream(self._d2h):
            h_labels.copy_(out_l, non_blocking=True)
            h_dists.copy_(out_d, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._d2h)
        self._down[j] = done
        return done

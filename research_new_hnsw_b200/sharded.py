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

// merge_launch.cuh -- host-side launcher of the (dist, label) top-k merge kernels in search_kernel.cuh.
//
// merge_tree() reduces `shards` lists of k pairs per query to one.  Few candidates (<= 1024 per query): one warp per
// query ranks them in shared memory.  Up to ~25 k candidates: merge_select_kernel, one CTA per query, O(total).  More:
// groups of lists small enough for the warp kernel are merged level by level (ping-pong between the input and a
// scratch array); only k > 2048 falls back to the global-memory warp kernel in one step.
#pragma once
#include <algorithm>

#include "search_kernel.cuh"

namespace b200 {

// one level: groups of `group` consecutive shards -> one list per group, written at out + g*ostride
// sorted_rows: every input row is in ascending distance order (rows returned by a search are; the per-slice lists of the
// brute-force scans are NOT)
inline cudaError_t merge_level(const uint64_t *l, const float *d, size_t ls, size_t ds, size_t shards, size_t group,
                               size_t nq, size_t k, uint64_t *ol, float *od, size_t ols, size_t ods, cudaStream_t st,
                               bool sorted_rows = false) {
    const size_t total = std::min(group, shards) * k;
    unsigned warps = 4;
    while (warps > 1 && warps * total * 12 > 48 * 1024) warps >>= 1;
    const dim3 grid((unsigned)((nq + warps - 1) / warps), (unsigned)((shards + group - 1) / group));
    if (warps * total * 12 <= 48 * 1024) {
        if (sorted_rows)
            merge_topk_smem_kernel<true><<<grid, warps * 32, warps * total * 12, st>>>(l, d, ls, ds, (uint32_t)shards,
                                                                                     (uint32_t)group, (uint32_t)nq,
                                                                                     (uint32_t)k, ol, od, ols, ods);
        else
            merge_topk_smem_kernel<false><<<grid, warps * 32, warps * total * 12, st>>>(l, d, ls, ds, (uint32_t)shards,
                                                                                      (uint32_t)group, (uint32_t)nq,
                                                                                      (uint32_t)k, ol, od, ols, ods);
    } else
        merge_topk_kernel<<<grid, warps * 32, 0, st>>>(l, d, ls, ds, (uint32_t)shards, (uint32_t)group, (uint32_t)nq,
                                                       (uint32_t)k, ol, od, ols, ods);
    return cudaGetLastError();
}

// One CTA per query over ALL candidates: O(total) selection instead of O(total^2) ranking.  The k-th smallest
// distance key T is found by 32 rounds of counting (bisection over the ordered-float bits, keys in shared memory);
// everything below T is selected, ties at T by smallest (label, position); the k selected pairs are then ranked
// among themselves.  Same result as the warp kernels: the k smallest by (dist, label, position), ascending.
// Dynamic shared memory: (2 * total + k) * 4 bytes.
static __global__ void __launch_bounds__(1024) merge_select_kernel(const uint64_t *__restrict__ labels_in,
                                                                   const float *__restrict__ dists_in, size_t lstride,
                                                                   size_t dstride, uint32_t shards, uint32_t nq,
                                                                   uint32_t k, uint64_t *__restrict__ labels_out,
                                                                   float *__restrict__ dists_out) {
    extern __shared__ __align__(16) unsigned char select_smem[];
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t nsel, nties;
    const uint32_t qi = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t total = shards * k;
    uint32_t *keys = reinterpret_cast<uint32_t *>(select_smem);
    uint32_t *ties = keys + total;
    uint32_t *sel = ties + total;
    auto label_of = [&](uint32_t c) { return labels_in[(size_t)(c / k) * lstride + (size_t)qi * k + c % k]; };
    for (uint32_t c = tid; c < total; c += nthr)
        keys[c] = f2ord(dists_in[(size_t)(c / k) * dstride + (size_t)qi * k + c % k]);
    if (tid == 0) { nsel = 0; nties = 0; }
    __syncthreads();
    auto count_below = [&](uint32_t bound) {  // block-wide number of keys < bound (all threads get it)
        uint32_t n = 0;
        for (uint32_t c = tid; c < total; c += nthr) n += keys[c] < bound ? 1u : 0u;
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) wsum[warp] = n;
        __syncthreads();
        uint32_t all = 0;
        for (uint32_t w = 0; w < nthr / 32; w++) all += wsum[w];
        __syncthreads();
        return all;
    };
    uint32_t T = 0;  // ends as the k-th smallest key: the largest value with fewer than k keys below it
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t cand = T | (1u << bit);
        if (count_below(cand) < k) T = cand;
    }
    const uint32_t below = count_below(T), need = k - below;
    for (uint32_t c = tid; c < total; c += nthr) {
        const uint32_t key = keys[c];
        if (key < T) sel[atomicAdd(&nsel, 1u)] = c;
        else if (key == T) ties[atomicAdd(&nties, 1u)] = c;
    }
    __syncthreads();
    const uint32_t nt = nties;
    for (uint32_t t = tid; t < nt; t += nthr) {  // `need` of the ties, smallest (label, position) first
        const uint32_t c = ties[t];
        uint32_t rank = 0;
        if (nt > need) {
            const uint64_t lc = label_of(c);
            for (uint32_t u = 0; u < nt && rank < need; u++) {
                const uint32_t cu = ties[u];
                const uint64_t lu = label_of(cu);
                rank += (lu < lc || (lu == lc && cu < c)) ? 1u : 0u;
            }
        }
        if (rank < need) sel[atomicAdd(&nsel, 1u)] = c;
    }
    __syncthreads();
    for (uint32_t e = tid; e < k; e += nthr) {  // order the k selected pairs
        const uint32_t c = sel[e], key = keys[c];
        const uint64_t lc = label_of(c);
        uint32_t rank = 0;
        for (uint32_t f = 0; f < k; f++) {
            const uint32_t cf = sel[f], kf = keys[cf];
            if (kf < key) {
                rank++;
            } else if (kf == key) {
                const uint64_t lf = label_of(cf);
                rank += (lf < lc || (lf == lc && cf < c)) ? 1u : 0u;
            }
        }
        labels_out[(size_t)qi * k + rank] = lc;
        dists_out[(size_t)qi * k + rank] = ord2f(key);
    }
}

// number of scratch elements merge_tree needs for `shards` lists (0 when one level is enough)
inline size_t merge_tree_group(size_t k) { return std::max<size_t>(2, 1024 / std::max<size_t>(1, k)); }
inline bool merge_select_fits(size_t shards, size_t k) { return (2 * shards * k + k) * 4 <= 200 * 1024; }
inline size_t merge_tree_scratch(size_t shards, size_t nq, size_t k) {
    const size_t g = merge_tree_group(k);
    return (shards <= g || merge_select_fits(shards, k)) ? 0 : (shards + g - 1) / g * nq * k;
}

// in: [shards][nq][k] at (l, d) (dense), scratch: merge_tree_scratch() elements each; `l`/`d` are overwritten when
// more than two levels are needed.  Up to 1024 candidates per query: one warp-per-query launch; up to ~25 k: one
// CTA-per-query selection; beyond: the tree.  Returns the number of kernel launches through *launches.
inline cudaError_t merge_tree(uint64_t *l, float *d, size_t shards, size_t nq, size_t k, uint64_t *sl, float *sd,
                              uint64_t *ol, float *od, cudaStream_t st, unsigned *launches = nullptr) {
    const size_t g = merge_tree_group(k), row = nq * k;
    if (shards > g && merge_select_fits(shards, k)) {
        const size_t bytes = (2 * shards * k + k) * 4;
        cudaError_t e = cudaFuncSetAttribute(merge_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        const unsigned threads = shards * k >= 8192 ? 1024 : 256;
        merge_select_kernel<<<(unsigned)nq, threads, bytes, st>>>(l, d, row, row, (uint32_t)shards, (uint32_t)nq,
                                                                  (uint32_t)k, ol, od);
        if (launches) *launches = 1;
        return cudaGetLastError();
    }
    uint64_t *cl = l, *nl = sl;
    float *cd = d, *nd = sd;
    unsigned count = 0;
    while (shards > g) {
        const size_t groups = (shards + g - 1) / g;
        cudaError_t e = merge_level(cl, cd, row, row, shards, g, nq, k, nl, nd, row, row, st);
        if (e != cudaSuccess) return e;
        count++;
        std::swap(cl, nl);
        std::swap(cd, nd);
        shards = groups;
    }
    cudaError_t e = merge_level(cl, cd, row, row, shards, shards, nq, k, ol, od, 0, 0, st);
    count++;
    if (launches) *launches = count;
    return e;
}

}  // namespace b200

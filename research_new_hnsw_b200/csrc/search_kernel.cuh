// search_kernel.cuh -- batched HNSW searchKnn on sm_100a: one 128-thread CTA per query.
//
// Replaces, for a whole batch of queries per launch (all file:line under /root/reference/hnswlib):
//   hnswalg.h:1270-1324  searchKnn            -> prologue (upper-layer greedy descent) + epilogue (top-k, labels)
//   hnswalg.h:309-440    searchBaseLayerST<bare_bone_search=true> -> the hop loop below
//   visited_list_pool.h  VisitedList tags     -> per-query open-addressing hash in shared memory
//   space_l2.h:97-143 / space_ip.h:255-303   -> sub-warp fp32 reductions (LPV lanes x 128-bit loads per vector)
//
// Equivalences used (each argued from the cited reference code):
//  * In bare-bone mode a node enters candidate_set and top_candidates together (:398-408) and leaves
//    top_candidates only as the current worst (:418-429); lowerBound never increases once the heap is full and
//    the loop stops when the best unexpanded candidate is > lowerBound (:347-358).  The live frontier is
//    therefore exactly the not-yet-expanded entries of the top-ef set: ONE sorted buffer of <= ef keys
//    (dist, id, expanded-bit) replaces both std::priority_queues.
//  * All unvisited neighbours of the expanded node are evaluated in parallel and filtered against the
//    pre-expansion bound -- a superset of what the sequential loop admits (:395); extras fall off the end of the
//    sorted buffer at the merge.  Same final buffer (up to exact float ties).
//  * "visited" means "distance already evaluated" (:385-386).  When the hash exceeds half load it is rebuilt
//    from the ids in the buffer; a forgotten node can only be re-evaluated and is then rejected by the bound
//    (it was rejected or evicted at a bound >= the current one), so results do not change.
//
// Data layout (device_index.cuh): vec[N][d4] float4 rows (zero padded), links0[N][maxM0] u32 padded with kEmpty,
// links_up[list][maxM] u32 padded with kEmpty, up_base[N] = index of the node's level-1 list or kEmpty.
#pragma once
#include "common.cuh"

namespace b200 {

// Threads per query ("team" = one CTA).  128 gives the shortest per-query latency (32 vectors in flight per hop);
// 64 doubles the number of resident queries per SM and wins on throughput for large batches.
constexpr int kTeamMax = 128;

struct SearchArgs {
    const float4 *vec;        // [n][d4]
    const uint32_t *links0;   // [n][maxM0]
    const uint32_t *up_base;  // [n]
    const uint32_t *links_up; // [lists][maxM]
    const uint64_t *labels;   // [n]
    const float *Q;           // [nq][dim]
    uint64_t *out_labels;     // [nq][k]
    float *out_dists;         // [nq][k]
    uint32_t *out_counts;     // [nq] or null
    uint32_t *out_work;       // [nq][4] or null: D, H0, Hup, resets
    const uint8_t *flags;     // [n] bit 0 = deleted (hnswalg.h:934-937); null when nothing is deleted
    const uint4 *vec16;       // [n][d16] bf16 rows (storage variant) or null
    uint32_t d16;
    uint32_t bufcap;          // entries per candidate buffer: ef (bare-bone), or ef + room for the rejected entries inside the bound
    uint32_t n, entry;
    int32_t maxlevel;
    uint32_t dim, d4, maxM, maxM0;
    uint32_t nq, k, ef;
    uint32_t hash_bits;
    uint32_t pf;              // GraphView::pf (L2 prefetch policy)
};

// shared-memory carve-up, shared by host (size) and device (pointers)
struct SearchSmem {
    uint32_t off_buf0, off_buf1, off_acc, off_ids, off_dist, off_pref, off_q, off_hash, total;
    __host__ __device__ SearchSmem(uint32_t bufcap, uint32_t list_cap, uint32_t d4, uint32_t hash_bits) {
        uint32_t o = 0;
        off_buf0 = o; o += bufcap * 8;
        off_buf1 = o; o += bufcap * 8;
        off_acc = o;  o += list_cap * 8;
        off_ids = o;  o += list_cap * 4;
        off_dist = o; o += list_cap * 4;
        off_pref = o; o += list_cap * 4;
        o = (o + 15) & ~15u;
        off_q = o;    o += d4 * 16;
        off_hash = o; o += (1u << hash_bits) * 4;
        total = o;
    }
};

// Sum over the LPV lanes of a group.  Groups of one warp may run different trip counts, so the shuffle names only
// the lanes of its own group.
template <int LPV>
__device__ __forceinline__ float group_sum(float v, uint32_t gmask) {
#pragma unroll
    for (int o = LPV / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}

// L2 prefetch (prefetch.global.L2: fire and forget, per-thread address, no register or shared-memory cost).  The
// gather of a hop runs in rounds of 2 rows per lane group (the bytes in flight are paid in registers); prefetching the
// rows of the LATER rounds when the hop starts turns their DRAM latency into an L2 hit.  (cp.async.bulk.prefetch.L2
// would cover a whole row per instruction, but it takes a warp-uniform address: ptxas serialises divergent lanes in
// a 7-instruction loop per lane, more than the four per-line prefetches it replaces.)
#ifndef B200_PF_LINE
#define B200_PF_LINE 128
#endif
constexpr uint32_t kPfLine = B200_PF_LINE;
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// rows ids[first..n) of `bytes` each, one thread per 128-byte line (consecutive threads take the consecutive lines of a
// row).  One thread per ROW -- one id read and one address per row, four back-to-back prefetches -- executes fewer
// instructions but measured 0.968 vs 0.809 ms per 10 k queries: 32 different rows per prefetch instruction.
template <int TEAM>
__device__ __forceinline__ void prefetch_rows(const char *base, uint32_t bytes, const uint32_t *ids, int first, int n) {
    const uint32_t lpr = (bytes + kPfLine - 1) / kPfLine;
    const uint32_t total = (uint32_t)max(n - first, 0) * lpr;
    for (uint32_t i = threadIdx.x; i < total; i += TEAM) {
        const uint32_t r = i / lpr, l = i - r * lpr;
        prefetch_l2(base + (size_t)ids[first + r] * bytes + l * kPfLine);
    }
}
__device__ __forceinline__ void prefetch_span(const char *p, uint32_t bytes) {
    for (uint32_t o = 0; o < bytes; o += kPfLine) prefetch_l2(p + o);
}
// GraphView::pf bits
constexpr uint32_t kPfRows = 1;      // beam search: rows of gather rounds >= 2 of the current hop
constexpr uint32_t kPfGreedy = 2;    // greedy descent: same
constexpr uint32_t kPfNewBest = 4;   // list of an admitted neighbour that beats the predicted next expansion
constexpr uint32_t kPfRound1 = 8;    // also the rows of round 1 (A/B only)
constexpr uint32_t kPfSpec = 16;     // rows of the predicted next expansion's neighbours (speculative)
constexpr uint32_t kPfSpecFilter = 32;  // ... skipping neighbours already in the visited table
constexpr uint32_t kPfEarly = 64;    // every new neighbour's row as soon as its id passed the visited filter (before the
                                     // compaction barrier) / every list entry's row in the greedy descent

// One 128-bit chunk of the distance sum on sm_100a's packed fp32 pipe (FFMA2: two fused multiply-adds per
// instruction): L2 -> diff = fma(v, -1, q), acc = fma(diff, diff, acc); IP -> acc = fma(q, v, acc).  The two halves
// of the accumulator are added once per vector.
template <int METRIC>
__device__ __forceinline__ float2 acc4(float2 acc, const float4 &q, const float4 &v) {
#ifdef B200_NO_FFMA2
    if (METRIC == 0) {
        float a = q.x - v.x, b = q.y - v.y, c = q.z - v.z, d = q.w - v.w;
        acc.x = fmaf(a, a, acc.x); acc.y = fmaf(b, b, acc.y); acc.x = fmaf(c, c, acc.x); acc.y = fmaf(d, d, acc.y);
    } else {
        acc.x = fmaf(q.x, v.x, acc.x); acc.y = fmaf(q.y, v.y, acc.y);
        acc.x = fmaf(q.z, v.z, acc.x); acc.y = fmaf(q.w, v.w, acc.y);
    }
    return acc;
#endif
    if (METRIC == 0) {
        const float2 m1 = make_float2(-1.f, -1.f);
        const float2 d0 = __ffma2_rn(make_float2(v.x, v.y), m1, make_float2(q.x, q.y));
        const float2 d1 = __ffma2_rn(make_float2(v.z, v.w), m1, make_float2(q.z, q.w));
        acc = __ffma2_rn(d0, d0, acc);
        acc = __ffma2_rn(d1, d1, acc);
    } else {
        acc = __ffma2_rn(make_float2(q.x, q.y), make_float2(v.x, v.y), acc);
        acc = __ffma2_rn(make_float2(q.z, q.w), make_float2(v.z, v.w), acc);
    }
    return acc;
}

// Distances from the query (register slices q[]) to ids[0..n): each group of LPV lanes owns one vector at a time,
// two vectors (2*CPL 128-bit loads per lane) are in flight per group.  dists[j] is written by the group leader.
template <int TEAM, int LPV, int CPL, int METRIC, bool CACHED = false>
__device__ __forceinline__ void eval_list(const float4 (&q)[CPL], const float4 *__restrict__ vec, uint32_t d4,
                                          const uint32_t *ids, int n, float *dists, int grp, int sub) {
    constexpr int NGRP = TEAM / LPV;
    const uint32_t gmask = LPV == 32 ? 0xffffffffu : (((1u << LPV) - 1u) << ((threadIdx.x & 31) / LPV * LPV));
    for (int j = grp; j < n; j += 2 * NGRP) {
        const int j2 = j + NGRP;
        const bool has2 = j2 < n;
        const float4 *ra = vec + (size_t)ids[j] * d4;
        const float4 *rb = vec + (size_t)ids[has2 ? j2 : j] * d4;
        float4 va[CPL], vb[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const uint32_t idx = sub + c * LPV;
            va[c] = idx < d4 ? (CACHED ? __ldg(ra + idx) : ldg_stream(ra + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const uint32_t idx = sub + c * LPV;
            vb[c] = (has2 && idx < d4) ? (CACHED ? __ldg(rb + idx) : ldg_stream(rb + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const uint32_t idx = sub + c * LPV;
            if (idx < d4) {
                pa = acc4<METRIC>(pa, q[c], va[c]);
                pb = acc4<METRIC>(pb, q[c], vb[c]);
            }
        }
        float sa = group_sum<LPV>(pa.x + pa.y, gmask);
        float sb = group_sum<LPV>(pb.x + pb.y, gmask);
        if (METRIC == 1) { sa = 1.0f - sa; sb = 1.0f - sb; }
        if (sub == 0) {
            dists[j] = sa;
            if (has2) dists[j2] = sb;
        }
    }
}

// Same gather as eval_list, fused with the admission test of searchBaseLayerST (hnswalg.h:395: size < ef ||
// lowerBound > dist): the group leader appends the key of an admitted neighbour to acc[] (order irrelevant, the merge
// ranks keys).
template <int TEAM, int LPV, int CPL, int METRIC, bool NB = false>
__device__ __forceinline__ void eval_admit(const float4 (&q)[CPL], const float4 *__restrict__ vec, uint32_t d4,
                                           const uint32_t *ids, int n, bool full, float bound, uint64_t *acc,
                                           int *s_acc, int grp, int sub, const uint8_t *__restrict__ flags = nullptr,
                                           const uint32_t *__restrict__ pf_lists = nullptr, uint32_t pf_bytes = 0,
                                           float pf_below = 0.f) {
    constexpr int NGRP = TEAM / LPV;
    const uint32_t gmask = LPV == 32 ? 0xffffffffu : (((1u << LPV) - 1u) << ((threadIdx.x & 31) / LPV * LPV));
    for (int j = grp; j < n; j += 2 * NGRP) {
        const int j2 = j + NGRP;
        const bool has2 = j2 < n;
        const uint32_t ida = ids[j], idb = ids[has2 ? j2 : j];
        const float4 *ra = vec + (size_t)ida * d4;
        const float4 *rb = vec + (size_t)idb * d4;
        uint32_t dela = 0, delb = 0;  // deleted marks travel in bit 30 of the key (non-bare-bone search only)
        if (NB && sub == 0) {
            dela = (uint32_t)(__ldg(flags + ida) & 1u) << 30;
            delb = (uint32_t)(__ldg(flags + idb) & 1u) << 30;
        }
        float4 va[CPL], vb[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const uint32_t idx = sub + c * LPV;
            va[c] = idx < d4 ? ldg_stream(ra + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const uint32_t idx = sub + c * LPV;
            vb[c] = (has2 && idx < d4) ? ldg_stream(rb + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const uint32_t idx = sub + c * LPV;
            if (idx < d4) {
                pa = acc4<METRIC>(pa, q[c], va[c]);
                pb = acc4<METRIC>(pb, q[c], vb[c]);
            }
        }
        float sa = group_sum<LPV>(pa.x + pa.y, gmask);
        float sb = group_sum<LPV>(pb.x + pb.y, gmask);
        if (METRIC == 1) { sa = 1.0f - sa; sb = 1.0f - sb; }
        if (sub == 0) {
            if (!full || sa < bound) acc[atomicAdd(s_acc, 1)] = make_key(sa, ida | dela);
            if (has2 && (!full || sb < bound)) acc[atomicAdd(s_acc, 1)] = make_key(sb, idb | delb);
            if (pf_lists) {  // a neighbour closer than the predicted next expansion will be expanded before it
                if (sa < pf_below) prefetch_span((const char *)pf_lists + (size_t)ida * pf_bytes, pf_bytes);
                if (has2 && sb < pf_below) prefetch_span((const char *)pf_lists + (size_t)idb * pf_bytes, pf_bytes);
            }
        }
    }
}

// bf16 storage variant of the gather: rows are 128-bit chunks of 8 bf16, the query stays fp32 in registers (two
// float4 per chunk), accumulation is fp32 (FFMA2).  Four rows in flight per group carry the same bytes as two fp32
// rows, so a hop needs half as many dependent round trips.
template <int METRIC>
__device__ __forceinline__ float2 acc8(float2 acc, const float4 &qa, const float4 &qb, const uint4 &v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const float2 qp[4] = {make_float2(qa.x, qa.y), make_float2(qa.z, qa.w), make_float2(qb.x, qb.y), make_float2(qb.z, qb.w)};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 e = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
        if (METRIC == 0) {
            const float2 d = __ffma2_rn(e, make_float2(-1.f, -1.f), qp[i]);
            acc = __ffma2_rn(d, d, acc);
        } else {
            acc = __ffma2_rn(qp[i], e, acc);
        }
    }
    return acc;
}

template <int TEAM, int LPV, int C16, int METRIC>
__device__ __forceinline__ void eval_admit_bf16(const float4 (&q)[C16][2], const uint4 *__restrict__ vec16, uint32_t d16,
                                                const uint32_t *ids, int n, bool full, float bound, uint64_t *acc,
                                                int *s_acc, int grp, int sub) {
    constexpr int NGRP = TEAM / LPV;
    constexpr int ROWS = 4;
    const uint32_t gmask = LPV == 32 ? 0xffffffffu : (((1u << LPV) - 1u) << ((threadIdx.x & 31) / LPV * LPV));
    for (int j = grp; j < n; j += ROWS * NGRP) {
        uint32_t id[ROWS];
        bool has[ROWS];
        uint4 v[ROWS][C16];
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const int jr = j + r * NGRP;
            has[r] = jr < n;
            id[r] = ids[has[r] ? jr : j];
        }
#pragma unroll
        for (int r = 0; r < ROWS; r++)
#pragma unroll
            for (int c = 0; c < C16; c++) {
                const uint32_t idx = sub + c * LPV;
                v[r][c] = (has[r] && idx < d16) ? ldg_stream_u4(vec16 + (size_t)id[r] * d16 + idx) : make_uint4(0, 0, 0, 0);
            }
        float2 p[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) p[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < C16; c++) {
            const uint32_t idx = sub + c * LPV;
            if (idx < d16) {
#pragma unroll
                for (int r = 0; r < ROWS; r++) p[r] = acc8<METRIC>(p[r], q[c][0], q[c][1], v[r][c]);
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            float sr = group_sum<LPV>(p[r].x + p[r].y, gmask);
            if (METRIC == 1) sr = 1.0f - sr;
            if (sub == 0 && has[r] && (!full || sr < bound)) acc[atomicAdd(s_acc, 1)] = make_key(sr, id[r]);
        }
    }
}

__device__ __forceinline__ bool hash_insert(uint32_t *tab, uint32_t bits, uint32_t id) {
    const uint32_t mask = (1u << bits) - 1u;
    uint32_t h = (id * 0x9E3779B1u) >> (32 - bits);
    for (;;) {
        const uint32_t old = atomicCAS(&tab[h], kEmpty, id);
        if (old == kEmpty) return true;
        if (old == id) return false;
        h = (h + 1) & mask;
    }
}

// Read-only view of the graph in HBM (device_index.cuh).
struct GraphView {
    const float4 *vec;        // [n][d4]
    const uint32_t *links0;   // [n][maxM0]
    const uint32_t *up_base;  // [n]
    const uint32_t *links_up; // [lists][maxM]
    uint32_t d4, maxM, maxM0;
    const uint4 *vec16 = nullptr;  // optional bf16 copy [n][d16], 8 elements per 128-bit chunk (storage variant)
    uint32_t d16 = 0;
    uint32_t pf = 0;               // kPf* bits: L2 prefetch policy of the traversal
    __device__ __forceinline__ const uint32_t *list(uint32_t node, int level) const {
        return level == 0 ? links0 + (size_t)node * maxM0
                          : links_up + ((size_t)__ldg(up_base + node) + (uint32_t)(level - 1)) * maxM;
    }
    __device__ __forceinline__ uint32_t list_len(int level) const { return level == 0 ? maxM0 : maxM; }
};

// Per-CTA scratch shared by the search and the construction kernels.
struct TeamCtx {
    uint64_t *buf_a, *buf_b, *acc;
    uint32_t *ids;
    float *dist;
    uint32_t *pref;   // neighbour list prefetched for the predicted next expansion
    uint32_t *hash;
    int *s_cnt, *s_next, *s_best, *s_acc, *s_pref, *s_size, *s_nr, *s_pd;
    uint32_t hash_bits;
    template <class Smem>
    __device__ __forceinline__ void bind(unsigned char *smem, const Smem &L, int *ints, uint32_t bits) {
        buf_a = (uint64_t *)(smem + L.off_buf0);
        buf_b = (uint64_t *)(smem + L.off_buf1);
        acc = (uint64_t *)(smem + L.off_acc);
        ids = (uint32_t *)(smem + L.off_ids);
        dist = (float *)(smem + L.off_dist);
        pref = (uint32_t *)(smem + L.off_pref);
        hash = (uint32_t *)(smem + L.off_hash);
        s_cnt = ints; s_next = ints + 1; s_best = ints + 2; s_acc = ints + 3; s_pref = ints + 4;
        s_size = ints + 5; s_nr = ints + 6; s_pd = ints + 7;
        hash_bits = bits;
    }
};
constexpr int kTeamInts = 8;

struct WorkCounters {
    uint32_t D = 0, H0 = 0, Hup = 0, resets = 0;
};

// Greedy descent on one upper level (hnswalg.h:1278-1303 / :1216-1238): scan ALL neighbours of the current node,
// move to the closest if it improves, repeat until no change.  argmin with lowest-slot tie-break equals the
// reference's sequential strict '<' scan.
template <int TEAM, int LPV, int CPL, int METRIC>
__device__ __forceinline__ void greedy_level(const TeamCtx &c, const float4 (&q)[CPL], const GraphView &g, int level,
                                             uint32_t &cur, float &curdist, WorkCounters &w) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int sub = tid % LPV, grp = tid / LPV;
    bool changed = true;
    while (changed) {
        changed = false;
        __syncthreads();  // previous round's reads of ids/dist/s_best are done
        const uint32_t *lst = g.list(cur, level);
        int cnt = 0;  // lists are dense: valid slots are 0..cnt-1
        for (uint32_t b0 = 0; b0 < g.maxM; b0 += TEAM) {
            uint32_t nid = kEmpty;
            if (b0 + tid < g.maxM) {
                nid = __ldg(lst + b0 + tid);
                c.ids[b0 + tid] = nid;
                if ((g.pf & kPfEarly) && nid != kEmpty) prefetch_span((const char *)(g.vec + (size_t)nid * g.d4), g.d4 * 16);
            }
            cnt += __syncthreads_count(nid != kEmpty);
        }
        if ((g.pf & (kPfGreedy | kPfEarly)) == kPfGreedy)
            prefetch_rows<TEAM>((const char *)g.vec, g.d4 * 16, c.ids, 2 * (TEAM / LPV), cnt);
        eval_list<TEAM, LPV, CPL, METRIC>(q, g.vec, g.d4, c.ids, cnt, c.dist, grp, sub);
        __syncthreads();
        w.D += cnt;
        w.Hup += 1;
        if (tid < 32) {
            float bd = 3.402823466e+38f;
            int bj = 0x7fffffff;
            for (int j = lane; j < cnt; j += 32) {
                const float dj = c.dist[j];
                if (dj < bd) { bd = dj; bj = j; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
            }
            if (lane == 0) *c.s_best = (cnt > 0 && bd < curdist) ? bj : -1;
        }
        __syncthreads();
        const int b = *c.s_best;
        if (b >= 0) {
            curdist = c.dist[b];
            cur = c.ids[b];
            changed = true;
        }
    }
    __syncthreads();
}

// Best-first beam search on one level (searchBaseLayerST<true>, hnswalg.h:309-440; the construction variant
// searchBaseLayer :225-305 has the same result set when nothing is deleted): sorted top-ef buffer with expanded
// bits, starting from (cur, curdist).  The visited table must be empty on entry.  On return the result is the
// first `size` keys of (cb ? buf_b : buf_a), closest first.
//
// Per hop (4 block barriers): [list of the node to expand: from the prefetch buffer when the prediction was right,
// else one coalesced global read] -> visited filter + compaction -> gather + distance + admission (eval_admit)
// while the last warp prefetches the list of the best other unexpanded entry -> rank merge.
//
// NB = true is the non-bare-bone variant (searchBaseLayerST<false>, hnswalg.h:324-433 without filter / stop condition):
// deleted nodes are traversed but never enter top_candidates.  They stay in the same sorted buffer with bit 30 set;
// only non-deleted entries count towards ef, the bound is the ef-th non-deleted distance, and everything behind that
// entry is dropped after each merge (the reference would stop at the first such candidate, :346-358).  The buffer
// holds bufcap entries (sized by launch_search from the rejected fraction); rejected entries beyond it lose their
// farthest members.
template <int TEAM, int LPV, int CPL, int METRIC, bool NB = false, int STORE = 0>
__device__ __forceinline__ void beam_level(const TeamCtx &c, const float4 (&q)[STORE ? (CPL + 1) / 2 * 2 : CPL],
                                           const GraphView &g, int level, uint32_t ef, uint32_t cur, float curdist,
                                           int &cb, int &size, WorkCounters &w, const uint8_t *flags = nullptr,
                                           uint32_t bufcap = 0) {
    constexpr uint32_t IDM = NB ? 0x3FFFFFFFu : kIdMask;
    constexpr uint64_t KM = NB ? 0xFFFFFFFF3FFFFFFFull : kKeyMask;
    const uint32_t cap = NB ? bufcap : ef;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = tid % LPV, grp = tid / LPV;
    const uint32_t HS = 1u << c.hash_bits;
    const uint32_t llen = g.list_len(level);
    const bool can_pref = llen <= 64;  // the prefetching warp holds the list in two registers per lane
    cb = 0;
    size = 1;
    const uint32_t ep_del = NB ? (uint32_t)(__ldg(flags + cur) & 1u) : 0u;
    int nr = ep_del ? 0 : 1;  // non-deleted entries in the buffer (NB only; uniform)
    if (tid == 0) {
        c.buf_a[0] = make_key(curdist, cur | (ep_del << 30));
        hash_insert(c.hash, c.hash_bits, cur);
        *c.s_next = 0;
        *c.s_pref = -1;
    }
    uint32_t hcount = 1;      // ids in the visited table (uniform across threads)
    uint32_t pref_node = kEmpty;  // node whose list sits in c.pref[] (uniform)
    __syncthreads();

    for (;;) {
        const int next = *c.s_next;
        if (next >= size) break;
        uint64_t *src = cb ? c.buf_b : c.buf_a, *dst = cb ? c.buf_a : c.buf_b;
        const uint32_t node = (uint32_t)src[next] & IDM;
        const bool full = NB ? (uint32_t)nr == ef : (uint32_t)size == ef;
        const float bound = full ? ord2f((uint32_t)(src[size - 1] >> 32)) : 3.402823466e+38f;
        const uint32_t *lst = g.list(node, level);
        const bool hit = can_pref && node == pref_node;
        // issue the list read before the barrier so its latency overlaps the bookkeeping
        uint32_t nid0 = kEmpty;
        if ((uint32_t)tid < llen) nid0 = hit ? c.pref[tid] : __ldg(lst + tid);
        // visited table at > 5/8 load: rebuild it from the buffer (results unchanged, see header)
        if (hcount > HS / 8 * 5) {
            __syncthreads();
            for (uint32_t i = tid; i < HS; i += TEAM) c.hash[i] = kEmpty;
            __syncthreads();
            for (int i = tid; i < size; i += TEAM) hash_insert(c.hash, c.hash_bits, (uint32_t)src[i] & IDM);
            hcount = size;
            w.resets += 1;
        }
        __syncthreads();  // (A) everyone has read s_next / src[next] / pref[]
        if (tid == 0) {
            src[next] |= (uint64_t)kExpanded;
            *c.s_next = 0x7fffffff;
        }
        // neighbour list -> unvisited ids, compacted into ids[]
        for (uint32_t b0 = 0; b0 < llen; b0 += TEAM) {
            uint32_t nid = nid0;
            if (b0) {
                nid = kEmpty;
                if (b0 + tid < llen) nid = __ldg(lst + b0 + tid);
            }
            bool isnew = false;
            if (nid != kEmpty) isnew = hash_insert(c.hash, c.hash_bits, nid);
            if ((g.pf & kPfEarly) && isnew)
                prefetch_span(STORE ? (const char *)(g.vec16 + (size_t)nid * g.d16) : (const char *)(g.vec + (size_t)nid * g.d4),
                              STORE ? g.d16 * 16 : g.d4 * 16);
            const uint32_t m = __ballot_sync(0xffffffffu, isnew);
            int basepos = 0;
            if (lane == 0 && m) basepos = atomicAdd(c.s_cnt, __popc(m));
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            if (isnew) c.ids[basepos + __popc(m & ((1u << lane) - 1u))] = nid;
        }
        // the last warp predicts the next expansion: the first other unexpanded entry of the (pre-merge) buffer,
        // and starts reading its list; the data lands in registers while everybody gathers vectors
        uint32_t pf0 = kEmpty, pf1 = kEmpty, pnode = kEmpty;
        if (can_pref && warp == TEAM / 32 - 1) {
            const int pos = next + 1 + lane;
            const bool un = pos < size && !((uint32_t)src[pos] & kExpanded);
            const uint32_t b = __ballot_sync(0xffffffffu, un);
            float pd = 3.402823466e+38f;  // no other unexpanded entry: any admitted neighbour is expanded next
            if (b) {
                const uint64_t pk = src[next + __ffs(b)];
                pnode = (uint32_t)pk & IDM;
                pd = ord2f((uint32_t)(pk >> 32));
                const uint32_t *pl = g.list(pnode, level);
                if ((uint32_t)lane < llen) pf0 = __ldg(pl + lane);
                if ((uint32_t)lane + 32 < llen) pf1 = __ldg(pl + lane + 32);
            }
            if (lane == 0) *c.s_pd = __float_as_int(pd);
        }
        __syncthreads();  // (B) ids[] complete
        const int nnew = *c.s_cnt;
        w.H0 += 1;
        w.D += nnew;
        hcount += nnew;
        if ((g.pf & (kPfRows | kPfRound1)) && !(g.pf & kPfEarly)) {
            // rows of the later gather rounds -> L2 now, so only round 1 pays the DRAM latency
            constexpr int kInFlight = STORE ? 4 * (TEAM / LPV) : 2 * (TEAM / LPV);
            const uint32_t rb = STORE ? g.d16 * 16 : g.d4 * 16;
            const char *base = STORE ? (const char *)g.vec16 : (const char *)g.vec;
            prefetch_rows<TEAM>(base, rb, c.ids, (g.pf & kPfRound1) ? 0 : kInFlight, nnew);
        }
        if (STORE == 0) {
            const bool pfl = can_pref && level == 0 && (g.pf & kPfNewBest);
            eval_admit<TEAM, LPV, CPL, METRIC, NB>(reinterpret_cast<const float4(&)[CPL]>(q), g.vec, g.d4, c.ids, nnew, full,
                                                   bound, c.acc, c.s_acc, grp, sub, flags, pfl ? g.links0 : nullptr,
                                                   g.maxM0 * 4, pfl ? __int_as_float(*c.s_pd) : 0.f);
        } else {
            constexpr int C16 = (CPL + 1) / 2;
            eval_admit_bf16<TEAM, LPV, C16, METRIC>(reinterpret_cast<const float4(&)[C16][2]>(q), g.vec16, g.d16, c.ids, nnew,
                                                    full, bound, c.acc, c.s_acc, grp, sub);
        }
        if (can_pref && warp == TEAM / 32 - 1) {
            if ((uint32_t)lane < llen) c.pref[lane] = pf0;
            if ((uint32_t)lane + 32 < llen) c.pref[lane + 32] = pf1;
            if (lane == 0) *c.s_pref = (int)pnode;
            if ((g.pf & kPfSpec) && pnode != kEmpty) {
                // speculative: the rows the predicted expansion will gather (it is the best existing candidate, so
                // even when a new neighbour overtakes it, it is usually expanded a hop or two later, while the rows
                // are still in the 126 MB L2)
                const uint32_t rb = STORE ? g.d16 * 16 : g.d4 * 16;
                const char *base = STORE ? (const char *)g.vec16 : (const char *)g.vec;
                const uint32_t hm = (1u << c.hash_bits) - 1u;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t id = h ? pf1 : pf0;
                    bool want = id != kEmpty;
                    if (want && (g.pf & kPfSpecFilter)) {  // read-only probe of the visited table
                        uint32_t hh = (id * 0x9E3779B1u) >> (32 - c.hash_bits);
                        for (;;) {
                            const uint32_t v = c.hash[hh];
                            if (v == id) { want = false; break; }
                            if (v == kEmpty) break;
                            hh = (hh + 1) & hm;
                        }
                    }
                    if (want) prefetch_span(base + (size_t)id * rb, rb);
                }
            }
        }
        __syncthreads();  // (C) acc[] complete, prefetched list stored
        const int m = *c.s_acc;
        pref_node = (uint32_t)*c.s_pref;
        int local_min = 0x7fffffff;
        if (m == 0) {
            // nothing admitted: buffer unchanged, find the next unexpanded entry after `next`
            for (int i = next + 1 + tid; i < size; i += TEAM)
                if (!((uint32_t)src[i] & kExpanded)) { local_min = i; break; }
        } else {
            // merge by rank: final position = own index + number of smaller keys in the other list
            for (int i = tid; i < size; i += TEAM) {
                const uint64_t key = src[i];
                const uint64_t km = key & KM;
                int pos = i;
                for (int j = 0; j < m; j++) pos += ((c.acc[j] & KM) < km) ? 1 : 0;
                if ((uint32_t)pos < cap) {
                    dst[pos] = key;
                    if (!((uint32_t)key & kExpanded)) local_min = min(local_min, pos);
                }
            }
            for (int j = TEAM - 1 - tid; j < m; j += TEAM) {
                const uint64_t key = c.acc[j];
                const uint64_t km = key & KM;
                int r = 0;
                for (int i = 0; i < m; i++) r += ((c.acc[i] & KM) < km) ? 1 : 0;
                int lo = 0, hi = size;  // upper bound: equal keys (impossible by construction) stay distinct
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((src[mid] & KM) <= km) lo = mid + 1; else hi = mid;
                }
                const int pos = r + lo;
                if ((uint32_t)pos < cap) {
                    dst[pos] = key;
                    local_min = min(local_min, pos);
                }
            }
            size = min((int)cap, size + m);
            cb ^= 1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_min = min(local_min, __shfl_xor_sync(0xffffffffu, local_min, o));
        if (lane == 0 && local_min != 0x7fffffff) atomicMin(c.s_next, local_min);
        if (tid == 0) { *c.s_cnt = 0; *c.s_acc = 0; }
        __syncthreads();  // (D) merged buffer and s_next visible
        if (NB && m != 0) {
            // cut the buffer behind the ef-th non-deleted entry (the new lowerBound); one warp scans 32 keys a step
            if (warp == 0) {
                const uint64_t *buf = cb ? c.buf_b : c.buf_a;
                int seen = 0, cut = size;
                for (int b0 = 0; b0 < size; b0 += 32) {
                    const int i = b0 + lane;
                    const bool live = i < size && !((uint32_t)buf[i] & 0x40000000u);
                    const uint32_t bm = __ballot_sync(0xffffffffu, live);
                    const int cntb = __popc(bm);
                    if (seen + cntb >= (int)ef) {
                        cut = b0 + (int)__fns(bm, 0, (int)ef - seen) + 1;
                        seen = (int)ef;
                        break;
                    }
                    seen += cntb;
                }
                if (lane == 0) { *c.s_size = cut; *c.s_nr = seen; }
            }
            __syncthreads();
            size = *c.s_size;
            nr = *c.s_nr;
        }
    }
}

// FULL: the row is exactly LPV * CPL 128-bit chunks (dim = 4 * LPV * CPL: 32, 64, 96, 128, 192, 256, 384, 512, 768, 1024).
// Every "chunk index < d4" test of the gathers then folds at compile time -- no predicated loads, no zeroed registers
// (12 % of the instructions of the generic kernel at dim 128).
template <int TEAM, int LPV, int CPL, int METRIC, bool NB = false, int STORE = 0, bool FULL = false>
__global__ void __launch_bounds__(TEAM, CPL <= 4 ? 1024 / TEAM : 512 / TEAM) hnsw_search_kernel(const SearchArgs p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t list_cap = p.maxM0 > p.maxM ? p.maxM0 : p.maxM;
    const SearchSmem L(p.bufcap, list_cap, p.d4, p.hash_bits);
    __shared__ int s_ints[kTeamInts];
    TeamCtx c;
    c.bind(smem, L, s_ints, p.hash_bits);
    float *qs = (float *)(smem + L.off_q);
    const uint32_t d4 = FULL ? (uint32_t)(LPV * CPL) : p.d4;
    GraphView g{p.vec, p.links0, p.up_base, p.links_up, d4, p.maxM, p.maxM0};
    g.vec16 = p.vec16;
    g.d16 = FULL ? (uint32_t)(LPV * CPL / 2) : p.d16;
    g.pf = p.pf;

    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    const uint32_t qi = blockIdx.x;
    const uint32_t HS = 1u << p.hash_bits;
    // Programmatic dependent launch: batches are independent, so the next batch's grid may start filling SMs as soon
    // as every CTA of this one has been scheduled -- the tail of batch i (SMs draining their last queries) overlaps the
    // head of batch i+1.  Outputs are written only after griddepcontrol.wait (= the previous grid has completed and
    // flushed), so stream order of everything a consumer can observe is unchanged.  No-ops without the launch attribute.
    asm volatile("griddepcontrol.launch_dependents;");

    // ---- stage the query (rows of Q are only 4-byte aligned when dim % 4 != 0) and clear the visited table ----
    for (uint32_t i = tid; i < d4 * 4; i += TEAM) qs[i] = i < p.dim ? p.Q[(size_t)qi * p.dim + i] : 0.f;
    for (uint32_t i = tid; i < HS; i += TEAM) c.hash[i] = kEmpty;
    if (tid == 0) { *c.s_cnt = 0; *c.s_acc = 0; *c.s_next = 0; }
    __syncthreads();
    WorkCounters w;
    uint32_t cur = p.entry;
    float curdist;
    {   // ---- searchKnn prologue: distance to the entry point, greedy descent on levels maxlevel..1 (fp32 rows) ----
        float4 q[CPL];
#pragma unroll
        for (int cc = 0; cc < CPL; cc++) {
            const uint32_t idx = sub + cc * LPV;
            q[cc] = idx < d4 ? ((const float4 *)qs)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid == 0) c.ids[0] = cur;
        __syncthreads();
        eval_list<TEAM, LPV, CPL, METRIC>(q, g.vec, d4, c.ids, 1, c.dist, grp, sub);
        __syncthreads();
        curdist = c.dist[0];
        w.D += 1;
        for (int level = p.maxlevel; level > 0; --level) greedy_level<TEAM, LPV, CPL, METRIC>(c, q, g, level, cur, curdist, w);
        __syncthreads();
    }

    // ---- searchBaseLayerST on level 0 ----
    int cb, size;
    if (STORE == 0) {
        float4 q[CPL];
#pragma unroll
        for (int cc = 0; cc < CPL; cc++) {
            const uint32_t idx = sub + cc * LPV;
            q[cc] = idx < d4 ? ((const float4 *)qs)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        beam_level<TEAM, LPV, CPL, METRIC, NB, 0>(c, q, g, 0, p.ef, cur, curdist, cb, size, w, p.flags, p.bufcap);
    } else {
        constexpr int C16 = (CPL + 1) / 2;
        float4 q[C16 * 2];  // chunk idx covers query elements 8*idx .. 8*idx+7
#pragma unroll
        for (int cc = 0; cc < C16; cc++) {
            const uint32_t idx = sub + cc * LPV;
            q[2 * cc] = 2 * idx < d4 ? ((const float4 *)qs)[2 * idx] : make_float4(0.f, 0.f, 0.f, 0.f);
            q[2 * cc + 1] = 2 * idx + 1 < d4 ? ((const float4 *)qs)[2 * idx + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        beam_level<TEAM, LPV, CPL, METRIC, false, 1>(c, q, g, 0, p.ef, cur, curdist, cb, size, w, nullptr, p.bufcap);
    }

    // ---- epilogue: first k entries are the result, closest first (hnswalg.h:1315-1322) ----
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint64_t *res = cb ? c.buf_b : c.buf_a;
    if (STORE == 1) {
        // bf16 traversal: re-evaluate the final buffer with the fp32 rows and re-sort it, so distances are the fp32
        // ones and near-ties at the k-th place are decided exactly (ef * 4d extra bytes per query)
        uint64_t *other = cb ? c.buf_a : c.buf_b;
        uint32_t *rid = (uint32_t *)other;
        float *rd = (float *)(rid + p.bufcap);
        for (int i = tid; i < size; i += TEAM) rid[i] = (uint32_t)res[i] & kIdMask;
        __syncthreads();
        {
            float4 q[CPL];
#pragma unroll
            for (int cc = 0; cc < CPL; cc++) {
                const uint32_t idx = sub + cc * LPV;
                q[cc] = idx < d4 ? ((const float4 *)qs)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            eval_list<TEAM, LPV, CPL, METRIC>(q, g.vec, d4, rid, size, rd, grp, sub);
        }
        __syncthreads();
        uint64_t *srt = const_cast<uint64_t *>(res);  // the traversal keys are no longer needed
        for (int i = tid; i < size; i += TEAM) {
            const uint64_t key = make_key(rd[i], rid[i]);
            int r = 0;
            for (int j = 0; j < size; j++) r += make_key(rd[j], rid[j]) < key ? 1 : 0;
            srt[r] = key;
        }
        __syncthreads();
        w.D += size;
    }
    if (NB) {
        // results are the non-deleted entries only: compact them to the front (same buffer, other half as scratch)
        uint64_t *tmp = cb ? c.buf_a : c.buf_b;
        if (tid < 32) {
            int outn = 0;
            for (int b0 = 0; b0 < size; b0 += 32) {
                const int i = b0 + tid;
                const bool live = i < size && !((uint32_t)res[i] & 0x40000000u);
                const uint32_t bm = __ballot_sync(0xffffffffu, live);
                if (live) tmp[outn + __popc(bm & ((1u << tid) - 1u))] = res[i];
                outn += __popc(bm);
            }
            if (tid == 0) *c.s_size = outn;
        }
        __syncthreads();
        res = tmp;
        size = *c.s_size;
    }
    for (uint32_t j = tid; j < p.k; j += TEAM) {
        uint64_t lab = 0xFFFFFFFFFFFFFFFFull;
        float dj = __int_as_float(0x7f800000);
        if (j < (uint32_t)size) {
            const uint64_t key = res[j];
            lab = __ldg(p.labels + ((uint32_t)key & (NB ? 0x3FFFFFFFu : kIdMask)));
            dj = ord2f((uint32_t)(key >> 32));
        }
        p.out_labels[(size_t)qi * p.k + j] = lab;
        p.out_dists[(size_t)qi * p.k + j] = dj;
    }
    if (tid == 0) {
        if (p.out_counts) p.out_counts[qi] = min((uint32_t)size, p.k);
        if (p.out_work) {
            uint32_t *wo = p.out_work + (size_t)qi * 4;
            wo[0] = w.D; wo[1] = w.H0; wo[2] = w.Hup; wo[3] = w.resets;
        }
    }
}

// k-way merge of per-shard results: one warp per query over the candidates of one GROUP of shards.
// Shard s holds its rows at labels_in + s*lstride / dists_in + s*dstride (elements), so per-shard blocks may be packed
// [labels | dists] and exchanged with ONE all_gather.  blockIdx.y selects the group of `group` consecutive shards and
// the output list at labels_out + blockIdx.y*ols / dists_out + blockIdx.y*ods (merge_launch.cuh builds the tree).
// Candidates are ranked by (dist, label, position).  This version reads every pair from global memory (total^2
// dependent loads per warp): only used beyond 4096 candidates per warp.
static __global__ void merge_topk_kernel(const uint64_t *__restrict__ labels_in, const float *__restrict__ dists_in,
                                         size_t lstride, size_t dstride, uint32_t shards, uint32_t group, uint32_t nq,
                                         uint32_t k, uint64_t *__restrict__ labels_out, float *__restrict__ dists_out,
                                         size_t ols, size_t ods) {
    const uint32_t qi = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const uint32_t s0 = blockIdx.y * group;
    const uint32_t total = min(group, shards - s0) * k;
    labels_in += (size_t)s0 * lstride;
    dists_in += (size_t)s0 * dstride;
    labels_out += (size_t)blockIdx.y * ols;
    dists_out += (size_t)blockIdx.y * ods;
    for (uint32_t c = lane; c < total; c += 32) {
        const uint32_t s = c / k, j = c % k;
        const float dc = dists_in[s * dstride + (size_t)qi * k + j];
        const uint64_t lc = labels_in[s * lstride + (size_t)qi * k + j];
        uint32_t rank = 0;
        for (uint32_t o = 0; o < total && rank < k; o++) {
            const uint32_t so = o / k, jo = o % k;
            const float d2 = dists_in[so * dstride + (size_t)qi * k + jo];
            const uint64_t l2 = labels_in[so * lstride + (size_t)qi * k + jo];
            rank += (d2 < dc || (d2 == dc && (l2 < lc || (l2 == lc && o < c)))) ? 1u : 0u;
        }
        if (rank < k) {
            labels_out[(size_t)qi * k + rank] = lc;
            dists_out[(size_t)qi * k + rank] = dc;
        }
    }
}

// Same contract, candidates staged once into shared memory (coalesced per-shard row reads), ranks computed from
// shared memory with broadcast reads and an early exit once a candidate's rank reaches k.  The global-memory version
// is latency-bound and, on a high-priority exchange stream, crowds the search kernel out of the SMs.  Dynamic shared
// memory: warps_per_cta * group * k * 12 bytes.
// SORTED: every shard's row is in ascending distance order (search rows are: closest first, padding last), so the
// candidates of a shard that rank before a given one form a prefix of its row -- the scan of a shard stops at its first
// larger distance: at most rank + shards steps per candidate instead of up to shards * k.
template <bool SORTED>
static __global__ void merge_topk_smem_kernel(const uint64_t *__restrict__ labels_in, const float *__restrict__ dists_in,
                                              size_t lstride, size_t dstride, uint32_t shards, uint32_t group,
                                              uint32_t nq, uint32_t k, uint64_t *__restrict__ labels_out,
                                              float *__restrict__ dists_out, size_t ols, size_t ods) {
    extern __shared__ __align__(16) unsigned char merge_smem[];
    const uint32_t warps = blockDim.x / 32, w = threadIdx.x / 32;
    const uint32_t qi = blockIdx.x * warps + w;
    const int lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const uint32_t s0 = blockIdx.y * group;
    const uint32_t cap = min(group, shards) * k;            // per-warp slice of the shared arrays
    const uint32_t total = min(group, shards - s0) * k;
    labels_in += (size_t)s0 * lstride;
    dists_in += (size_t)s0 * dstride;
    labels_out += (size_t)blockIdx.y * ols;
    dists_out += (size_t)blockIdx.y * ods;
    uint64_t *sl = reinterpret_cast<uint64_t *>(merge_smem) + (size_t)w * cap;
    float *sd = reinterpret_cast<float *>(merge_smem + (size_t)warps * cap * 8) + (size_t)w * cap;
    for (uint32_t c = lane; c < total; c += 32) {
        const uint32_t s = c / k, j = c % k;
        sl[c] = labels_in[s * lstride + (size_t)qi * k + j];
        sd[c] = dists_in[s * dstride + (size_t)qi * k + j];
    }
    __syncwarp();
    for (uint32_t c = lane; c < total; c += 32) {
        const float dc = sd[c];
        const uint64_t lc = sl[c];
        uint32_t rank = 0;
        if (SORTED) {
            for (uint32_t s0_ = 0; s0_ < total && rank < k; s0_ += k) {
                for (uint32_t o = s0_; o < s0_ + k; o++) {
                    const float d2 = sd[o];
                    if (d2 < dc) {
                        rank++;
                    } else if (d2 == dc) {  // ties on distance are rare: only then look at the label
                        const uint64_t l2 = sl[o];
                        rank += (l2 < lc || (l2 == lc && o < c)) ? 1u : 0u;
                    } else {
                        break;  // the rest of this shard's row is farther
                    }
                }
            }
        } else {
            for (uint32_t o = 0; o < total; o++) {
                const float d2 = sd[o];
                if (d2 < dc) {
                    rank++;
                } else if (d2 == dc) {  // ties on distance are rare: only then look at the label
                    const uint64_t l2 = sl[o];
                    rank += (l2 < lc || (l2 == lc && o < c)) ? 1u : 0u;
                }
                if (rank >= k) break;
            }
        }
        if (rank < k) {
            labels_out[(size_t)qi * k + rank] = lc;
            dists_out[(size_t)qi * k + rank] = dc;
        }
    }
}

}  // namespace b200

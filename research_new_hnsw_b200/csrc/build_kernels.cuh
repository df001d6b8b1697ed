// build_kernels.cuh -- device code of the batched GPU graph construction and of updatePoint (see build.cu for the
// reference lines each kernel replaces).  Included by the translation units that instantiate one family each
// (build_insert.cu, build_update.cu, build_nb.cu) so that they compile in parallel; build.cu holds the host logic.
#pragma once
#include <algorithm>
#include <vector>

#include "hnsw_index.cuh"

namespace b200 {

// Threads per CTA of the build kernels.  64-thread teams keep 16 construction searches resident per SM (64 registers
// per thread) instead of 8: measured at C5 with a 2048-slot visited table, construction search 924 ms vs 1179 ms
// (128 threads, 4096 slots), link 244 vs 252 ms, reverse 63 vs 77 ms (gpurun_out/s2_build_team.log).
#ifndef B200_BUILD_TEAM
#define B200_BUILD_TEAM 64
#endif
constexpr int kTeam = B200_BUILD_TEAM;
constexpr uint32_t kCapIn = 32;  // incoming reverse edges kept per list and batch
#ifndef B200_PRUNE_AHEAD
#define B200_PRUNE_AHEAD 8
#endif
constexpr int kPruneAhead = B200_PRUNE_AHEAD;  // heuristic_prune: candidate rows prefetched into L2 this many rounds ahead

struct BuildArgs {
    float4 *vec;
    uint32_t *links0, *up_base, *links_up;
    const int32_t *plevel;     // [n] element levels
    uint64_t *cand;            // [lists][efc] sorted keys
    uint32_t *cand_cnt;        // [lists]
    const uint32_t *list_off;  // [batch] slot of the point's level-0 list; level l at slot + l
    const uint32_t *list_point, *list_level;  // [lists]
    uint32_t *incnt;           // [cap + up_lists_cap]
    uint64_t *incoming;        // [cap + up_lists_cap][kCapIn]
    uint32_t *aff_node, *aff_level, *aff_count;
    unsigned long long *work;  // [8] D, H0, Hup, resets, dropped reverse edges (atomicAdd)
    uint32_t cap, first, batch, lists;
    uint32_t entry;
    int32_t maxlevel;
    uint32_t d4, maxM, maxM0, M, efc, hash_bits;
    // update mode only (kernels instantiated with UPD = true re-link EXISTING points: repairConnectionsForUpdate,
    // hnswalg.h:1075-1139); kept at the end so the insert-mode kernels see the layout they were tuned with
    const uint32_t *batch_ids;  // [batch] ids of the points of this batch (insert mode: first + b)
    const uint8_t *flags;       // [n] delete marks; only read by the NB instantiations (elements marked deleted exist)
    uint32_t pf;                // GraphView::pf of the construction searches (L2 prefetch policy)
};

// Implemented one family per translation unit (they compile in parallel): insert / update mode, and the same two with
// NB = true, used while elements are marked deleted -- searchBaseLayer traverses a deleted element but never puts it
// into top_candidates (hnswalg.h:291-292), so it cannot become a neighbour of the point being linked.
int build_run_batch_insert(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st);
int build_run_batch_update(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st);
int build_run_batch_insert_nb(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st);
int build_run_batch_update_nb(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st);
int build_run_update_phase1(int metric, const BuildArgs &a, uint32_t *newlists, cudaStream_t st);

// B200HNSW_BUILD_PROFILE=1: CUDA events around every build launch, summed per kernel by flush() (diagnostic only)
struct BuildProfile {
    std::vector<cudaEvent_t> ev;
    void mark(cudaStream_t st) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
    }
};
BuildProfile *build_profile();  // build.cu; nullptr unless profiling is on

template <int LPV, int CPL>
__device__ __forceinline__ void load_row(float4 (&v)[CPL], const float4 *row, uint32_t d4, int sub) {
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        const uint32_t idx = sub + c * LPV;
        v[c] = idx < d4 ? __ldg(row + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Construction search for one new point: CTA b handles point first + b on all of its levels.
// FULL: rows are exactly LPV * CPL 128-bit chunks, so every "chunk index < d4" test of the gathers folds at compile
// time (same device as hnsw_search_kernel's FULL instantiation).
template <int LPV, int CPL, int METRIC, bool UPD, bool NB, bool FULL = false>
__global__ void __launch_bounds__(kTeam) build_search_kernel(const BuildArgs p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t list_cap = p.maxM0 > p.maxM ? p.maxM0 : p.maxM;
    const uint32_t bufcap = NB ? 2 * p.efc : p.efc;
    const uint32_t d4 = FULL ? (uint32_t)(LPV * CPL) : p.d4;
    const SearchSmem L(bufcap, list_cap, d4, p.hash_bits);
    __shared__ int s_ints[kTeamInts];
    TeamCtx c;
    c.bind(smem, L, s_ints, p.hash_bits);
    GraphView g{p.vec, p.links0, p.up_base, p.links_up, d4, p.maxM, p.maxM0};
    g.pf = p.pf;

    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    const uint32_t pid = UPD ? p.batch_ids[blockIdx.x] : p.first + blockIdx.x;
    const uint32_t HS = 1u << p.hash_bits;
    const int plevel = p.plevel[pid];

    float4 q[CPL];
    load_row<LPV, CPL>(q, p.vec + (size_t)pid * d4, d4, sub);
    if (tid == 0) { *c.s_cnt = 0; *c.s_acc = 0; *c.s_next = 0; c.ids[0] = p.entry; }
    __syncthreads();
    WorkCounters w;
    uint32_t cur = p.entry;
    eval_list<kTeam, LPV, CPL, METRIC>(q, g.vec, d4, c.ids, 1, c.dist, grp, sub);
    __syncthreads();
    float curdist = c.dist[0];
    w.D += 1;
    for (int level = p.maxlevel; level > plevel; --level) greedy_level<kTeam, LPV, CPL, METRIC>(c, q, g, level, cur, curdist, w);
    const uint32_t slot0 = p.list_off[blockIdx.x];
    for (int level = min(plevel, p.maxlevel); level >= 0; --level) {
        __syncthreads();
        for (uint32_t i = tid; i < HS; i += kTeam) c.hash[i] = kEmpty;
        __syncthreads();
        int cb, size;
        beam_level<kTeam, LPV, CPL, METRIC, NB>(c, q, g, level, p.efc, cur, curdist, cb, size, w, p.flags, bufcap);
        const uint64_t *res = cb ? c.buf_b : c.buf_a;
        if (NB) {  // candidates are the non-deleted entries only: compact them to the front of the other buffer
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
        uint64_t *out = p.cand + (size_t)(slot0 + level) * p.efc;
        for (int j = tid; j < size; j += kTeam) out[j] = res[j] & kKeyMask;
        if (tid == 0) p.cand_cnt[slot0 + level] = (uint32_t)size;
        // next level starts from the closest candidate (= selectedNeighbors.back(), hnswalg.h:524,629: the closest
        // candidate always survives the heuristic)
        if (size > 0) {
            cur = (uint32_t)res[0] & (NB ? 0x3FFFFFFFu : kIdMask);
            curdist = ord2f((uint32_t)(res[0] >> 32));
        }
    }
    if (tid == 0) {
        atomicAdd(p.work + 0, (unsigned long long)w.D); atomicAdd(p.work + 1, (unsigned long long)w.H0);
        atomicAdd(p.work + 2, (unsigned long long)w.Hup); atomicAdd(p.work + 3, (unsigned long long)w.resets);
    }
}

// Distances from the register-resident vector q to one or two rows, computed by ONE group of LPV lanes (same arithmetic
// and summation order as eval_list, so both give bit-identical values).
template <int LPV, int CPL, int METRIC>
__device__ __forceinline__ void group_dist2(const float4 (&q)[CPL], const float4 *__restrict__ ra,
                                            const float4 *__restrict__ rb, bool has2, uint32_t d4, int sub,
                                            uint32_t gmask, float &sa, float &sb) {
    float4 va[CPL], vb[CPL];
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        const uint32_t idx = sub + c * LPV;
        va[c] = idx < d4 ? __ldg(ra + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        const uint32_t idx = sub + c * LPV;
        vb[c] = (has2 && idx < d4) ? __ldg(rb + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    sa = group_sum<LPV>(pa.x + pa.y, gmask);
    sb = group_sum<LPV>(pb.x + pb.y, gmask);
    if (METRIC == 1) { sa = 1.0f - sa; sb = 1.0f - sb; }
}

// getNeighborsByHeuristic2 (hnswalg.h:443-483) for one candidate list sorted closest-first: accept c iff every already
// accepted r has dist(r, c) >= dist(base, c); stop at Mlimit.  Selected keys end up in sel[0..ns), their ids in ids[].
//
// The reference's loop is sequential in the candidates; here a WINDOW of W = kTeam / LPV candidates is in flight, one
// per lane group, with the same decisions:
//   A. every group holds its candidate's vector in registers and checks it against the rows accepted BEFORE the window,
//      two at a time, stopping at the first row that is closer to the candidate than the base point is (the sequential
//      loop evaluated every accepted row for every candidate, one block barrier per candidate);
//   B. the window is resolved in order: the first candidate that survived is accepted; the later survivors of the window
//      are checked against it (one more row each) before the next one is looked at.  Block barriers: one per window plus
//      one per ACCEPTED candidate (<= Mlimit per list) instead of three per candidate.
// The rows of the next window are prefetched into L2 while the current one is resolved.
template <int LPV, int CPL, int METRIC>
__device__ __forceinline__ int heuristic_prune(const GraphView &g, const uint64_t *cand, int n, int Mlimit,
                                               uint64_t *sel, uint32_t *ids, float *dist, uint32_t &evals) {
    constexpr int W = kTeam / LPV;
    __shared__ uint32_t s_bad[W];
    __shared__ uint32_t s_ev;
    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    const uint32_t gmask = LPV == 32 ? 0xffffffffu : (((1u << LPV) - 1u) << ((threadIdx.x & 31) / LPV * LPV));
    if (n < Mlimit) {  // hnswalg.h:446-448
        for (int j = tid; j < n; j += kTeam) { sel[j] = cand[j]; ids[j] = (uint32_t)cand[j] & kIdMask; }
        __syncthreads();
        return n;
    }
    const uint32_t row_bytes = g.d4 * 16, lpr = (row_bytes + kPfLine - 1) / kPfLine;
    auto prefetch_window = [&](int first) {
        const int cnt = min(W, n - first);
        for (uint32_t i = tid; i < (uint32_t)max(cnt, 0) * lpr; i += kTeam)
            prefetch_l2((const char *)(g.vec + (size_t)((uint32_t)cand[first + i / lpr] & kIdMask) * g.d4) + (i % lpr) * kPfLine);
    };
    prefetch_window(0);
    if (tid == 0) s_ev = 0;  // (ordered before its use by the barriers of the first window)
    int ns = 0;
    uint32_t my_evals = 0;
    for (int base = 0; base < n && ns < Mlimit; base += W) {
        const int wcount = min(W, n - base);
        const bool has = grp < wcount;
        const uint64_t key = has ? cand[base + grp] : 0;
        const float dq = ord2f((uint32_t)(key >> 32));
        float4 v[CPL];
        if (has) load_row<LPV, CPL>(v, g.vec + (size_t)((uint32_t)key & kIdMask) * g.d4, g.d4, sub);
        prefetch_window(base + W);
        // ---- A: against the rows accepted before this window ----
        bool bad = !has;
        for (int r = 0; r < ns && !bad; r += 2) {
            const bool two = r + 1 < ns;
            float da, db;
            group_dist2<LPV, CPL, METRIC>(v, g.vec + (size_t)ids[r] * g.d4, g.vec + (size_t)ids[two ? r + 1 : r] * g.d4, two,
                                          g.d4, sub, gmask, da, db);
            my_evals += two ? 2u : 1u;
            bad = da < dq || (two && db < dq);
        }
        if (sub == 0) s_bad[grp] = bad ? 1u : 0u;
        __syncthreads();
        // ---- B: resolve the window in candidate order ----
        for (int w = 0; w < wcount && ns < Mlimit; w++) {
            if (s_bad[w]) continue;  // (uniform: every thread reads the same word)
            const uint64_t kw = cand[base + w];
            const uint32_t idw = (uint32_t)kw & kIdMask;
            if (tid == 0) { sel[ns] = kw; ids[ns] = idw; }
            ns++;
            if (w + 1 < wcount && ns < Mlimit) {
                if (has && grp > w && !bad) {
                    float da, db;
                    group_dist2<LPV, CPL, METRIC>(v, g.vec + (size_t)idw * g.d4, g.vec + (size_t)idw * g.d4, false, g.d4, sub,
                                                  gmask, da, db);
                    my_evals += 1u;
                    if (da < dq) {
                        bad = true;
                        if (sub == 0) s_bad[grp] = 1u;
                    }
                }
                __syncthreads();  // the later survivors' verdicts before the next candidate of the window is looked at
            }
        }
        __syncthreads();  // ids[] / sel[] of this window are visible, s_bad[] may be rewritten
    }
    if (sub == 0 && my_evals) atomicAdd(&s_ev, my_evals);  // one count per group
    __syncthreads();
    evals += s_ev;
    return ns;
}

struct LinkSmem {
    uint32_t off_sel, off_raw, off_srt, off_ids, off_dist, total;
    __host__ __device__ explicit LinkSmem(uint32_t cap) {
        uint32_t o = 0;
        off_sel = o; o += cap * 8;
        off_raw = o; o += cap * 8;
        off_srt = o; o += cap * 8;
        off_ids = o; o += cap * 4;
        off_dist = o; o += cap * 4;
        total = o;
    }
};

__device__ __forceinline__ uint32_t list_id(const BuildArgs &p, uint32_t node, uint32_t level) {
    return level == 0 ? node : p.cap + p.up_base[node] + level - 1;
}
__device__ __forceinline__ uint32_t *list_ptr(const BuildArgs &p, uint32_t node, uint32_t level) {
    return level == 0 ? p.links0 + (size_t)node * p.maxM0
                      : p.links_up + ((size_t)p.up_base[node] + level - 1) * p.maxM;
}

// One CTA per (new point, level): prune the candidates to M, write the forward list, stage the reverse edges.
template <int LPV, int CPL, int METRIC, bool UPD>
__global__ void __launch_bounds__(kTeam) build_link_kernel(const BuildArgs p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t capc = max(p.maxM0, p.maxM) + kCapIn;
    const LinkSmem L(capc);
    uint64_t *sel = (uint64_t *)(smem + L.off_sel);
    uint32_t *ids = (uint32_t *)(smem + L.off_ids);
    float *dist = (float *)(smem + L.off_dist);
    GraphView g{p.vec, p.links0, p.up_base, p.links_up, p.d4, p.maxM, p.maxM0};
    const int tid = threadIdx.x;
    const uint32_t slot = blockIdx.x;
    const uint32_t pid = p.list_point[slot], level = p.list_level[slot];
    int n = (int)p.cand_cnt[slot];
    uint64_t *cand = p.cand + (size_t)slot * p.efc;
    if constexpr (UPD) {
        // the point is already in the graph, so its own search finds it: drop it from the candidates
        // (filteredTopCandidates, hnswalg.h:1117-1123) by closing the gap in the sorted list
        __shared__ int s_self;
        if (tid == 0) s_self = n;
        __syncthreads();
        for (int j = tid; j < n; j += kTeam)
            if (((uint32_t)cand[j] & kIdMask) == pid) s_self = j;
        __syncthreads();
        const int self = s_self;
        if (self < n) {
            for (int base = self; base < n - 1; base += kTeam) {
                const int j = base + tid;
                const uint64_t v = j < n - 1 ? cand[j + 1] : 0;
                __syncthreads();
                if (j < n - 1) cand[j] = v;
                __syncthreads();
            }
            n--;
        }
        if (n == 0) return;  // nothing but the point itself on this level: its links stay (hnswalg.h:1127)
    }
    uint32_t evals = 0;
    const int ns = heuristic_prune<LPV, CPL, METRIC>(g, cand, n, (int)p.M, sel, ids, dist, evals);
    uint32_t *mine = list_ptr(p, pid, level);
    if constexpr (UPD) {  // the old forward list is replaced, not extended
        const int Mcur = (int)(level ? p.maxM : p.maxM0);
        for (int j = ns + tid; j < Mcur; j += kTeam) mine[j] = kEmpty;
    }
    for (int j = tid; j < ns; j += kTeam) {
        const uint32_t r = ids[j];
        mine[j] = r;
        const uint32_t lid = list_id(p, r, level);
        const uint32_t s = atomicAdd(p.incnt + lid, 1u);
        if (s == 0) {
            const uint32_t pos = atomicAdd(p.aff_count, 1u);
            p.aff_node[pos] = r;
            p.aff_level[pos] = level;
        }
        if (s < kCapIn) p.incoming[(size_t)lid * kCapIn + s] = (sel[j] & 0xFFFFFFFF00000000ull) | pid;
        else atomicAdd(p.work + 4, 1ull);  // more than kCapIn new points chose r in ONE batch: edge r -> pid not offered
    }
    if (tid == 0) atomicAdd(p.work + 0, (unsigned long long)evals);
}

// One touched list: append the incoming new points while there is room, otherwise re-run the heuristic over
// existing + incoming neighbours (distances to this node) and rewrite the list.  Every branch is CTA-uniform.
template <int LPV, int CPL, int METRIC, bool UPD>
__device__ __forceinline__ void reverse_one(const BuildArgs &p, const GraphView &g, uint32_t item, unsigned char *smem) {
    const uint32_t capc = max(p.maxM0, p.maxM) + kCapIn;
    const LinkSmem L(capc);
    uint64_t *sel = (uint64_t *)(smem + L.off_sel);
    uint64_t *raw = (uint64_t *)(smem + L.off_raw);
    uint64_t *srt = (uint64_t *)(smem + L.off_srt);
    uint32_t *ids = (uint32_t *)(smem + L.off_ids);
    float *dist = (float *)(smem + L.off_dist);
    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    const uint32_t node = p.aff_node[item], level = p.aff_level[item];
    const uint32_t lid = list_id(p, node, level);
    int t = (int)min(p.incnt[lid], kCapIn);
    const int Mcur = (int)(level ? p.maxM : p.maxM0);
    uint32_t *lst = list_ptr(p, node, level);
    int deg = 0;
    for (int b0 = 0; b0 < Mcur; b0 += kTeam) {
        uint32_t v = kEmpty;
        if (b0 + tid < Mcur) { v = lst[b0 + tid]; ids[b0 + tid] = v; }
        deg += __syncthreads_count(v != kEmpty);
    }
    for (int j = tid; j < t; j += kTeam) raw[deg + j] = p.incoming[(size_t)lid * kCapIn + j];
    __syncthreads();
    if (tid == 0) p.incnt[lid] = 0;  // ready for the next batch
    if constexpr (UPD) {
        // a re-linked point may already be a neighbour of this node (is_cur_c_present, hnswalg.h:566-580): keep the
        // existing edge, drop the incoming duplicate
        __shared__ int s_keep;
        if (tid == 0) {
            int keep = 0;
            for (int j = 0; j < t; j++) {
                const uint64_t key = raw[deg + j];
                bool present = false;
                for (int i = 0; i < deg; i++) present |= ids[i] == (uint32_t)key;
                if (!present) raw[deg + keep++] = key;
            }
            s_keep = keep;
        }
        __syncthreads();
        t = s_keep;
        if (t == 0) return;
    }
    if (deg + t <= Mcur) {
        // room for all (hnswalg.h:586-588); ordered by new id so the list does not depend on atomic arrival order
        for (int j = tid; j < t; j += kTeam) {
            const uint32_t id = (uint32_t)raw[deg + j];
            int r = 0;
            for (int i = 0; i < t; i++) r += ((uint32_t)raw[deg + i] < id) ? 1 : 0;
            lst[deg + r] = id;
        }
        return;
    }
    // full: candidates = existing (distances to this node evaluated now, :597-601) + incoming
    float4 q[CPL];
    load_row<LPV, CPL>(q, p.vec + (size_t)node * p.d4, p.d4, sub);
    eval_list<kTeam, LPV, CPL, METRIC>(q, g.vec, p.d4, ids, deg, dist, grp, sub);
    __syncthreads();
    for (int j = tid; j < deg; j += kTeam) raw[j] = make_key(dist[j], ids[j]);
    __syncthreads();
    const int n = deg + t;
    for (int j = tid; j < n; j += kTeam) {  // rank sort, closest first
        const uint64_t key = raw[j];
        int r = 0;
        for (int i = 0; i < n; i++) r += (raw[i] < key || (raw[i] == key && i < j)) ? 1 : 0;
        srt[r] = key;
    }
    __syncthreads();
    uint32_t evals = (uint32_t)deg;
    const int ns = heuristic_prune<LPV, CPL, METRIC>(g, srt, n, Mcur, sel, ids, dist, evals);
    for (int j = tid; j < Mcur; j += kTeam) lst[j] = j < ns ? ids[j] : kEmpty;
    if (tid == 0) atomicAdd(p.work + 0, (unsigned long long)evals);
}

// Persistent grid over the lists touched by this batch; their number is read from device memory (aff_count), so the
// host never waits for it.
template <int LPV, int CPL, int METRIC, bool UPD>
__global__ void __launch_bounds__(kTeam) build_reverse_kernel(const BuildArgs p) {
    extern __shared__ __align__(16) unsigned char smem[];
    GraphView g{p.vec, p.links0, p.up_base, p.links_up, p.d4, p.maxM, p.maxM0};
    const uint32_t n_aff = *p.aff_count;
    for (uint32_t item = blockIdx.x; item < n_aff; item += gridDim.x) {
        __syncthreads();  // the previous item's shared arrays are no longer read
        reverse_one<LPV, CPL, METRIC, UPD>(p, g, item, smem);
    }
}

template <int LPV, int CPL, int METRIC, bool UPD, bool NB>
static int run_batch(const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st) {
    static bool configured[16] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (d < 16 && !configured[d]) {
        cudaFuncAttributes fa;
        int optin = 0;
        B200_CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d));
        B200_CUDA_OK(cudaFuncGetAttributes(&fa, build_search_kernel<LPV, CPL, METRIC, UPD, NB>));
        B200_CUDA_OK(cudaFuncSetAttribute(build_search_kernel<LPV, CPL, METRIC, UPD, NB>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        if constexpr (!UPD && !NB)
            B200_CUDA_OK(cudaFuncSetAttribute(build_search_kernel<LPV, CPL, METRIC, UPD, NB, true>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        configured[d] = true;
    }
    BuildProfile *prof = build_profile();
    if (prof) prof->mark(st);
    if constexpr (!UPD && !NB) {  // the bulk-build family also exists with the row length folded in
        if (a.d4 == (uint32_t)(LPV * CPL))
            build_search_kernel<LPV, CPL, METRIC, UPD, NB, true><<<a.batch, kTeam, smem_search, st>>>(a);
        else
            build_search_kernel<LPV, CPL, METRIC, UPD, NB><<<a.batch, kTeam, smem_search, st>>>(a);
    } else {
        build_search_kernel<LPV, CPL, METRIC, UPD, NB><<<a.batch, kTeam, smem_search, st>>>(a);
    }
    if (prof) prof->mark(st);
    build_link_kernel<LPV, CPL, METRIC, UPD><<<a.lists, kTeam, smem_link, st>>>(a);
    if (prof) prof->mark(st);
    // at most M lists are touched per (point, level); the grid is sized for the GPU (16 resident CTAs per SM), not for
    // that bound
    const unsigned rev_grid = (unsigned)std::min<size_t>((size_t)a.lists * a.M, (size_t)148 * 16);
    build_reverse_kernel<LPV, CPL, METRIC, UPD><<<rev_grid, kTeam, smem_link, st>>>(a);
    if (prof) prof->mark(st);
    B200_CUDA_OK(cudaMemsetAsync(a.aff_count, 0, 4, st));
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- updatePoint, first phase (hnswalg.h:1009-1069) --------------------------------------------------------------
// For every level of the moved point `id`: sCand = {id} + its neighbours + their neighbours; every neighbour `neigh`
// keeps the ef_construction candidates of sCand \ {neigh} closest to ITSELF and re-prunes its list from them with the
// heuristic (Mcurmax).  All lists are read as they were before the update (the reference builds sCand before it
// rewrites anything), so the phase is three launches: claim -> prune into scratch -> apply.
//
// update_claim_kernel: one CTA per (point, level) list of the batch.  When two points of one batch share a neighbour,
// the one with the larger slot re-prunes it (atomicMax on the per-list word that build_link_kernel uses later as its
// incoming counter; it is zero between batches and update_apply_kernel zeroes it again).
static __global__ void update_claim_kernel(const BuildArgs p) {
    const uint32_t slot = blockIdx.x;
    const uint32_t id = p.list_point[slot], level = p.list_level[slot];
    const uint32_t Mcur = level ? p.maxM : p.maxM0;
    const uint32_t *l1 = list_ptr(p, id, level);
    for (uint32_t j = threadIdx.x; j < Mcur; j += blockDim.x) {
        const uint32_t neigh = l1[j];
        if (neigh != kEmpty) atomicMax(p.incnt + list_id(p, neigh, level), slot + 1);
    }
}

struct UpdSmem {
    uint32_t off_bufa, off_bufb, off_acc, off_sel, off_ids, off_dist, total;
    __host__ __device__ UpdSmem(uint32_t efc, uint32_t list_cap) {
        uint32_t o = 0;
        off_bufa = o; o += efc * 8;
        off_bufb = o; o += efc * 8;
        off_acc = o;  o += (list_cap + 1) * 8;
        off_sel = o;  o += (list_cap + 1) * 8;
        off_ids = o;  o += (list_cap + 1) * 4;
        off_dist = o; o += (list_cap + 1) * 4;
        total = o;
    }
};

// update_prune_kernel: CTA (j, slot) re-prunes neighbour j of list `slot`.  The candidate set is streamed in chunks
// (chunk -1 = {id} + the list itself, chunk c = the list of neighbour c); each chunk is evaluated against neigh's
// vector and merged BY RANK into a sorted buffer of the ef_construction closest, duplicates (same id => same key)
// dropped at the merge -- the same result as the reference's set + bounded max-heap (:1026-1051).
template <int LPV, int CPL, int METRIC>
__global__ void __launch_bounds__(kTeam) update_prune_kernel(const BuildArgs p, uint32_t *__restrict__ newlists) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t list_cap = max(p.maxM0, p.maxM);
    const UpdSmem L(p.efc, list_cap);
    uint64_t *buf[2] = {(uint64_t *)(smem + L.off_bufa), (uint64_t *)(smem + L.off_bufb)};
    uint64_t *acc = (uint64_t *)(smem + L.off_acc);
    uint64_t *sel = (uint64_t *)(smem + L.off_sel);
    uint32_t *ids = (uint32_t *)(smem + L.off_ids);
    float *dist = (float *)(smem + L.off_dist);
    __shared__ int s_n;
    GraphView g{p.vec, p.links0, p.up_base, p.links_up, p.d4, p.maxM, p.maxM0};
    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    const uint32_t slot = blockIdx.y, j = blockIdx.x;
    const uint32_t id = p.list_point[slot], level = p.list_level[slot];
    const uint32_t Mcur = level ? p.maxM : p.maxM0;
    if (j >= Mcur) return;
    const uint32_t *l1 = list_ptr(p, id, level);
    const uint32_t neigh = l1[j];
    if (neigh == kEmpty) return;
    if (p.incnt[list_id(p, neigh, level)] != slot + 1) return;  // another point of this batch re-prunes this list
    float4 q[CPL];
    load_row<LPV, CPL>(q, p.vec + (size_t)neigh * p.d4, p.d4, sub);
    const int cap = (int)p.efc;
    int size = 0, cb = 0;
    uint32_t evals = 0;
    for (int c = -1; c < (int)Mcur; c++) {
        const uint32_t *src_list = l1;
        if (c >= 0) {
            const uint32_t el = l1[c];
            if (el == kEmpty) break;  // lists are dense
            src_list = list_ptr(p, el, level);
        }
        __syncthreads();  // previous chunk's ids / acc are no longer read
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (uint32_t t = tid; t < Mcur + (c < 0 ? 1u : 0u); t += kTeam) {
            const uint32_t v = (c < 0 && t == Mcur) ? id : src_list[t];
            if (v != kEmpty && v != neigh) ids[atomicAdd(&s_n, 1)] = v;
        }
        __syncthreads();
        const int n = s_n;
        if (n == 0) continue;
        eval_list<kTeam, LPV, CPL, METRIC>(q, g.vec, p.d4, ids, n, dist, grp, sub);
        evals += (uint32_t)n;
        __syncthreads();
        const uint64_t *src = buf[cb];
        uint64_t *dst = buf[cb ^ 1];
        // new keys; one that is already in the buffer becomes the maximum key and is dropped
        if (tid == 0) s_n = 0;  // reused as the count of fresh keys (n is already in a register)
        __syncthreads();
        for (int t = tid; t < n; t += kTeam) {
            const uint64_t key = make_key(dist[t], ids[t]);
            int lo = 0, hi = size;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (src[mid] < key) lo = mid + 1; else hi = mid;
            }
            const bool fresh = !(lo < size && src[lo] == key);
            acc[t] = fresh ? key : ~0ull;
            if (fresh) atomicAdd(&s_n, 1);
        }
        __syncthreads();
        const int m = s_n;
        if (m == 0) continue;
        for (int i = tid; i < size; i += kTeam) {
            const uint64_t key = src[i];
            int pos = i;
            for (int t = 0; t < n; t++) pos += (acc[t] < key) ? 1 : 0;
            if (pos < cap) dst[pos] = key;
        }
        for (int t0 = tid; t0 < n; t0 += kTeam) {
            const uint64_t key = acc[t0];
            if (key == ~0ull) continue;
            int r = 0;
            for (int t = 0; t < n; t++) r += (acc[t] < key) ? 1 : 0;
            int lo = 0, hi = size;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (src[mid] < key) lo = mid + 1; else hi = mid;
            }
            if (r + lo < cap) dst[r + lo] = key;
        }
        size = min(cap, size + m);
        cb ^= 1;
    }
    __syncthreads();
    const int ns = heuristic_prune<LPV, CPL, METRIC>(g, buf[cb], size, (int)Mcur, sel, ids, dist, evals);
    __syncthreads();
    // the reference pops its max-heap into the list: farthest first (:1058-1063)
    uint32_t *out = newlists + ((size_t)slot * p.maxM0 + j) * p.maxM0;
    for (int t = tid; t < (int)Mcur; t += kTeam) out[t] = t < ns ? ids[ns - 1 - t] : kEmpty;
    if (tid == 0) atomicAdd(p.work + 0, (unsigned long long)evals);
}

// update_apply_kernel: the re-pruned lists replace the old ones; the claim word goes back to zero.
static __global__ void update_apply_kernel(const BuildArgs p, const uint32_t *__restrict__ newlists) {
    const uint32_t slot = blockIdx.y, j = blockIdx.x;
    const uint32_t id = p.list_point[slot], level = p.list_level[slot];
    const uint32_t Mcur = level ? p.maxM : p.maxM0;
    if (j >= Mcur) return;
    const uint32_t neigh = list_ptr(p, id, level)[j];
    if (neigh == kEmpty) return;
    const uint32_t lid = list_id(p, neigh, level);
    if (p.incnt[lid] != slot + 1) return;
    __syncthreads();
    const uint32_t *in = newlists + ((size_t)slot * p.maxM0 + j) * p.maxM0;
    uint32_t *lst = list_ptr(p, neigh, level);
    for (uint32_t t = threadIdx.x; t < Mcur; t += blockDim.x) lst[t] = in[t];
    if (threadIdx.x == 0) p.incnt[lid] = 0;
}

template <int LPV, int CPL, int METRIC>
static int run_update_phase1(const BuildArgs &a, uint32_t *newlists, cudaStream_t st) {
    const uint32_t list_cap = std::max(a.maxM0, a.maxM);
    const UpdSmem L(a.efc, list_cap);
    static bool configured[16] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (d < 16 && !configured[d]) {
        B200_CUDA_OK(cudaFuncSetAttribute(update_prune_kernel<LPV, CPL, METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024));
        configured[d] = true;
    }
    if (L.total > 200 * 1024) {
        set_error("ef_construction too large for the update kernel's shared memory");
        return B200HNSW_E_UNSUPPORTED;
    }
    update_claim_kernel<<<a.lists, 128, 0, st>>>(a);
    update_prune_kernel<LPV, CPL, METRIC><<<dim3(a.maxM0, a.lists), kTeam, L.total, st>>>(a, newlists);
    update_apply_kernel<<<dim3(a.maxM0, a.lists), 64, 0, st>>>(a, newlists);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int METRIC>
static int run_update_phase1_metric(const BuildArgs &a, uint32_t *newlists, cudaStream_t st) {
    const uint32_t d4 = a.d4;
    if (d4 <= 8) return run_update_phase1<8, 1, METRIC>(a, newlists, st);
    if (d4 <= 16) return run_update_phase1<8, 2, METRIC>(a, newlists, st);
    if (d4 <= 24) return run_update_phase1<8, 3, METRIC>(a, newlists, st);
    if (d4 <= 32) return run_update_phase1<8, 4, METRIC>(a, newlists, st);
    if (d4 <= 48) return run_update_phase1<16, 3, METRIC>(a, newlists, st);
    if (d4 <= 64) return run_update_phase1<16, 4, METRIC>(a, newlists, st);
    if (d4 <= 96) return run_update_phase1<32, 3, METRIC>(a, newlists, st);
    if (d4 <= 128) return run_update_phase1<32, 4, METRIC>(a, newlists, st);
    if (d4 <= 192) return run_update_phase1<32, 6, METRIC>(a, newlists, st);
    if (d4 <= 256) return run_update_phase1<32, 8, METRIC>(a, newlists, st);
    set_error("dimension > 1024 is not supported by the build kernels");
    return B200HNSW_E_UNSUPPORTED;
}

template <int METRIC, bool UPD, bool NB>
static int run_batch_metric(const BuildArgs &a, size_t s1, size_t s2, cudaStream_t st) {
    const uint32_t d4 = a.d4;
    if (d4 <= 8) return run_batch<8, 1, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 16) return run_batch<8, 2, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 24) return run_batch<8, 3, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 32) return run_batch<8, 4, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 48) return run_batch<16, 3, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 64) return run_batch<16, 4, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 96) return run_batch<32, 3, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 128) return run_batch<32, 4, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 192) return run_batch<32, 6, METRIC, UPD, NB>(a, s1, s2, st);
    if (d4 <= 256) return run_batch<32, 8, METRIC, UPD, NB>(a, s1, s2, st);
    set_error("dimension > 1024 is not supported by the build kernels");
    return B200HNSW_E_UNSUPPORTED;
}

}  // namespace b200

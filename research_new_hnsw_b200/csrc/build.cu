// build.cu -- batched GPU graph construction behind addPoint.
//
// Reference path replaced (file:line under /root/reference/hnswlib/hnswalg.h):
//   :1153-1267  addPoint(data, label, level)      -> HnswIndex::add_batch (staging, eager public fields) + flush()
//   :1214-1239  greedy descent above the new level -> build_search_kernel prologue (greedy_level)
//   :225-305    searchBaseLayer (ef_construction)  -> build_search_kernel (beam_level on every level <= the new one)
//   :443-483    getNeighborsByHeuristic2           -> heuristic_prune (device function)
//   :506-630    mutuallyConnectNewElement          -> build_link_kernel (forward lists + reverse-edge staging) and
//                                                     build_reverse_kernel (append, or re-prune to Mcurmax when full)
//
// Batching.  The reference inserts one point at a time; here points are linked in insertion order in batches that
// see the graph as of the start of their batch.  A batch is never larger than (linked points) / build_ratio, so at
// most a 1/build_ratio (default 1/32) fraction of a point's true neighbours is invisible to it, and a point that raises the
// maximum level always ends its batch (it becomes the entry point of everything after it, :1262-1265).  Levels,
// element count, entry point and max level are assigned at add time with the reference's generator and are
// bit-identical to the reference for the same insertion order (:207-211,1187-1198).
//
// Reverse links are lock-free: build_link_kernel appends (dist, new id) to a per-list incoming buffer with one
// atomicAdd per edge and records each touched list once; build_reverse_kernel then gives every touched list to one
// CTA which appends while there is room (:586-588) and otherwise re-runs the heuristic over existing + incoming
// (:590-612), once per batch instead of once per edge.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hnsw_index.cuh"

namespace b200 {

constexpr int kTeam = 128;      // threads per CTA of the build kernels
constexpr uint32_t kCapIn = 32;  // incoming reverse edges kept per list and batch

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
    unsigned long long *work;  // [4] D, H0, Hup, resets (atomicAdd)
    uint32_t cap, first, batch, lists;
    uint32_t entry;
    int32_t maxlevel;
    uint32_t d4, maxM, maxM0, M, efc, hash_bits;
    // update mode only (kernels instantiated with UPD = true re-link EXISTING points: repairConnectionsForUpdate,
    // hnswalg.h:1075-1139); kept at the end so the insert-mode kernels see the layout they were tuned with
    const uint32_t *batch_ids;  // [batch] ids of the points of this batch (insert mode: first + b)
};

template <int LPV, int CPL>
__device__ __forceinline__ void load_row(float4 (&v)[CPL], const float4 *row, uint32_t d4, int sub) {
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        const uint32_t idx = sub + c * LPV;
        v[c] = idx < d4 ? __ldg(row + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Construction search for one new point: CTA b handles point first + b on all of its levels.
template <int LPV, int CPL, int METRIC, bool UPD>
__global__ void __launch_bounds__(kTeam) build_search_kernel(const BuildArgs p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t list_cap = p.maxM0 > p.maxM ? p.maxM0 : p.maxM;
    const SearchSmem L(p.efc, list_cap, p.d4, p.hash_bits);
    __shared__ int s_ints[kTeamInts];
    TeamCtx c;
    c.bind(smem, L, s_ints, p.hash_bits);
    GraphView g{p.vec, p.links0, p.up_base, p.links_up, p.d4, p.maxM, p.maxM0};

    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    const uint32_t pid = UPD ? p.batch_ids[blockIdx.x] : p.first + blockIdx.x;
    const uint32_t HS = 1u << p.hash_bits;
    const int plevel = p.plevel[pid];

    float4 q[CPL];
    load_row<LPV, CPL>(q, p.vec + (size_t)pid * p.d4, p.d4, sub);
    if (tid == 0) { *c.s_cnt = 0; *c.s_acc = 0; *c.s_next = 0; c.ids[0] = p.entry; }
    __syncthreads();
    WorkCounters w;
    uint32_t cur = p.entry;
    eval_list<kTeam, LPV, CPL, METRIC>(q, g.vec, p.d4, c.ids, 1, c.dist, grp, sub);
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
        beam_level<kTeam, LPV, CPL, METRIC>(c, q, g, level, p.efc, cur, curdist, cb, size, w);
        const uint64_t *res = cb ? c.buf_b : c.buf_a;
        uint64_t *out = p.cand + (size_t)(slot0 + level) * p.efc;
        for (int j = tid; j < size; j += kTeam) out[j] = res[j] & kKeyMask;
        if (tid == 0) p.cand_cnt[slot0 + level] = (uint32_t)size;
        // next level starts from the closest candidate (= selectedNeighbors.back(), hnswalg.h:524,629: the closest
        // candidate always survives the heuristic)
        cur = (uint32_t)res[0] & kIdMask;
        curdist = ord2f((uint32_t)(res[0] >> 32));
    }
    if (tid == 0) {
        atomicAdd(p.work + 0, (unsigned long long)w.D); atomicAdd(p.work + 1, (unsigned long long)w.H0);
        atomicAdd(p.work + 2, (unsigned long long)w.Hup); atomicAdd(p.work + 3, (unsigned long long)w.resets);
    }
}

// getNeighborsByHeuristic2 (hnswalg.h:443-483) for one candidate list sorted closest-first: accept c iff every already
// accepted r has dist(r, c) >= dist(base, c); stop at Mlimit.  Selected keys end up in sel[0..ns), their ids in ids[].
// The candidate's vector is register-resident and the next candidate's is prefetched while the current one is
// compared against the accepted set (whose rows are re-read through L1/L2).
template <int LPV, int CPL, int METRIC>
__device__ __forceinline__ int heuristic_prune(const GraphView &g, const uint64_t *cand, int n, int Mlimit,
                                               uint64_t *sel, uint32_t *ids, float *dist, uint32_t &evals) {
    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    if (n < Mlimit) {  // hnswalg.h:446-448
        for (int j = tid; j < n; j += kTeam) { sel[j] = cand[j]; ids[j] = (uint32_t)cand[j] & kIdMask; }
        __syncthreads();
        return n;
    }
    int ns = 0;
    float4 v[CPL], vn[CPL];
    load_row<LPV, CPL>(v, g.vec + (size_t)((uint32_t)cand[0] & kIdMask) * g.d4, g.d4, sub);
    for (int ci = 0; ci < n && ns < Mlimit; ci++) {
        const uint64_t key = cand[ci];
        const float dq = ord2f((uint32_t)(key >> 32));
        if (ci + 1 < n) load_row<LPV, CPL>(vn, g.vec + (size_t)((uint32_t)cand[ci + 1] & kIdMask) * g.d4, g.d4, sub);
        eval_list<kTeam, LPV, CPL, METRIC, true>(v, g.vec, g.d4, ids, ns, dist, grp, sub);
        evals += ns;
        __syncthreads();
        bool bad = false;
        for (int j = tid; j < ns; j += kTeam) bad |= dist[j] < dq;
        bad = __syncthreads_or(bad);
        if (!bad) {
            if (tid == 0) { sel[ns] = key; ids[ns] = (uint32_t)key & kIdMask; }
            ns++;
        }
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < CPL; cc++) v[cc] = vn[cc];
    }
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
    }
    if (tid == 0) atomicAdd(p.work + 0, (unsigned long long)evals);
}

// One CTA per touched list: append the incoming new points while there is room, otherwise re-run the heuristic over
// existing + incoming neighbours (distances to this node) and rewrite the list.
template <int LPV, int CPL, int METRIC, bool UPD>
__global__ void __launch_bounds__(kTeam) build_reverse_kernel(const BuildArgs p, uint32_t n_aff) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t capc = max(p.maxM0, p.maxM) + kCapIn;
    const LinkSmem L(capc);
    uint64_t *sel = (uint64_t *)(smem + L.off_sel);
    uint64_t *raw = (uint64_t *)(smem + L.off_raw);
    uint64_t *srt = (uint64_t *)(smem + L.off_srt);
    uint32_t *ids = (uint32_t *)(smem + L.off_ids);
    float *dist = (float *)(smem + L.off_dist);
    GraphView g{p.vec, p.links0, p.up_base, p.links_up, p.d4, p.maxM, p.maxM0};
    const int tid = threadIdx.x;
    const int sub = tid % LPV, grp = tid / LPV;
    if (blockIdx.x >= n_aff) return;
    const uint32_t node = p.aff_node[blockIdx.x], level = p.aff_level[blockIdx.x];
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

template <int LPV, int CPL, int METRIC, bool UPD>
static int run_batch(const BuildArgs &a, size_t smem_search, size_t smem_link, uint32_t *h_aff, cudaStream_t st) {
    static bool configured[16] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (d < 16 && !configured[d]) {
        cudaFuncAttributes fa;
        int optin = 0;
        B200_CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d));
        B200_CUDA_OK(cudaFuncGetAttributes(&fa, build_search_kernel<LPV, CPL, METRIC, UPD>));
        B200_CUDA_OK(cudaFuncSetAttribute(build_search_kernel<LPV, CPL, METRIC, UPD>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        configured[d] = true;
    }
    build_search_kernel<LPV, CPL, METRIC, UPD><<<a.batch, kTeam, smem_search, st>>>(a);
    build_link_kernel<LPV, CPL, METRIC, UPD><<<a.lists, kTeam, smem_link, st>>>(a);
    B200_CUDA_OK(cudaMemcpyAsync(h_aff, a.aff_count, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
    const uint32_t n_aff = *h_aff;
    if (n_aff) build_reverse_kernel<LPV, CPL, METRIC, UPD><<<n_aff, kTeam, smem_link, st>>>(a, n_aff);
    B200_CUDA_OK(cudaMemsetAsync(a.aff_count, 0, 4, st));
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int METRIC, bool UPD = false>
static int run_batch_metric(const BuildArgs &a, size_t s1, size_t s2, uint32_t *h_aff, cudaStream_t st) {
    const uint32_t d4 = a.d4;
    if (d4 <= 8) return run_batch<8, 1, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 16) return run_batch<8, 2, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 24) return run_batch<8, 3, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 32) return run_batch<8, 4, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 48) return run_batch<16, 3, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 64) return run_batch<16, 4, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 96) return run_batch<32, 3, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 128) return run_batch<32, 4, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 192) return run_batch<32, 6, METRIC, UPD>(a, s1, s2, h_aff, st);
    if (d4 <= 256) return run_batch<32, 8, METRIC, UPD>(a, s1, s2, h_aff, st);
    set_error("dimension > 1024 is not supported by the build kernels");
    return B200HNSW_E_UNSUPPORTED;
}

// ---- host side ---------------------------------------------------------------------------------------------

static size_t env_size(const char *name, size_t dflt) {
    if (const char *e = getenv(name)) {
        const long long v = atoll(e);
        if (v > 0) return (size_t)v;
    }
    return dflt;
}

// addPoint staging (hnswalg.h:1153-1211,1255-1265): everything that does not need a distance.
int HnswIndex::add_batch(const float *X, const uint64_t *labels, size_t n, bool replace_deleted) {
    std::lock_guard<std::mutex> g(mu);
    HostImage &m = host;
    std::vector<uint32_t> updates;
    for (size_t i = 0; i < n; i++) {
        const uint64_t lab = labels ? labels[i] : (uint64_t)m.cur;
        auto known = m.label_lookup.find(lab);
        if (known != m.label_lookup.end()) {
            // existing label: update instead of insert (hnswalg.h:1157-1174)
            const uint32_t c = known->second;
            if (m.deleted(c)) {
                if (prm.allow_replace_deleted) {
                    set_error("Can't use addPoint to update deleted elements if replacement of deleted elements is enabled.");
                    return B200HNSW_E_STATE;
                }
                *((unsigned char *)m.rec(c) + 2) &= (unsigned char)~1;  // unmarkDeletedInternal
                m.num_deleted--;
                flags_dirty = true;
            }
            memcpy(m.rec(c) + m.off_data, X + i * m.dim, m.dim * 4);
            if (c < linked) updates.push_back(c);  // a staged point is simply linked with its new vector later
            continue;
        }
        if (replace_deleted && m.num_deleted > 0) {
            // vacant place: a deleted element takes the new label and vector (hnswalg.h:965-992)
            size_t c = replace_scan < m.cur ? replace_scan : 0;
            while (!m.deleted(c)) c = c + 1 < m.cur ? c + 1 : 0;  // num_deleted > 0: terminates
            replace_scan = c + 1;
            uint64_t old_label;
            memcpy(&old_label, m.rec(c) + m.off_label, 8);
            m.label_lookup.erase(old_label);
            m.label_lookup[lab] = (uint32_t)c;
            memcpy(m.rec(c) + m.off_label, &lab, 8);
            *((unsigned char *)m.rec(c) + 2) &= (unsigned char)~1;
            m.num_deleted--;
            flags_dirty = true;
            memcpy(m.rec(c) + m.off_data, X + i * m.dim, m.dim * 4);
            if (c < linked) updates.push_back((uint32_t)c);
            continue;
        }
        if (m.cur >= m.max_elements) {
            set_error("The number of elements exceeds the specified limit");
            return B200HNSW_E_CAPACITY;
        }
        const size_t c = m.cur++;
        m.label_lookup[lab] = (uint32_t)c;
        const int level = m.random_level();
        m.levels[c] = level;
        memset(m.rec(c), 0, m.size_data);
        memcpy(m.rec(c) + m.off_label, &lab, 8);
        memcpy(m.rec(c) + m.off_data, X + i * m.dim, m.dim * 4);
        m.upper[c].assign((size_t)level * m.size_links, 0);
        if (c == 0) {
            m.enterpoint = 0;
            m.maxlevel = level;
        } else if (level > m.maxlevel) {
            m.enterpoint = (uint32_t)c;
            m.maxlevel = level;
        }
    }
    return updates.empty() ? 0 : relink_points(std::move(updates));
}

// updatePoint (hnswalg.h:995-1139) for points that are already part of the device graph, batched.  The new vector is
// uploaded, then repairConnectionsForUpdate runs on the GPU with the construction kernels in update mode: search from the
// entry point, the point itself dropped from its candidates, heuristic selection, forward list REPLACED, reverse edges
// added unless already present (mutuallyConnectNewElement with isUpdate, :485-639).  The first phase of the reference --
// re-pruning every old neighbour over the 1-hop/2-hop neighbourhood (:1009-1069) -- is NOT performed: old neighbours keep
// their edge to the moved point.
int HnswIndex::relink_points(std::vector<uint32_t> ids) {
    HostImage &m = host;
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    B200_CUDA_OK(cudaSetDevice(dev.device));
    {   // new vectors -> padded device rows
        std::vector<float> row(dev.d4 * 4, 0.f);
        for (uint32_t id : ids) {
            memcpy(row.data(), m.rec(id) + m.off_data, m.dim * 4);
            B200_CUDA_OK(cudaMemcpy((float *)dev.vec + (size_t)id * dev.d4 * 4, row.data(), dev.d4 * 16, cudaMemcpyHostToDevice));
            B200_CUDA_OK(cudaMemcpy(dev.labels + id, m.rec(id) + m.off_label, 8, cudaMemcpyHostToDevice));  // replace_deleted relabels
            const int rc16 = sync_bf16(id, 1);
            if (rc16) return rc16;
        }
    }
    if (linked <= 1) return 0;  // a single element has nothing to connect to (hnswalg.h:1001-1003)
    const size_t build_ratio = env_size("B200HNSW_BUILD_RATIO", 32);
    const size_t max_batch = env_size("B200HNSW_BUILD_BATCH", 16384);
    if (!bld.plevel) B200_CUDA_OK(cudaMalloc(&bld.plevel, std::max<size_t>(dev.cap, 1) * 4));
    B200_CUDA_OK(cudaMemcpy(bld.plevel, m.levels.data(), linked * 4, cudaMemcpyHostToDevice));
    BuildArgs a{};
    size_t max_lists = 0, smem_search = 0, smem_link = 0;
    int rc = prepare_build(&a, max_batch, &max_lists, &smem_search, &smem_link);
    if (rc) return rc;
    if (!bld.batch_ids) B200_CUDA_OK(cudaMalloc(&bld.batch_ids, max_batch * 4));
    std::vector<uint32_t> off, lp, ll;
    uint32_t h_aff = 0;
    uint64_t launches = 0;
    for (size_t b0 = 0; b0 < ids.size() && rc == 0;) {
        size_t B = std::max<size_t>(1, std::min(max_batch, linked / build_ratio));
        B = std::min(B, ids.size() - b0);
        off.clear(); lp.clear(); ll.clear();
        size_t used = 0;
        for (; used < B; used++) {
            const uint32_t id = ids[b0 + used];
            const int top = std::min(m.levels[id], dev_maxlevel);
            if (lp.size() + (size_t)top + 1 > max_lists) break;
            off.push_back((uint32_t)lp.size());
            for (int l = 0; l <= top; l++) { lp.push_back(id); ll.push_back((uint32_t)l); }
        }
        B = used;
        B200_CUDA_OK(cudaMemcpyAsync(bld.batch_ids, ids.data() + b0, B * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_off, off.data(), B * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_point, lp.data(), lp.size() * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_level, ll.data(), ll.size() * 4, cudaMemcpyHostToDevice, stream));
        a.first = 0; a.batch = (uint32_t)B; a.lists = (uint32_t)lp.size();
        a.entry = dev_entry; a.maxlevel = dev_maxlevel;
        a.batch_ids = bld.batch_ids;
        rc = prm.metric == B200HNSW_L2 ? run_batch_metric<0, true>(a, smem_search, smem_link, &h_aff, stream)
                                       : run_batch_metric<1, true>(a, smem_search, smem_link, &h_aff, stream);
        launches += 3;
        b0 += B;
    }
    if (rc == 0) B200_CUDA_OK(cudaStreamSynchronize(stream));
    stats.kernel_launches += launches;
    mirror_dirty = true;
    return rc;
}

// Device graph -> reference-layout host mirror (needed by saveIndex and get_linklist*).
int HnswIndex::sync_host_mirror() {
    if (!mirror_dirty) return 0;
    B200_CUDA_OK(cudaSetDevice(dev.device));
    HostImage &m = host;
    const size_t n = m.cur;
    const size_t chunk = std::max<size_t>(1, (size_t)(64u << 20) / (m.maxM0 * 4));
    std::vector<uint32_t> tmp(chunk * m.maxM0);
    for (size_t s = 0; s < n; s += chunk) {
        const size_t c = std::min(chunk, n - s);
        B200_CUDA_OK(cudaMemcpy(tmp.data(), dev.links0 + s * m.maxM0, c * m.maxM0 * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < c; i++) {
            const uint32_t *src = tmp.data() + i * m.maxM0;
            uint32_t *dst = m.list(s + i, 0);
            unsigned cnt = 0;
            while (cnt < m.maxM0 && src[cnt] != kEmpty) cnt++;
            HostImage::set_count(dst, (unsigned short)cnt);
            memcpy(dst + 1, src, cnt * 4);
            memset(dst + 1 + cnt, 0, (m.maxM0 - cnt) * 4);
        }
    }
    if (dev.up_lists) {
        std::vector<uint32_t> up(dev.up_lists * m.maxM);
        B200_CUDA_OK(cudaMemcpy(up.data(), dev.links_up, up.size() * 4, cudaMemcpyDeviceToHost));
        size_t li = 0;
        for (size_t i = 0; i < n; i++)
            for (int l = 1; l <= m.levels[i]; l++, li++) {
                const uint32_t *src = up.data() + li * m.maxM;
                uint32_t *dst = m.list(i, l);
                unsigned cnt = 0;
                while (cnt < m.maxM && src[cnt] != kEmpty) cnt++;
                dst[0] = 0;
                HostImage::set_count(dst, (unsigned short)cnt);
                memcpy(dst + 1, src, cnt * 4);
                memset(dst + 1 + cnt, 0, (m.maxM - cnt) * 4);
            }
    }
    mirror_dirty = false;
    return 0;
}

// Scratch of the build kernels (sized for batches of max_batch points) and the kernel arguments that do not depend on
// the batch; shared by flush() (new points) and relink_points() (updatePoint).
int HnswIndex::prepare_build(void *args, size_t max_batch, size_t *max_lists_out, size_t *smem_search, size_t *smem_link) {
    HostImage &m = host;
    const size_t nl = dev.cap + dev.up_lists_cap;
    if (!bld.plevel) B200_CUDA_OK(cudaMalloc(&bld.plevel, std::max<size_t>(dev.cap, 1) * 4));
    if (!bld.incnt) {
        B200_CUDA_OK(cudaMalloc(&bld.incnt, nl * 4));
        B200_CUDA_OK(cudaMemset(bld.incnt, 0, nl * 4));
        B200_CUDA_OK(cudaMalloc(&bld.incoming, nl * kCapIn * 8));
    }
    const size_t max_lists = max_batch + max_batch / 2 + 64;  // level-0 list per point + the rare upper lists
    *max_lists_out = max_lists;
    if (!bld.cand || bld.cand_efc != m.efc) {
        cudaFree(bld.cand); cudaFree(bld.cand_cnt); cudaFree(bld.list_off); cudaFree(bld.list_point);
        cudaFree(bld.list_level); cudaFree(bld.aff_node); cudaFree(bld.aff_level); cudaFree(bld.aff_count);
        cudaFree(bld.work);
        B200_CUDA_OK(cudaMalloc(&bld.cand, max_lists * m.efc * 8));
        B200_CUDA_OK(cudaMalloc(&bld.cand_cnt, max_lists * 4));
        B200_CUDA_OK(cudaMalloc(&bld.list_off, max_batch * 4));
        B200_CUDA_OK(cudaMalloc(&bld.list_point, max_lists * 4));
        B200_CUDA_OK(cudaMalloc(&bld.list_level, max_lists * 4));
        B200_CUDA_OK(cudaMalloc(&bld.aff_node, max_lists * m.M * 4));
        B200_CUDA_OK(cudaMalloc(&bld.aff_level, max_lists * m.M * 4));
        B200_CUDA_OK(cudaMalloc(&bld.aff_count, 4));
        B200_CUDA_OK(cudaMemset(bld.aff_count, 0, 4));
        B200_CUDA_OK(cudaMalloc(&bld.work, 32));
        bld.cand_efc = m.efc;
    }
    B200_CUDA_OK(cudaMemset(bld.work, 0, 32));

    const size_t list_cap = std::max(m.maxM, m.maxM0);
    BuildArgs &a = *(BuildArgs *)args;
    a = BuildArgs{};
    a.vec = dev.vec; a.links0 = dev.links0; a.up_base = dev.up_base; a.links_up = dev.links_up;
    a.plevel = bld.plevel; a.cand = bld.cand; a.cand_cnt = bld.cand_cnt; a.list_off = bld.list_off;
    a.list_point = bld.list_point; a.list_level = bld.list_level; a.incnt = bld.incnt; a.incoming = bld.incoming;
    a.aff_node = bld.aff_node; a.aff_level = bld.aff_level; a.aff_count = bld.aff_count; a.work = bld.work;
    a.cap = (uint32_t)dev.cap; a.d4 = (uint32_t)dev.d4; a.maxM = (uint32_t)m.maxM; a.maxM0 = (uint32_t)m.maxM0;
    a.M = (uint32_t)m.M; a.efc = (uint32_t)m.efc;
    // construction searches evaluate ~40 * efc nodes; a table of ~32 * efc slots is rebuilt about once in four searches
    // and lets twice as many CTAs share an SM as the no-rebuild size (measured: -30 % build time, same graph)
    {
        size_t want = std::min<size_t>(8192, 32 * m.efc + 1024);
        want = std::max(want, 2 * (m.efc + list_cap));
        a.hash_bits = 10;
        while ((1ull << a.hash_bits) < want) a.hash_bits++;
    }
    const SearchSmem SL(a.efc, (uint32_t)list_cap, a.d4, a.hash_bits);
    const LinkSmem LL((uint32_t)list_cap + kCapIn);
    if (SL.total > 226 * 1024) {
        set_error("ef_construction too large for the build kernel's shared memory");
        return B200HNSW_E_UNSUPPORTED;
    }
    *smem_search = SL.total;
    *smem_link = LL.total;
    return 0;
}


// Link every staged point (ids [linked, host.cur)) into the device graph.
int HnswIndex::flush() {
    std::lock_guard<std::mutex> g(mu);
    HostImage &m = host;
    if (linked >= m.cur) return 0;
    B200_CUDA_OK(cudaSetDevice(dev.device));
    const size_t n_new = m.cur - linked;
    const size_t rec = m.size_data;
    const size_t build_ratio = env_size("B200HNSW_BUILD_RATIO", 32);
    const size_t max_batch = env_size("B200HNSW_BUILD_BATCH", 16384);

    // ---- upload vectors + labels of the staged points (records carry empty lists) ----
    {
        const size_t chunk = std::max<size_t>(1, std::min<size_t>(n_new, (size_t)(256u << 20) / rec));
        uint32_t *raw = nullptr;
        B200_CUDA_OK(cudaMalloc(&raw, chunk * rec));
        for (size_t first = linked; first < m.cur; first += chunk) {
            const size_t cnt = std::min(chunk, m.cur - first);
            cudaError_t e = cudaMemcpy(raw, m.level0 + first * rec, cnt * rec, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) {
                deinterleave_kernel<<<(unsigned)((cnt * 32 + 255) / 256), 256>>>(
                    raw, rec / 4, (uint32_t)first, (uint32_t)cnt, (uint32_t)m.maxM0, (uint32_t)m.dim, (uint32_t)dev.d4,
                    (uint32_t)m.cur, (float *)dev.vec, dev.links0, dev.labels, dev.err_flag);
                e = cudaDeviceSynchronize();
            }
            if (e != cudaSuccess) {
                cudaFree(raw);
                set_error(std::string("CUDA error during upload: ") + cudaGetErrorString(e));
                return B200HNSW_E_CUDA;
            }
        }
        cudaFree(raw);
    }
    {
        int rc16 = sync_bf16(linked, n_new);
        if (rc16) return rc16;
    }
    // ---- upper-level list slots of the staged points (appended after the existing ones) ----
    {
        std::vector<uint32_t> base(n_new, kEmpty);
        size_t lists = dev.up_lists;
        for (size_t i = 0; i < n_new; i++)
            if (m.levels[linked + i] > 0) {
                base[i] = (uint32_t)lists;
                lists += (size_t)m.levels[linked + i];
            }
        if (lists > dev.up_lists_cap) {
            const size_t ncap = std::max(lists, dev.up_lists_cap * 2);
            uint32_t *nu = nullptr;
            B200_CUDA_OK(cudaMalloc(&nu, ncap * m.maxM * 4));
            B200_CUDA_OK(cudaMemset(nu, 0xFF, ncap * m.maxM * 4));
            if (dev.up_lists)
                B200_CUDA_OK(cudaMemcpy(nu, dev.links_up, dev.up_lists * m.maxM * 4, cudaMemcpyDeviceToDevice));
            cudaFree(dev.links_up);
            dev.links_up = nu;
            dev.up_lists_cap = ncap;
            cudaFree(bld.incnt); cudaFree(bld.incoming);  // sized by cap + up_lists_cap
            bld.incnt = nullptr; bld.incoming = nullptr;
        } else if (lists > dev.up_lists) {
            B200_CUDA_OK(cudaMemset(dev.links_up + dev.up_lists * m.maxM, 0xFF, (lists - dev.up_lists) * m.maxM * 4));
        }
        dev.up_lists = lists;
        B200_CUDA_OK(cudaMemcpy(dev.up_base + linked, base.data(), n_new * 4, cudaMemcpyHostToDevice));
    }
    // ---- build scratch + kernel arguments ----
    if (!bld.plevel) B200_CUDA_OK(cudaMalloc(&bld.plevel, std::max<size_t>(dev.cap, 1) * 4));
    B200_CUDA_OK(cudaMemcpy(bld.plevel + linked, m.levels.data() + linked, n_new * 4, cudaMemcpyHostToDevice));
    BuildArgs a{};
    size_t max_lists = 0, smem_search = 0, smem_link = 0;
    {
        const int rcp = prepare_build(&a, max_batch, &max_lists, &smem_search, &smem_link);
        if (rcp) return rcp;
    }

    cudaEvent_t e0, e1;
    B200_CUDA_OK(cudaEventCreate(&e0));
    B200_CUDA_OK(cudaEventCreate(&e1));
    B200_CUDA_OK(cudaEventRecord(e0, stream));
    std::vector<uint32_t> off, lp, ll;
    uint32_t h_aff = 0;
    uint64_t launches = 0;
    int rc = 0;
    while (linked < m.cur && rc == 0) {
        if (linked == 0) {  // first element: nothing to link (hnswalg.h:1255-1259)
            dev_entry = 0;
            dev_maxlevel = m.levels[0];
            linked = 1;
            continue;
        }
        size_t B = std::max<size_t>(1, std::min(max_batch, linked / build_ratio));
        B = std::min(B, m.cur - linked);
        for (size_t j = 0; j < B; j++)
            if (m.levels[linked + j] > dev_maxlevel) { B = j + 1; break; }  // new top level ends the batch
        off.resize(B); lp.clear(); ll.clear();
        for (size_t j = 0; j < B; j++) {
            off[j] = (uint32_t)lp.size();
            const int top = std::min(m.levels[linked + j], dev_maxlevel);
            for (int l = 0; l <= top; l++) { lp.push_back((uint32_t)(linked + j)); ll.push_back((uint32_t)l); }
        }
        if (lp.size() > max_lists) {  // cannot happen with P(level >= 1) = 1/M, but never overrun the pool
            size_t j = B;
            while (j > 1 && off[j - 1] + 8 > max_lists) j--;
            B = j;
            off.resize(B);
            size_t keep = 0;
            for (size_t i = 0; i < lp.size(); i++) if (lp[i] < linked + B) keep = i + 1;
            lp.resize(keep); ll.resize(keep);
        }
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_off, off.data(), B * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_point, lp.data(), lp.size() * 4, cudaMemcpyHostToDevice, stream));
        B200_CUDA_OK(cudaMemcpyAsync(bld.list_level, ll.data(), ll.size() * 4, cudaMemcpyHostToDevice, stream));
        a.first = (uint32_t)linked; a.batch = (uint32_t)B; a.lists = (uint32_t)lp.size();
        a.entry = dev_entry; a.maxlevel = dev_maxlevel;
        rc = prm.metric == B200HNSW_L2 ? run_batch_metric<0>(a, smem_search, smem_link, &h_aff, stream)
                                       : run_batch_metric<1>(a, smem_search, smem_link, &h_aff, stream);
        launches += 3;
        const size_t last = linked + B - 1;
        if (m.levels[last] > dev_maxlevel) {
            dev_entry = (uint32_t)last;
            dev_maxlevel = m.levels[last];
        }
        linked += B;
    }
    if (rc == 0) {
        B200_CUDA_OK(cudaEventRecord(e1, stream));
        B200_CUDA_OK(cudaStreamSynchronize(stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long w[4] = {0, 0, 0, 0};
        B200_CUDA_OK(cudaMemcpy(w, bld.work, 32, cudaMemcpyDeviceToHost));
        stats.queries = n_new;
        stats.dist_evals = w[0]; stats.hops_base = w[1]; stats.hops_upper = w[2]; stats.visited_resets = w[3];
        stats.kernel_launches += launches;
        stats.last_kernel_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev.n = linked;
    mirror_dirty = true;
    flags_dirty = true;
    return rc;
}

void BuildScratch::release() {
    cudaFree(plevel); cudaFree(cand); cudaFree(cand_cnt); cudaFree(list_off); cudaFree(list_point);
    cudaFree(list_level); cudaFree(incnt); cudaFree(incoming); cudaFree(aff_node); cudaFree(aff_level);
    cudaFree(aff_count); cudaFree(work); cudaFree(batch_ids);
    *this = BuildScratch();
}

}  // namespace b200

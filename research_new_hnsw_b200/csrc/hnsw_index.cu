// hnsw_index.cu -- HBM image management and search dispatch for HnswIndex.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "search_launch.cuh"

namespace b200 {

SearchCtx::~SearchCtx() {
    cudaFree(dQ); cudaFree(dLabels); cudaFree(dDists); cudaFree(dCounts); cudaFree(dWork);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev2) cudaEventDestroy(ev2);
    if (stream) cudaStreamDestroy(stream);
    if (stream2) cudaStreamDestroy(stream2);
}

int SearchCtx::ensure(size_t nq, size_t k, size_t dim) {
    if (!stream) {
        B200_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        B200_CUDA_OK(cudaEventCreate(&ev0));
        B200_CUDA_OK(cudaEventCreate(&ev1));
    }
    if (nq <= scratch_q && k <= scratch_k) return 0;
    const size_t q = std::max(nq, scratch_q), kk = std::max(k, scratch_k);
    cudaFree(dQ); cudaFree(dLabels); cudaFree(dDists); cudaFree(dCounts); cudaFree(dWork);
    dQ = nullptr; dLabels = nullptr; dDists = nullptr; dCounts = dWork = nullptr;
    scratch_q = scratch_k = 0;
    B200_CUDA_OK(cudaMalloc(&dQ, q * dim * 4));
    B200_CUDA_OK(cudaMalloc(&dLabels, q * kk * 8));
    B200_CUDA_OK(cudaMalloc(&dDists, q * kk * 4));
    B200_CUDA_OK(cudaMalloc(&dCounts, q * 4));
    B200_CUDA_OK(cudaMalloc(&dWork, q * 16));
    scratch_q = q;
    scratch_k = kk;
    return 0;
}

// A free context, or a new one while the pool is below its size (B200HNSW_SEARCH_CTXS, default 4), else wait.
SearchCtx *HnswIndex::acquire_ctx(bool wait) {
    static const size_t max_ctx = [] {
        const char *e = getenv("B200HNSW_SEARCH_CTXS");
        const int v = e ? atoi(e) : 4;
        return (size_t)std::min(16, std::max(1, v));
    }();
    std::unique_lock<std::mutex> lk(ctx_mu);
    for (;;) {
        for (auto &c : ctxs)
            if (!c->busy) { c->busy = true; return c.get(); }
        if (ctxs.size() < max_ctx) {
            ctxs.emplace_back(new SearchCtx());
            ctxs.back()->busy = true;
            return ctxs.back().get();
        }
        if (!wait) return nullptr;
        ctx_cv.wait(lk);
    }
}

void HnswIndex::release_ctx(SearchCtx *c) {
    {
        std::lock_guard<std::mutex> lk(ctx_mu);
        c->busy = false;
    }
    ctx_cv.notify_one();
}

HnswIndex::~HnswIndex() {
    for (auto &c : ctxs)  // a launch of search_submit that nobody waited for must not outlive the arrays it reads
        if (c->stream) cudaStreamSynchronize(c->stream);
    if (dev.cap || dev.err_flag) {
        cudaSetDevice(dev.device);
        dev.release();
        bld.release();
    }
    ctxs.clear();
    if (stream) cudaStreamDestroy(stream);
}

int HnswIndex::init_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
        return B200HNSW_E_CUDA;
    }
    int d = prm.device;
    if (d < 0) B200_CUDA_OK(cudaGetDevice(&d));
    if (d >= count) {
        set_error("device ordinal out of range");
        return B200HNSW_E_ARG;
    }
    dev.device = d;
    B200_CUDA_OK(cudaSetDevice(d));
    B200_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return 0;
}

int HnswIndex::alloc_device(size_t cap) {
    B200_CUDA_OK(cudaSetDevice(dev.device));
    dev.release();
    bld.release();
    dev.cap = cap;
    dev.dim = host.dim;
    dev.d4 = (host.dim + 3) / 4;
    dev.maxM = host.maxM;
    dev.maxM0 = host.maxM0;
    const size_t c = cap ? cap : 1;
    B200_CUDA_OK(cudaMalloc(&dev.vec, c * dev.d4 * 16));
    B200_CUDA_OK(cudaMalloc(&dev.links0, c * dev.maxM0 * 4));
    B200_CUDA_OK(cudaMalloc(&dev.up_base, c * 4));
    B200_CUDA_OK(cudaMalloc(&dev.labels, c * 8));
    B200_CUDA_OK(cudaMalloc(&dev.err_flag, 4));
    B200_CUDA_OK(cudaMemset(dev.err_flag, 0, 4));
    B200_CUDA_OK(cudaMemset(dev.up_base, 0xFF, c * 4));
    B200_CUDA_OK(cudaMemset(dev.links0, 0xFF, c * dev.maxM0 * 4));
    if (prm.storage == B200HNSW_BF16) {
        dev.d16 = (host.dim + 7) / 8;
        B200_CUDA_OK(cudaMalloc(&dev.vec16, c * dev.d16 * 16));
    }
    return 0;
}

// refresh the bf16 copy of rows [first, first+count) from the fp32 rows already in HBM
// bf16 copy of rows [first, first + count) (storage variant).  With a stream: enqueued there, no synchronisation.
int HnswIndex::sync_bf16(size_t first, size_t count, cudaStream_t st) {
    if (prm.storage != B200HNSW_BF16 || !count) return 0;
    B200_CUDA_OK(cudaSetDevice(dev.device));
    const size_t tot = count * dev.d16;
    rows_to_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(dev.vec, (uint32_t)dev.d4, (uint32_t)dev.d16,
                                                                     (uint32_t)first, (uint32_t)count, dev.vec16);
    B200_CUDA_OK(cudaGetLastError());
    if (!st) B200_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int HnswIndex::upload_upper() {
    const size_t n = host.cur;
    std::vector<uint32_t> base(n ? n : 1, kEmpty);
    size_t lists = 0;
    for (size_t i = 0; i < n; i++)
        if (host.levels[i] > 0) {
            base[i] = (uint32_t)lists;
            lists += (size_t)host.levels[i];
        }
    std::vector<uint32_t> up((lists ? lists : 1) * host.maxM, kEmpty);
    for (size_t i = 0; i < n; i++) {
        for (int l = 1; l <= host.levels[i]; l++) {
            const uint32_t *ll = host.list(i, l);
            const unsigned cnt = HostImage::count_of(ll);
            if (cnt > host.maxM) {
                set_error("Index seems to be corrupted or unsupported");
                return B200HNSW_E_CORRUPT;
            }
            uint32_t *dst = up.data() + ((size_t)base[i] + (l - 1)) * host.maxM;
            for (unsigned j = 0; j < cnt; j++) {
                const uint32_t v = ll[1 + j];
                // a neighbour on level l must itself own a level-l list (hnswalg.h:547-548)
                if (v >= n || host.levels[v] < l) {
                    set_error("cand error");
                    return B200HNSW_E_CAND;
                }
                dst[j] = v;
            }
        }
    }
    B200_CUDA_OK(cudaSetDevice(dev.device));
    if (lists > dev.up_lists_cap || !dev.links_up) {
        cudaFree(dev.links_up);
        dev.links_up = nullptr;
        dev.up_lists_cap = std::max<size_t>(lists, 1);
        B200_CUDA_OK(cudaMalloc(&dev.links_up, dev.up_lists_cap * host.maxM * 4));
    }
    dev.up_lists = lists;
    B200_CUDA_OK(cudaMemcpy(dev.links_up, up.data(), std::max<size_t>(lists, 1) * host.maxM * 4, cudaMemcpyHostToDevice));
    if (n) B200_CUDA_OK(cudaMemcpy(dev.up_base, base.data(), n * 4, cudaMemcpyHostToDevice));
    return 0;
}

int HnswIndex::upload_all() {
    B200_CUDA_OK(cudaSetDevice(dev.device));
    const size_t n = host.cur;
    dev.n = n;
    if (n) {
        // stream the raw level-0 block through a bounded staging buffer and de-interleave it on the device
        const size_t rec = host.size_data;
        const size_t chunk = std::max<size_t>(1, std::min<size_t>(n, (size_t)(256u << 20) / rec));
        uint32_t *raw = nullptr;
        B200_CUDA_OK(cudaMalloc(&raw, chunk * rec));
        for (size_t first = 0; first < n; first += chunk) {
            const size_t cnt = std::min(chunk, n - first);
            cudaError_t e = cudaMemcpy(raw, host.level0 + first * rec, cnt * rec, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) {
                const unsigned threads = 256;
                const unsigned blocks = (unsigned)((cnt * 32 + threads - 1) / threads);
                deinterleave_kernel<<<blocks, threads>>>(raw, rec / 4, (uint32_t)first, (uint32_t)cnt,
                                                         (uint32_t)host.maxM0, (uint32_t)host.dim, (uint32_t)dev.d4,
                                                         (uint32_t)n, (float *)dev.vec, dev.links0, dev.labels,
                                                         dev.err_flag);
                e = cudaDeviceSynchronize();
            }
            if (e != cudaSuccess) {
                cudaFree(raw);
                set_error(std::string("CUDA error during upload: ") + cudaGetErrorString(e));
                return B200HNSW_E_CUDA;
            }
        }
        cudaFree(raw);
        uint32_t flag = 0;
        B200_CUDA_OK(cudaMemcpy(&flag, dev.err_flag, 4, cudaMemcpyDeviceToHost));
        if (flag) {
            set_error(flag & 2 ? "Index seems to be corrupted or unsupported" : "cand error");
            return flag & 2 ? B200HNSW_E_CORRUPT : B200HNSW_E_CAND;
        }
    }
    linked = n;
    dev_entry = host.enterpoint;
    dev_maxlevel = host.maxlevel;
    mirror_dirty = false;
    int rc16 = sync_bf16(0, n);
    if (rc16) return rc16;
    return upload_upper();
}

// delete marks (byte 2 of the level-0 record header, hnswalg.h:934-937) -> device byte array
// `allowed` (nullable, one byte per internal id) is a BaseFilterFunctor evaluated by the caller: a node that is not
// allowed is kept out of top_candidates exactly like a deleted one (hnswalg.h:406-407), so both share the flag byte.
int HnswIndex::upload_flags(const uint8_t *allowed, const uint32_t *extra, size_t n_extra) {
    if (!flags_dirty && dev.flags && !allowed && !n_extra) return 0;
    B200_CUDA_OK(cudaSetDevice(dev.device));
    if (!dev.flags) {
        B200_CUDA_OK(cudaMalloc(&dev.flags, std::max<size_t>(dev.cap, 1)));
        flags_on_device_valid = false;
    }
    std::vector<uint8_t> f(host.cur);
    for (size_t i = 0; i < host.cur; i++) f[i] = (host.deleted(i) || (allowed && !allowed[i])) ? 1 : 0;
    for (size_t i = 0; i < n_extra; i++)
        if (extra[i] < host.cur) f[extra[i]] = 1;
    {
        size_t rej = 0;
        for (uint8_t v : f) rej += v;
        flags_rejected = rej;  // sizes the candidate buffer of the non-bare search (launch_search)
    }
    // calls that pass the same filter over and over (the shim evaluates a functor into the same verdicts each time) do
    // not pay the upload again: the device array is only rewritten when its content changes
    uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)host.cur;
    {
        size_t i = 0;
        for (; i + 8 <= f.size(); i += 8) {
            uint64_t w;
            memcpy(&w, f.data() + i, 8);
            h = (h ^ w) * 0x100000001b3ull;
            h ^= h >> 29;
        }
        for (; i < f.size(); i++) h = (h ^ f[i]) * 0x100000001b3ull;
    }
    if (!(flags_on_device_valid && h == flags_on_device_hash)) {
        if (host.cur) B200_CUDA_OK(cudaMemcpy(dev.flags, f.data(), host.cur, cudaMemcpyHostToDevice));
        flags_on_device_hash = h;
        flags_on_device_valid = true;
    }
    flags_dirty = allowed != nullptr || n_extra != 0;  // a per-call state must not outlive its call
    return 0;
}

// Visited-table size.  128-thread teams (8 resident queries per SM): large enough that a typical query at this ef
// never rebuilds it (D ~ 30*ef + 500 on the 1M x 128, M=32 graph).  Smaller teams trade table
// size for resident queries: the table is rebuilt from the buffer at 5/8 load (re-evaluations only).
uint32_t pick_hash_bits(size_t ef, size_t list_cap, int team) {
    // The table is rebuilt from the buffer when it is more than 5/8 full at the START of a hop, and a hop inserts up to
    // list_cap ids: 5/8 * size + list_cap must stay below the size (or hash_insert's probe loop would never find a free
    // slot), and the rebuilt table (<= ef entries) must itself be below the mark.
    const size_t need = std::max(2 * (ef + list_cap), 8 * list_cap / 3 + 64);
    if (const char *e = getenv("B200HNSW_HASH_BITS")) {
        const int b = atoi(e);
        if (b >= 8 && b <= 15) {
            uint32_t bits = (uint32_t)b;
            while ((1ull << bits) < need) bits++;
            return bits;
        }
    }
    size_t want = 64 * ef + 1024;
    if (want > 16384) want = 16384;
    // Throughput teams (64 / 32 threads): resident queries per SM matter more than avoiding rebuilds.  Measured on the
    // 1M x 128 graph, 10 k queries (ms per batch, default-size table vs this policy): ef=64 3.31 -> 2.79, ef=128
    // 8.82 -> 5.18, ef=256 20.3 -> 10.0, with 7-20 % more evaluations from the rebuilds.
    // (re-measured with the round-2 kernel, 1024 / 2048 / 4096 slots, ms per 10 k queries: ef=64 2.02 / 1.93 / 2.12,
    // ef=128 4.05 / 3.98 / 4.53, ef=256 8.25 / 8.15 / 9.86 -- 2048 slots at every ef)
    if (team <= 64) want = 2048;
    if (want < need) want = need;
    uint32_t bits = 10;
    while ((1ull << bits) < want) bits++;
    return bits;
}

// Team size: a batch that cannot fill the GPU with 64-thread teams is latency-bound -> 128 threads per query.
int pick_team(size_t nq) {
    if (const char *e = getenv("B200HNSW_TEAM")) {
        const int t = atoi(e);
        if (t == 32 || t == 64 || t == 128) return t;
    }
    return nq >= 148 * 16 ? 64 : 128;
}

int HnswIndex::launch_search(const float *dQ_, size_t nq, size_t k, size_t ef_, uint64_t *dl, float *dd,
                             uint32_t *dc, uint32_t *dw, cudaStream_t st, const uint8_t *allowed) {
    if (nq == 0) return 0;
    if (k == 0 || !dQ_ || !dl || !dd) {
        set_error("search: null pointer or k == 0");
        return B200HNSW_E_ARG;
    }
    B200_CUDA_OK(cudaSetDevice(dev.device));
    const bool nonbare = host.num_deleted != 0 || allowed;  // hnswalg.h:1306: bare_bone = !num_deleted_ && !isIdAllowed
    if (nonbare) {
        if (linked >= (1u << 30)) {
            set_error("search with deleted elements supports at most 2^30 elements");
            return B200HNSW_E_UNSUPPORTED;
        }
        // `allowed` or dirty marks: the caller holds `rw` exclusively (search_host / search_device arrange that)
        if (allowed || flags_dirty || !dev.flags) {
            int rc = upload_flags(allowed);
            if (rc) return rc;
        }
    }
    if (linked == 0) {  // hnswalg.h:1273: empty index -> empty result
        B200_CUDA_OK(cudaMemsetAsync(dl, 0xFF, nq * k * 8, st));
        fill_pad_rows(dd, dc, dw, nq, k, st);  // dist = +inf, counts = 0
        return 0;
    }
    size_t efx = ef_ ? ef_ : ef.load();
    efx = std::max(efx, k);  // hnswalg.h:1309
    if (efx > 4096) {
        set_error("ef > 4096 is not supported");
        return B200HNSW_E_UNSUPPORTED;
    }
    const size_t list_cap = std::max(host.maxM, host.maxM0);
    SearchArgs a{};
    a.vec = dev.vec; a.links0 = dev.links0; a.up_base = dev.up_base; a.links_up = dev.links_up;
    a.labels = dev.labels; a.Q = dQ_; a.out_labels = dl; a.out_dists = dd; a.out_counts = dc; a.out_work = dw;
    a.n = (uint32_t)linked; a.entry = dev_entry; a.maxlevel = dev_maxlevel;
    a.dim = (uint32_t)host.dim; a.d4 = (uint32_t)dev.d4; a.maxM = (uint32_t)host.maxM; a.maxM0 = (uint32_t)host.maxM0;
    a.nq = (uint32_t)nq; a.k = (uint32_t)k; a.ef = (uint32_t)efx;
    int team = nonbare ? 128 : pick_team(nq);
    if (prm.storage == B200HNSW_BF16 && team == 32) team = 64;
    const bool bf16 = prm.storage == B200HNSW_BF16 && !nonbare;  // deleted elements: fp32 rows (non-bare kernel)
    a.vec16 = bf16 ? dev.vec16 : nullptr;
    a.d16 = (uint32_t)dev.d16;
    a.flags = nonbare ? dev.flags : nullptr;
    a.bufcap = (uint32_t)efx;
    if (nonbare) {
        // The reference keeps EVERY visited element inside the bound in candidate_set, deleted or not (hnswalg.h:395-408);
        // only the non-deleted ones count towards ef.  With a fraction f of the elements rejected (deleted or filtered
        // out) about ef * f / (1 - f) rejected entries sit inside the bound of a full result set: the buffer holds three
        // times that expectation (at least ef, at most 15 * ef) on top of the ef results, as far as shared memory allows.
        const double f = std::min(0.999, (double)flags_rejected / (double)std::max<size_t>(linked, 1));
        size_t extra = (size_t)(3.0 * (double)efx * f / (1.0 - f)) + 32;
        extra = std::min(std::max(extra, efx), 15 * efx);
        size_t cap = efx + extra;
        while (cap > 2 * efx &&
               SearchSmem((uint32_t)cap, (uint32_t)list_cap, a.d4, pick_hash_bits(cap, list_cap, team)).total > 96 * 1024)
            cap = std::max(2 * efx, cap * 3 / 4);
        a.bufcap = (uint32_t)cap;
    }
    a.hash_bits = pick_hash_bits(a.bufcap, list_cap, team);
    {
        static const int pf_env = getenv("B200HNSW_PF") ? atoi(getenv("B200HNSW_PF")) : -1;
        a.pf = pf_env >= 0 ? (uint32_t)pf_env : (kPfRows | kPfGreedy | kPfRound1);
    }
    const SearchSmem L(a.bufcap, (uint32_t)list_cap, a.d4, a.hash_bits);
    if (L.total > 226 * 1024) {
        set_error("search configuration needs more than 226 KB of shared memory");
        return B200HNSW_E_UNSUPPORTED;
    }
    {
        std::lock_guard<std::mutex> sg(stats_mu);
        stats.kernel_launches += 1;
    }
    const bool l2 = prm.metric == B200HNSW_L2;
    if (nonbare) return l2 ? search_launch_l2_var(a, L.total, 1, st) : search_launch_ip_var(a, L.total, 1, st);
    if (bf16) return l2 ? search_launch_l2_var(a, L.total, team == 64 ? 2 : 3, st) : search_launch_ip_var(a, L.total, team == 64 ? 2 : 3, st);
    return l2 ? search_launch_l2_bare(a, L.total, team, st) : search_launch_ip_bare(a, L.total, team, st);
}

__global__ void fill_pad_kernel(float *dd, uint32_t *dc, uint32_t *dw, size_t nq, size_t k) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * k) dd[i] = __int_as_float(0x7f800000);
    if (dc && i < nq) dc[i] = 0;
    if (dw && i < nq * 4) dw[i] = 0;
}

void fill_pad_rows(float *dd, uint32_t *dc, uint32_t *dw, size_t nq, size_t k, cudaStream_t st) {
    const size_t tot = std::max(nq * k, nq * 4);
    fill_pad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(dd, dc, dw, nq, k);
}

// Micro-batching of concurrent single-query calls: hnsw_service's pool threads each call
// searchKnn with one query (main.cpp:59-67).  Callers queue their request; the first one becomes the leader, gives
// the others a few microseconds to arrive, runs ONE batched launch for every queued request with the same (k, ef)
// and hands the rows back.  A lone caller pays only the short gather window.
int HnswIndex::search_coalesced(const float *Q, size_t k, size_t ef_, uint64_t *labels, float *dists, uint32_t *counts,
                                uint32_t *work) {
    Pending me{Q, k, ef_, labels, dists, counts, work, 0, false, std::string()};
    std::unique_lock<std::mutex> lk(co_mu);
    co_queue.push_back(&me);
    for (;;) {
        if (me.done) break;
        if (co_leader) {  // somebody else is running a batch: wait for it (ours may be in it, or in the next one)
            co_cv.wait(lk);
            continue;
        }
        co_leader = true;
        // gather window: concurrent callers are typically microseconds apart; sleep on the condition variable until the
        // deadline instead of spinning (followers do not notify, a finished batch of another leader cannot exist here)
        co_cv.wait_for(lk, std::chrono::microseconds(co_window_us));
        std::vector<Pending *> batch;
        for (auto it = co_queue.begin(); it != co_queue.end();) {
            if ((*it)->k == me.k && (*it)->ef == me.ef && batch.size() < 1024) {
                batch.push_back(*it);
                it = co_queue.erase(it);
            } else {
                ++it;
            }
        }
        lk.unlock();
        const size_t nb = batch.size(), d = host.dim;
        int rc = 0;
        std::string err;
        std::vector<uint64_t> l;
        std::vector<float> dd;
        std::vector<uint32_t> c, w;
        try {  // whatever happens here, every request of the batch is completed and the leader role is released
            std::vector<float> q(nb * d);
            l.resize(nb * k);
            dd.resize(nb * k);
            c.resize(nb);
            w.resize(nb * 4);
            for (size_t i = 0; i < nb; i++) memcpy(q.data() + i * d, batch[i]->q, d * 4);
            rc = search_host(q.data(), nb, k, ef_, l.data(), dd.data(), c.data(), w.data());
            if (rc) err = b200hnsw_last_error();
        } catch (const std::bad_alloc &) {
            rc = B200HNSW_E_NOMEM;
            err = "Not enough memory";
        } catch (const std::exception &e) {
            rc = B200HNSW_E_STATE;
            err = std::string("internal error: ") + e.what();
        }
        lk.lock();
        for (size_t i = 0; i < nb; i++) {
            Pending *p = batch[i];
            if (!rc) {
                memcpy(p->labels, l.data() + i * k, k * 8);
                memcpy(p->dists, dd.data() + i * k, k * 4);
                if (p->counts) *p->counts = c[i];
                if (p->work) memcpy(p->work, w.data() + i * 4, 16);
            }
            p->rc = rc;
            p->err = err;
            p->done = true;
        }
        co_batches++;
        co_queries += nb;
        co_leader = false;
        co_cv.notify_all();
    }
    lk.unlock();
    if (me.rc) set_error(me.err);
    return me.rc;
}

// Shared lock for a search once the delete marks on the device are current; exclusive when they are not (they are
// uploaded by launch_search) or when the call brings its own filter mask (the mask lives in the same device array).
struct SearchLock {
    std::shared_lock<std::shared_mutex> sh;
    std::unique_lock<std::shared_mutex> ex;
    SearchLock(HnswIndex &ix, bool filtered) {
        if (!filtered) {
            sh = std::shared_lock<std::shared_mutex>(ix.rw);
            if (!(ix.host.num_deleted != 0 && (ix.flags_dirty || !ix.dev.flags))) return;
            sh.unlock();
        }
        ex = std::unique_lock<std::shared_mutex>(ix.rw);
        ix.drain_async();  // the marks on the device are about to change under launches still in flight
    }
};

int HnswIndex::search_device(const float *dQ_, size_t nq, size_t k, size_t ef_, uint64_t *dl, float *dd, uint32_t *dc,
                             uint32_t *dw, cudaStream_t st) {
    int rc = flush();
    if (rc) return rc;
    SearchLock lock(*this, false);
    return launch_search(dQ_, nq, k, ef_, dl, dd, dc, dw, st);
}

int HnswIndex::search_host(const float *Q, size_t nq, size_t k, size_t ef_, uint64_t *labels, float *dists,
                           uint32_t *counts, uint32_t *work, const uint8_t *allowed) {
    if (nq == 0) return 0;
    if (!Q || !labels || !dists || k == 0) {
        set_error("search: null pointer or k == 0");
        return B200HNSW_E_ARG;
    }
    int rc = flush();
    if (rc) return rc;
    SearchCtx *c = acquire_ctx();
    {
        SearchLock lock(*this, allowed != nullptr);
        rc = search_host_locked(*c, Q, nq, k, ef_, labels, dists, counts, work, allowed);
    }
    release_ctx(c);
    return rc;
}

void HnswIndex::drain_async() {
    std::vector<cudaStream_t> pending;
    {
        std::lock_guard<std::mutex> lk(ctx_mu);
        for (auto &c : ctxs)
            if (c->async_pending && c->stream) pending.push_back(c->stream);
    }
    for (cudaStream_t st : pending) cudaStreamSynchronize(st);
}

int HnswIndex::search_submit(const float *Q, size_t nq, size_t k, size_t ef_, uint64_t *labels, float *dists,
                             uint32_t *counts, uint64_t *ticket_out) {
    if (!ticket_out) { set_error("ticket_out is null"); return B200HNSW_E_ARG; }
    *ticket_out = 0;
    if (nq == 0) return 0;
    if (!Q || !labels || !dists || k == 0) {
        set_error("search: null pointer or k == 0");
        return B200HNSW_E_ARG;
    }
    int rc = flush();
    if (rc) return rc;
    B200_CUDA_OK(cudaSetDevice(dev.device));
    void *dq = nullptr, *dl = nullptr, *dd = nullptr, *dc = nullptr;
    bool ok = cudaHostGetDevicePointer(&dq, (void *)Q, 0) == cudaSuccess &&
              cudaHostGetDevicePointer(&dl, (void *)labels, 0) == cudaSuccess &&
              cudaHostGetDevicePointer(&dd, (void *)dists, 0) == cudaSuccess;
    if (ok && counts) ok = cudaHostGetDevicePointer(&dc, (void *)counts, 0) == cudaSuccess;
    if (!ok) {  // pageable buffers: nothing to overlap with, the call completes here and the ticket is 0
        cudaGetLastError();
        return search_host(Q, nq, k, ef_, labels, dists, counts, nullptr);
    }
    // never block here: the caller may be the very thread that has to wait for the batches in flight
    SearchCtx *c = acquire_ctx(false);
    if (!c) {
        set_error("search_submit: too many batches in flight (wait for one of them first)");
        return B200HNSW_E_STATE;
    }
    rc = c->ensure(1, 1, host.dim);  // stream + events only
    if (!rc) {
        SearchLock lock(*this, false);
        rc = launch_search((const float *)dq, nq, k, ef_, (uint64_t *)dl, (float *)dd, (uint32_t *)dc, nullptr, c->stream);
        if (!rc) {
            std::lock_guard<std::mutex> lk(ctx_mu);
            c->async_pending = true;
        }
    }
    if (rc) {
        release_ctx(c);
        return rc;
    }
    std::lock_guard<std::mutex> lk(ctx_mu);
    for (size_t i = 0; i < ctxs.size(); i++)
        if (ctxs[i].get() == c) *ticket_out = i + 1;
    return 0;
}

int HnswIndex::search_wait(uint64_t ticket) {
    if (ticket == 0) return 0;
    SearchCtx *c = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx_mu);
        if (ticket > ctxs.size() || !ctxs[ticket - 1]->async_pending) {
            set_error("search_wait: no launch is pending for this ticket");
            return B200HNSW_E_ARG;
        }
        c = ctxs[ticket - 1].get();
    }
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    {
        std::lock_guard<std::mutex> lk(ctx_mu);
        c->async_pending = false;
    }
    release_ctx(c);
    if (e != cudaSuccess) {
        set_error(std::string("CUDA error: ") + cudaGetErrorString(e));
        return B200HNSW_E_CUDA;
    }
    return 0;
}

int HnswIndex::search_host_locked(SearchCtx &c, const float *Q, size_t nq, size_t k, size_t ef_, uint64_t *labels,
                                  float *dists, uint32_t *counts, uint32_t *work, const uint8_t *allowed) {
    B200_CUDA_OK(cudaSetDevice(dev.device));
    int rc = c.ensure(nq, k, host.dim);
    if (rc) return rc;
    // Large batches are cut into chunks that alternate between two streams, so the H2D copy of one chunk and the
    // D2H copy of the previous one overlap the kernel of the chunk in between (kernels of both streams share the SMs).
    // Only worth it when every host buffer is page-locked: a copy to or from pageable memory blocks the calling thread.
    auto pinned = [](const void *p) {
        if (!p) return true;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool all_pinned = pinned(Q) && pinned(labels) && pinned(dists) && pinned(counts) && pinned(work);
    // Page-locked buffers are device-accessible under unified addressing: the kernel reads its query (d floats per CTA)
    // and stores its k result rows straight over PCIe -- the copies are a few KB per query spread over the whole launch,
    // so the call costs one kernel launch and nothing is staged.  (B200HNSW_ZEROCOPY=0: staged copies as below.)
    static const bool zc_on = !(getenv("B200HNSW_ZEROCOPY") && atoi(getenv("B200HNSW_ZEROCOPY")) == 0);
    if (zc_on && all_pinned && !allowed && nq >= 64) {
        void *dq = nullptr, *dl = nullptr, *dd = nullptr, *dc = nullptr, *dw = nullptr;
        bool ok = cudaHostGetDevicePointer(&dq, (void *)Q, 0) == cudaSuccess &&
                  cudaHostGetDevicePointer(&dl, (void *)labels, 0) == cudaSuccess &&
                  cudaHostGetDevicePointer(&dd, (void *)dists, 0) == cudaSuccess;
        if (ok && counts) ok = cudaHostGetDevicePointer(&dc, (void *)counts, 0) == cudaSuccess;
        if (ok && work) ok = cudaHostGetDevicePointer(&dw, (void *)work, 0) == cudaSuccess;
        if (ok) {
            B200_CUDA_OK(cudaEventRecord(c.ev0, c.stream));
            rc = launch_search((const float *)dq, nq, k, ef_, (uint64_t *)dl, (float *)dd, (uint32_t *)dc, (uint32_t *)dw,
                               c.stream);
            if (rc) return rc;
            B200_CUDA_OK(cudaEventRecord(c.ev1, c.stream));
            B200_CUDA_OK(cudaStreamSynchronize(c.stream));
            float ms = 0;
            cudaEventElapsedTime(&ms, c.ev0, c.ev1);
            std::lock_guard<std::mutex> sg(stats_mu);
            stats.last_kernel_ms = ms;
            stats.queries = nq;
            stats.dist_evals = stats.hops_base = stats.hops_upper = stats.visited_resets = 0;
            if (work)
                for (size_t i = 0; i < nq; i++) {
                    stats.dist_evals += work[i * 4 + 0];
                    stats.hops_base += work[i * 4 + 1];
                    stats.hops_upper += work[i * 4 + 2];
                    stats.visited_resets += work[i * 4 + 3];
                }
            return 0;
        }
        cudaGetLastError();  // not mappable after all: staged copies
    }
    const bool async_ok = !allowed && nq >= 4096 && all_pinned;
    size_t chunks = async_ok ? 3 : 1;  // measured: 3 chunks beat 2 and 4+ (smaller kernels lose more to their tails)
    if (async_ok)
        if (const char *e = getenv("B200HNSW_CHUNKS")) chunks = (size_t)std::min(16, std::max(1, atoi(e)));
    const size_t per = (nq + chunks - 1) / chunks;
    if (chunks > 1 && !c.stream2) {
        B200_CUDA_OK(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
        B200_CUDA_OK(cudaEventCreateWithFlags(&c.ev2, cudaEventDisableTiming));
    }
    const size_t d = host.dim;
    B200_CUDA_OK(cudaEventRecord(c.ev0, c.stream));
    if (chunks > 1) {  // the second stream starts after ev0 so the event pair brackets everything
        B200_CUDA_OK(cudaStreamWaitEvent(c.stream2, c.ev0, 0));
    }
    for (size_t ch = 0; ch < chunks; ch++) {
        const size_t off = ch * per;
        if (off >= nq) break;
        const size_t n = std::min(per, nq - off);
        cudaStream_t st = (ch & 1) ? c.stream2 : c.stream;
        B200_CUDA_OK(cudaMemcpyAsync(c.dQ + off * d, Q + off * d, n * d * 4, cudaMemcpyHostToDevice, st));
        rc = launch_search(c.dQ + off * d, n, k, ef_, c.dLabels + off * k, c.dDists + off * k, c.dCounts + off,
                           c.dWork + off * 4, st, allowed);
        if (rc) return rc;
        B200_CUDA_OK(cudaMemcpyAsync(labels + off * k, c.dLabels + off * k, n * k * 8, cudaMemcpyDeviceToHost, st));
        B200_CUDA_OK(cudaMemcpyAsync(dists + off * k, c.dDists + off * k, n * k * 4, cudaMemcpyDeviceToHost, st));
        if (counts) B200_CUDA_OK(cudaMemcpyAsync(counts + off, c.dCounts + off, n * 4, cudaMemcpyDeviceToHost, st));
        if (work) B200_CUDA_OK(cudaMemcpyAsync(work + off * 4, c.dWork + off * 4, n * 16, cudaMemcpyDeviceToHost, st));
    }
    if (chunks > 1) {
        B200_CUDA_OK(cudaEventRecord(c.ev2, c.stream2));
        B200_CUDA_OK(cudaStreamWaitEvent(c.stream, c.ev2, 0));
    }
    B200_CUDA_OK(cudaEventRecord(c.ev1, c.stream));
    B200_CUDA_OK(cudaStreamSynchronize(c.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    std::lock_guard<std::mutex> sg(stats_mu);
    stats.last_kernel_ms = ms;  // kernel (+ overlapped copies when chunked)
    stats.queries = nq;
    stats.dist_evals = stats.hops_base = stats.hops_upper = stats.visited_resets = 0;
    if (work)
        for (size_t i = 0; i < nq; i++) {
            stats.dist_evals += work[i * 4 + 0];
            stats.hops_base += work[i * 4 + 1];
            stats.hops_upper += work[i * 4 + 2];
            stats.visited_resets += work[i * 4 + 3];
        }
    return 0;
}

}  // namespace b200

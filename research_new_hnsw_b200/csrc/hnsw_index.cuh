// hnsw_index.cuh -- HierarchicalNSW<float> replacement: host mirror + HBM image + kernel dispatch.
#pragma once
#include <atomic>
#include <condition_variable>
#include <list>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "device_index.cuh"
#include "search_kernel.cuh"

namespace b200 {

// Device scratch of the batched graph build (build.cu); allocated on first flush.
struct BuildScratch {
    int32_t *plevel = nullptr;
    uint64_t *cand = nullptr;
    uint32_t *cand_cnt = nullptr, *list_off = nullptr, *list_point = nullptr, *list_level = nullptr, *batch_ids = nullptr;
    uint32_t *incnt = nullptr;
    uint64_t *incoming = nullptr;
    uint32_t *aff_node = nullptr, *aff_level = nullptr, *aff_count = nullptr;
    unsigned long long *work = nullptr;
    size_t cand_efc = 0;
    // updatePoint: re-pruned neighbour lists of the first phase, staging of the new rows / labels
    uint32_t *newlists = nullptr;
    size_t newlists_cap = 0;
    float *stage_rows = nullptr;
    uint64_t *stage_labels = nullptr;
    size_t stage_cap = 0;
    // flush(): batch plan on the device and the upload pipeline, kept between calls (a cudaFree at the end of every
    // flush synchronises the device and was seen to take up to 0.8 s)
    uint32_t *plan_off = nullptr, *plan_lp = nullptr, *plan_ll = nullptr;
    size_t plan_off_cap = 0, plan_lists_cap = 0;
    uint32_t *up_raw[2] = {nullptr, nullptr};
    size_t up_raw_bytes = 0;
    cudaEvent_t up_ev[2] = {nullptr, nullptr};
    cudaStream_t up_stream = nullptr;
    void release();
};

// Device scratch + streams of ONE host-pointer search call.  Concurrent callers each take a context from a small pool, so
// batch searches from several host threads overlap on the GPU instead of queueing behind one mutex.
struct SearchCtx {
    float *dQ = nullptr;
    uint64_t *dLabels = nullptr;
    float *dDists = nullptr;
    uint32_t *dCounts = nullptr, *dWork = nullptr;
    size_t scratch_q = 0, scratch_k = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    bool busy = false;
    bool async_pending = false;  // a search_submit launch has not been waited for yet
    ~SearchCtx();
    int ensure(size_t nq, size_t k, size_t dim);
};

struct HnswIndex {
    b200hnsw_params prm{};
    HostImage host;
    DeviceIndex dev;
    std::atomic<size_t> ef{10};  // hnswalg.h:115 (setEf is an unsynchronised write in the reference; atomic here)
    // Locking (the reference: label-op / link-list locks for writers, lock-free const searchKnn, hnswalg.h:40-43,59):
    // every call that changes the host image, the device graph or the delete marks holds `rw` exclusively; searches and
    // read-only accessors hold it shared.  C-ABI entry points take the lock; the *_locked / launch_* members assume it.
    std::shared_mutex rw;
    std::atomic<bool> has_staged{false};   // add_batch staged points that flush() has not linked yet
    std::mutex stats_mu;
    b200hnsw_stats stats{};
    // staged insertions (add_batch before flush): ids [linked, host.cur) are not yet in the graph
    size_t linked = 0;
    uint32_t dev_entry = (uint32_t)-1;  // entry point / max level of the linked part of the graph
    int dev_maxlevel = -1;
    bool mirror_dirty = false;          // device link lists are newer than the host mirror
    std::atomic<bool> flags_dirty{true};  // delete marks changed since the last upload
    size_t flags_rejected = 0;          // elements marked in the last uploaded flag array (deleted or filtered out)
    BuildScratch bld;
    cudaStream_t stream = nullptr;      // build / update stream
    // pool of search contexts (host-pointer search path)
    std::mutex ctx_mu;
    std::condition_variable ctx_cv;
    std::vector<std::unique_ptr<SearchCtx>> ctxs;
    SearchCtx *acquire_ctx(bool wait = true);  // wait = false: nullptr when every context is busy
    void release_ctx(SearchCtx *c);
    // asynchronous form of search_host for page-locked buffers: submit enqueues the zero-copy launch on a context of
    // its own and returns a ticket, wait blocks until that launch has finished.  Writers drain pending launches first.
    int search_submit(const float *Q, size_t nq, size_t k, size_t ef_, uint64_t *labels, float *dists, uint32_t *counts,
                      uint64_t *ticket_out);
    int search_wait(uint64_t ticket);
    void drain_async();  // caller holds `rw` exclusively

    // micro-batching of concurrent single-query calls
    struct Pending {
        const float *q; size_t k, ef; uint64_t *labels; float *dists; uint32_t *counts; uint32_t *work;
        int rc; bool done; std::string err;
    };
    std::mutex co_mu;
    std::condition_variable co_cv;
    std::list<Pending *> co_queue;
    bool co_leader = false;
    unsigned co_window_us = 20;
    uint64_t co_batches = 0, co_queries = 0;
    int search_coalesced(const float *Q, size_t k, size_t ef_, uint64_t *labels, float *dists, uint32_t *counts,
                         uint32_t *work);

    ~HnswIndex();
    int init_device();
    int alloc_device(size_t cap);
    int upload_all();                       // host mirror -> HBM (after load)
    int upload_upper();                     // rebuild up_base / links_up from the host mirror
    // extra: internal ids uploaded as deleted although the host image says live (relink_points)
    int upload_flags(const uint8_t *allowed = nullptr, const uint32_t *extra = nullptr, size_t n_extra = 0);
    bool revived_on_device = false;
    uint64_t flags_on_device_hash = 0;   // content hash of dev.flags as last uploaded
    bool flags_on_device_valid = false;  // false after anything else wrote dev.flags (scatter_rows_kernel) or reallocated it
    int sync_bf16(size_t first, size_t count, cudaStream_t st = nullptr);
    // kernel launch only; the caller holds `rw` (shared is enough unless `allowed` is given or the marks are dirty)
    int launch_search(const float *dQ_, size_t nq, size_t k, size_t ef_, uint64_t *dl, float *dd, uint32_t *dc,
                      uint32_t *dw, cudaStream_t st, const uint8_t *allowed = nullptr);
    // entry points of the C ABI: link staged points, bring the delete marks up to date, take the lock, search
    int search_device(const float *dQ_, size_t nq, size_t k, size_t ef_, uint64_t *dl, float *dd, uint32_t *dc,
                      uint32_t *dw, cudaStream_t st);
    int search_host(const float *Q, size_t nq, size_t k, size_t ef_, uint64_t *labels, float *dists,
                    uint32_t *counts, uint32_t *work, const uint8_t *allowed = nullptr);
    int search_host_locked(SearchCtx &c, const float *Q, size_t nq, size_t k, size_t ef_, uint64_t *labels, float *dists,
                           uint32_t *counts, uint32_t *work, const uint8_t *allowed);
    // build.cu: addPoint staging and the batched GPU graph build
    int add_batch(const float *X, const uint64_t *labels, size_t n, bool replace_deleted = false);
    void stage_records(const float *X, const uint64_t *labels, size_t cur0, const std::vector<size_t> &rows);
    size_t replace_scan = 0;  // where the search for a deleted slot resumes (replace_deleted)
    int flush();              // takes `rw` exclusively when there is something to link
    int flush_locked();       // caller holds `rw` exclusively
    // build.cu: scratch + batch-independent kernel arguments (args is a BuildArgs*)
    int prepare_build(void *args, size_t max_batch, size_t *max_lists, size_t *smem_search, size_t *smem_link);
    // build.cu: updatePoint for ids that are already linked (vectors in the host image are the new ones); rw held
    int relink_points(std::vector<uint32_t> ids, const std::vector<uint32_t> &revived = {});
    int sync_host_mirror();   // caller holds `rw` exclusively
};

uint32_t pick_hash_bits(size_t ef, size_t list_cap, int team = 128);
int pick_team(size_t nq);
void fill_pad_rows(float *dd, uint32_t *dc, uint32_t *dw, size_t nq, size_t k, cudaStream_t st);

}  // namespace b200

// the opaque handle of include/b200hnsw.h
struct b200hnsw_index { b200::HnswIndex ix; };

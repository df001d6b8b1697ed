// capi.cu -- extern "C" surface of libb200hnsw.so (include/b200hnsw.h).  No exception crosses this boundary.
#include <cstring>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <string>

#include "bruteforce.cuh"
#include "hnsw_index.cuh"
#include "merge_launch.cuh"

struct b200bf_index { b200::BruteIndex ix; };

namespace b200 {
static thread_local std::string t_last_error;
void set_error(const std::string &msg) { t_last_error = msg; }
}  // namespace b200

using b200::set_error;

#define B200_GUARD_BEGIN try {
#define B200_GUARD_END                                            \
    }                                                             \
    catch (const std::bad_alloc &) {                              \
        set_error("Not enough memory");                           \
        return B200HNSW_E_NOMEM;                                  \
    }                                                             \
    catch (const std::exception &e) {                             \
        set_error(std::string("internal error: ") + e.what());    \
        return B200HNSW_E_STATE;                                  \
    }

static int check_params(const b200hnsw_params *p) {
    if (!p) { set_error("params is null"); return B200HNSW_E_ARG; }
    if (p->dim == 0) { set_error("dim must be > 0"); return B200HNSW_E_ARG; }
    if (p->metric != B200HNSW_L2 && p->metric != B200HNSW_IP) { set_error("unknown metric"); return B200HNSW_E_ARG; }
    if (p->storage != B200HNSW_F32 && p->storage != B200HNSW_BF16) { set_error("unknown storage type"); return B200HNSW_E_ARG; }
    return 0;
}

extern "C" {

const char *b200hnsw_last_error(void) { return b200::t_last_error.c_str(); }
int b200hnsw_abi_version(void) { return B200HNSW_ABI_VERSION; }

int b200hnsw_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error(std::string("CUDA error: ") + cudaGetErrorString(e));
        return B200HNSW_E_CUDA;
    }
    return n;
}

int b200hnsw_create(const b200hnsw_params *params, b200hnsw_index **out) {
    B200_GUARD_BEGIN
    if (!out) { set_error("out is null"); return B200HNSW_E_ARG; }
    *out = nullptr;
    int rc = check_params(params);
    if (rc) return rc;
    if (params->M < 2) { set_error("M must be >= 2"); return B200HNSW_E_ARG; }
    b200hnsw_index *h = new b200hnsw_index();
    h->ix.prm = *params;
    if (!h->ix.host.init(params->dim, params->max_elements, params->M, params->ef_construction, params->random_seed)) {
        delete h;
        set_error("Not enough memory");
        return B200HNSW_E_NOMEM;
    }
    rc = h->ix.init_device();
    if (!rc) rc = h->ix.alloc_device(params->max_elements);
    if (!rc) rc = h->ix.upload_upper();
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
    B200_GUARD_END
}

int b200hnsw_load(const char *path, const b200hnsw_params *params, b200hnsw_index **out) {
    B200_GUARD_BEGIN
    if (!out || !path) { set_error("null argument"); return B200HNSW_E_ARG; }
    *out = nullptr;
    int rc = check_params(params);
    if (rc) return rc;
    b200hnsw_index *h = new b200hnsw_index();
    h->ix.prm = *params;
    const int lr = h->ix.host.load(path, params->dim, params->max_elements);
    if (lr) {
        delete h;
        if (lr == -1) { set_error("Cannot open file"); return B200HNSW_E_OPEN; }
        if (lr == -3) { set_error("Not enough memory: loadIndex failed to allocate level0"); return B200HNSW_E_NOMEM; }
        set_error("Index seems to be corrupted or unsupported");
        return B200HNSW_E_CORRUPT;
    }
    h->ix.ef = 10;  // hnswalg.h:795
    rc = h->ix.init_device();
    if (!rc) rc = h->ix.alloc_device(h->ix.host.max_elements);
    if (!rc) rc = h->ix.upload_all();
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
    B200_GUARD_END
}

int b200hnsw_save(b200hnsw_index *h, const char *path) {
    B200_GUARD_BEGIN
    if (!h || !path) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::unique_lock<std::shared_mutex> lk(h->ix.rw);
    h->ix.drain_async();
    int rc = h->ix.flush_locked();
    if (!rc) rc = h->ix.sync_host_mirror();
    if (rc) return rc;
    if (h->ix.host.save(path)) { set_error("Cannot open file"); return B200HNSW_E_OPEN; }
    return 0;
    B200_GUARD_END
}

void b200hnsw_destroy(b200hnsw_index *h) { delete h; }

int b200hnsw_set_ef(b200hnsw_index *h, size_t ef) {
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    h->ix.ef = ef;
    return 0;
}

int b200hnsw_add_batch(b200hnsw_index *h, const float *X, const uint64_t *labels, size_t n) {
    B200_GUARD_BEGIN
    if (!h || (!X && n)) { set_error("null argument"); return B200HNSW_E_ARG; }
    return h->ix.add_batch(X, labels, n);
    B200_GUARD_END
}

int b200hnsw_add_batch_replace_deleted(b200hnsw_index *h, const float *X, const uint64_t *labels, size_t n) {
    B200_GUARD_BEGIN
    if (!h || (!X && n)) { set_error("null argument"); return B200HNSW_E_ARG; }
    if (!h->ix.prm.allow_replace_deleted) {
        set_error("Replacement of deleted elements is disabled in constructor");  // hnswalg.h:955-957
        return B200HNSW_E_STATE;
    }
    return h->ix.add_batch(X, labels, n, true);
    B200_GUARD_END
}

int b200hnsw_flush(b200hnsw_index *h) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.flush();
    B200_GUARD_END
}

int b200hnsw_search_batch(b200hnsw_index *h, const float *Q, size_t nq, size_t k, size_t ef, uint64_t *labels_out,
                          float *dists_out, uint32_t *counts_out, uint32_t *work_out) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    if (nq == 1 && Q && labels_out && dists_out && k)  // one query per call: coalesce concurrent callers into one launch
        return h->ix.search_coalesced(Q, k, ef, labels_out, dists_out, counts_out, work_out);
    return h->ix.search_host(Q, nq, k, ef, labels_out, dists_out, counts_out, work_out);
    B200_GUARD_END
}

int b200hnsw_search_batch_submit(b200hnsw_index *h, const float *Q, size_t nq, size_t k, size_t ef, uint64_t *labels_out,
                                 float *dists_out, uint32_t *counts_out, uint64_t *ticket_out) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.search_submit(Q, nq, k, ef, labels_out, dists_out, counts_out, ticket_out);
    B200_GUARD_END
}

int b200hnsw_search_batch_wait(b200hnsw_index *h, uint64_t ticket) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.search_wait(ticket);
    B200_GUARD_END
}

int b200hnsw_search_batch_filtered(b200hnsw_index *h, const float *Q, size_t nq, size_t k, size_t ef,
                                   const uint8_t *allowed, uint64_t *labels_out, float *dists_out, uint32_t *counts_out) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.search_host(Q, nq, k, ef, labels_out, dists_out, counts_out, nullptr, allowed);
    B200_GUARD_END
}

int b200hnsw_get_labels(b200hnsw_index *h, uint64_t *labels_out, size_t capacity) {
    if (!h || !labels_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::shared_lock<std::shared_mutex> lk(h->ix.rw);
    const b200::HostImage &m = h->ix.host;
    if (capacity < m.cur) { set_error("labels_out is smaller than cur_element_count"); return B200HNSW_E_ARG; }
    for (size_t i = 0; i < m.cur; i++) labels_out[i] = m.label(i);
    return 0;
}

int b200hnsw_search_batch_device(b200hnsw_index *h, const float *dQ, size_t nq, size_t k, size_t ef,
                                 uint64_t *d_labels_out, float *d_dists_out, uint32_t *d_counts_out,
                                 uint32_t *d_work_out, void *cuda_stream) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.search_device(dQ, nq, k, ef, d_labels_out, d_dists_out, d_counts_out, d_work_out,
                               (cudaStream_t)cuda_stream);
    B200_GUARD_END
}

int b200hnsw_get_info(b200hnsw_index *h, b200hnsw_info *o) {
    if (!h || !o) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::shared_lock<std::shared_mutex> lk(h->ix.rw);
    const b200::HostImage &m = h->ix.host;
    memset(o, 0, sizeof(*o));
    o->cur_element_count = m.cur; o->max_elements = m.max_elements; o->num_deleted = m.num_deleted;
    o->dim = m.dim; o->M = m.M; o->maxM = m.maxM; o->maxM0 = m.maxM0; o->ef_construction = m.efc; o->ef = h->ix.ef.load();
    o->size_data_per_element = m.size_data; o->size_links_per_element = m.size_links;
    o->size_links_level0 = m.size_links0; o->offset_data = m.off_data; o->label_offset = m.off_label;
    o->mult = m.mult; o->maxlevel = m.maxlevel; o->enterpoint_node = m.enterpoint;
    o->metric = h->ix.prm.metric; o->storage = h->ix.prm.storage; o->device = h->ix.dev.device;
    return 0;
}

int b200hnsw_get_levels(b200hnsw_index *h, const int32_t **levels_out) {
    if (!h || !levels_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    *levels_out = h->ix.host.levels.data();
    return 0;
}

int b200hnsw_get_linklist(b200hnsw_index *h, uint32_t id, int level, const uint32_t **ptr_out) {
    B200_GUARD_BEGIN
    if (!h || !ptr_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    int rc = 0;
    if (h->ix.has_staged || h->ix.mirror_dirty) {  // refresh the host mirror; otherwise a shared lock is enough
        std::unique_lock<std::shared_mutex> xl(h->ix.rw);
        h->ix.drain_async();
        rc = h->ix.flush_locked();
        if (!rc) rc = h->ix.sync_host_mirror();
        if (rc) return rc;
    }
    std::shared_lock<std::shared_mutex> lk(h->ix.rw);
    const b200::HostImage &m = h->ix.host;
    if (id >= m.cur || level < 0 || level > m.levels[id]) { set_error("no such link list"); return B200HNSW_E_ARG; }
    *ptr_out = m.list(id, level);
    return 0;
    B200_GUARD_END
}

int b200hnsw_get_label(b200hnsw_index *h, uint32_t id, uint64_t *label_out) {
    if (!h || !label_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::shared_lock<std::shared_mutex> lk(h->ix.rw);
    if (id >= h->ix.host.cur) { set_error("internal id out of range"); return B200HNSW_E_ARG; }
    *label_out = h->ix.host.label(id);
    return 0;
}

int b200hnsw_get_data(b200hnsw_index *h, uint32_t id, const float **vec_out) {
    if (!h || !vec_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::shared_lock<std::shared_mutex> lk(h->ix.rw);
    if (id >= h->ix.host.cur) { set_error("internal id out of range"); return B200HNSW_E_ARG; }
    *vec_out = h->ix.host.vec(id);
    return 0;
}

int b200hnsw_get_data_by_label(b200hnsw_index *h, uint64_t label, float *vec_out) {
    B200_GUARD_BEGIN
    if (!h || !vec_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::shared_lock<std::shared_mutex> lk(h->ix.rw);
    const b200::HostImage &m = h->ix.host;
    auto it = m.label_lookup.find(label);
    if (it == m.label_lookup.end() || m.deleted(it->second)) { set_error("Label not found"); return B200HNSW_E_LABEL; }
    memcpy(vec_out, m.vec(it->second), m.dim * 4);
    return 0;
    B200_GUARD_END
}

int b200hnsw_mark_delete(b200hnsw_index *h, uint64_t label) {  // hnswalg.h:853-883
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    std::unique_lock<std::shared_mutex> lk(h->ix.rw);
    h->ix.drain_async();
    b200::HostImage &m = h->ix.host;
    auto it = m.label_lookup.find(label);
    if (it == m.label_lookup.end()) { set_error("Label not found"); return B200HNSW_E_LABEL; }
    unsigned char *f = (unsigned char *)m.rec(it->second) + 2;
    if (*f & 1) { set_error("The requested to delete element is already deleted"); return B200HNSW_E_STATE; }
    *f |= 1;
    m.num_deleted++;
    h->ix.flags_dirty = true;
    return 0;
    B200_GUARD_END
}

int b200hnsw_unmark_delete(b200hnsw_index *h, uint64_t label) {  // hnswalg.h:892-917
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    std::unique_lock<std::shared_mutex> lk(h->ix.rw);
    h->ix.drain_async();
    b200::HostImage &m = h->ix.host;
    auto it = m.label_lookup.find(label);
    if (it == m.label_lookup.end()) { set_error("Label not found"); return B200HNSW_E_LABEL; }
    unsigned char *f = (unsigned char *)m.rec(it->second) + 2;
    if (!(*f & 1)) { set_error("The requested to undelete element is not deleted"); return B200HNSW_E_STATE; }
    *f &= (unsigned char)~1;
    m.num_deleted--;
    h->ix.flags_dirty = true;
    return 0;
    B200_GUARD_END
}

int b200hnsw_resize(b200hnsw_index *h, size_t new_max) {  // hnswalg.h:633-656
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    std::unique_lock<std::shared_mutex> lk(h->ix.rw);
    h->ix.drain_async();
    if (new_max < h->ix.host.cur) {
        set_error("Cannot resize, max element is less than the current number of elements");
        return B200HNSW_E_ARG;
    }
    int rc = h->ix.flush_locked();
    if (!rc) rc = h->ix.sync_host_mirror();
    if (rc) return rc;
    if (!h->ix.host.resize(new_max)) {
        set_error("Not enough memory: resizeIndex failed to allocate base layer");
        return B200HNSW_E_NOMEM;
    }
    rc = h->ix.alloc_device(new_max);
    if (!rc) rc = h->ix.upload_all();
    return rc;
    B200_GUARD_END
}

int b200hnsw_index_file_size(b200hnsw_index *h, uint64_t *bytes_out) {
    if (!h || !bytes_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::shared_lock<std::shared_mutex> lk(h->ix.rw);
    *bytes_out = h->ix.host.file_size();
    return 0;
}

int b200hnsw_get_stats(b200hnsw_index *h, b200hnsw_stats *out) {
    if (!h || !out) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::lock_guard<std::mutex> sg(h->ix.stats_mu);
    *out = h->ix.stats;
    return 0;
}

static int merge_launch(const uint64_t *l, const float *d, size_t ls, size_t ds, size_t shards, size_t nq, size_t k,
                        uint64_t *ol, float *od, void *cuda_stream) {
    if (!l || !d || !ol || !od || !shards || !k) {
        set_error("null argument");
        return B200HNSW_E_ARG;
    }
    if (nq == 0) return 0;
    B200_CUDA_OK(b200::merge_level(l, d, ls, ds, shards, shards, nq, k, ol, od, 0, 0, (cudaStream_t)cuda_stream,
                                   /*sorted_rows=*/true));
    return 0;
}

int b200hnsw_merge_topk_device(const uint64_t *d_labels_in, const float *d_dists_in, size_t shards, size_t nq,
                               size_t k, uint64_t *d_labels_out, float *d_dists_out, void *cuda_stream) {
    B200_GUARD_BEGIN
    return merge_launch(d_labels_in, d_dists_in, nq * k, nq * k, shards, nq, k, d_labels_out, d_dists_out, cuda_stream);
    B200_GUARD_END
}

int b200hnsw_merge_topk_packed_device(const void *d_blocks, size_t block_bytes, size_t shards, size_t nq, size_t k,
                                      uint64_t *d_labels_out, float *d_dists_out, void *cuda_stream) {
    B200_GUARD_BEGIN
    if (block_bytes < nq * k * 12 || block_bytes % 8) { set_error("block_bytes must be >= nq*k*12 and a multiple of 8"); return B200HNSW_E_ARG; }
    const uint64_t *l = (const uint64_t *)d_blocks;
    const float *d = (const float *)((const char *)d_blocks + nq * k * 8);
    return merge_launch(l, d, block_bytes / 8, block_bytes / 4, shards, nq, k, d_labels_out, d_dists_out, cuda_stream);
    B200_GUARD_END
}

// ---- BruteforceSearch ------------------------------------------------------------------------------------
int b200bf_create(const b200hnsw_params *params, b200bf_index **out) {
    B200_GUARD_BEGIN
    if (!out) { set_error("out is null"); return B200HNSW_E_ARG; }
    *out = nullptr;
    int rc = check_params(params);
    if (rc) return rc;
    b200bf_index *h = new b200bf_index();
    rc = h->ix.create(*params);
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
    B200_GUARD_END
}

int b200bf_load(const char *path, const b200hnsw_params *params, b200bf_index **out) {
    B200_GUARD_BEGIN
    if (!out || !path) { set_error("null argument"); return B200HNSW_E_ARG; }
    *out = nullptr;
    int rc = check_params(params);
    if (rc) return rc;
    b200bf_index *h = new b200bf_index();
    rc = h->ix.load(path, *params);
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
    B200_GUARD_END
}

int b200bf_save(b200bf_index *h, const char *path) {
    B200_GUARD_BEGIN
    if (!h || !path) { set_error("null argument"); return B200HNSW_E_ARG; }
    if (h->ix.host.save(path)) { set_error("Cannot open file"); return B200HNSW_E_OPEN; }
    return 0;
    B200_GUARD_END
}

void b200bf_destroy(b200bf_index *h) { delete h; }

int b200bf_add_batch(b200bf_index *h, const float *X, const uint64_t *labels, size_t n) {
    B200_GUARD_BEGIN
    if (!h || (!X && n)) { set_error("null argument"); return B200HNSW_E_ARG; }
    return h->ix.add_batch(X, labels, n);
    B200_GUARD_END
}

int b200bf_remove(b200bf_index *h, uint64_t label) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.remove(label);
    B200_GUARD_END
}

int b200bf_search_batch(b200bf_index *h, const float *Q, size_t nq, size_t k, uint64_t *labels_out,
                        float *dists_out, uint32_t *counts_out) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.search_host(Q, nq, k, labels_out, dists_out, counts_out);
    B200_GUARD_END
}

/* searchKnn with a BaseFilterFunctor (bruteforce.h:106-135, filter at :114 and :121) */
int b200bf_search_batch_filtered(b200bf_index *h, const float *Q, size_t nq, size_t k, const uint8_t *allowed,
                                 uint64_t *labels_out, float *dists_out, uint32_t *counts_out) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.search_host(Q, nq, k, labels_out, dists_out, counts_out, allowed);
    B200_GUARD_END
}

int b200bf_get_labels(b200bf_index *h, uint64_t *labels_out, size_t capacity) {
    if (!h || !labels_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::lock_guard<std::mutex> g(h->ix.mu);
    const size_t n = h->ix.host.cur;
    if (capacity < n) { set_error("labels_out is smaller than the element count"); return B200HNSW_E_ARG; }
    for (size_t i = 0; i < n; i++) labels_out[i] = h->ix.host.label(i);
    return 0;
}

int b200bf_search_batch_device(b200bf_index *h, const float *dQ, size_t nq, size_t k, uint64_t *d_labels_out,
                               float *d_dists_out, uint32_t *d_counts_out, void *cuda_stream) {
    B200_GUARD_BEGIN
    if (!h) { set_error("null handle"); return B200HNSW_E_ARG; }
    return h->ix.search_device(dQ, nq, k, d_labels_out, d_dists_out, d_counts_out, (cudaStream_t)cuda_stream);
    B200_GUARD_END
}

int b200bf_count(b200bf_index *h, uint64_t *count_out) {
    if (!h || !count_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    *count_out = h->ix.host.cur;
    return 0;
}

int b200bf_get_stats(b200bf_index *h, b200hnsw_stats *out) {
    if (!h || !out) { set_error("null argument"); return B200HNSW_E_ARG; }
    *out = h->ix.stats;
    return 0;
}

}  // extern "C"

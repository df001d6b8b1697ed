// sharded.cu -- one process, several GPUs: the data set split into sub-indexes, one per device, every query searched on
// every shard, per-shard top-k merged on the root device (north_star: "the dataset is sharded, with one
// sub-index per GPU, queries broadcast and per-shard top-k merged").
//
// Exchange.  The processes-per-GPU form of this path (research_new_hnsw_b200/sharded.py under torchrun) moves the result
// rows with one NCCL all_gather.  Inside ONE process no collective is needed: with peer access enabled every shard's
// search kernel writes its [nq][k] rows STRAIGHT into its block of the root device's packed buffer -- P2P stores over
// NVLink from the kernel's own epilogue, 12 bytes per result row -- an event per shard orders them before the merge
// kernel on the root device.  Without peer access (or for shards that share the root device) the block is written
// locally and copied with cudaMemcpyPeerAsync.  Nothing here computes on the host.
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "hnsw_index.cuh"
#include "merge_launch.cuh"

struct b200hnsw_sharded {
    std::vector<b200hnsw_index *> shards;
    std::vector<int> devices;
    int root = 0;
    std::vector<char> direct;            // shard writes into the root's buffer itself (same device or peer access)
    std::mutex mu;
    uint64_t next_label = 0;
    // root device
    unsigned char *blocks = nullptr;     // [n_shards][block_bytes]
    uint64_t *outL = nullptr;
    float *outD = nullptr;
    size_t cap_q = 0, cap_k = 0, block_bytes = 0;
    cudaStream_t root_stream = nullptr;
    // per shard
    std::vector<float *> dQ;
    std::vector<unsigned char *> local_block;
    std::vector<cudaStream_t> st;
    std::vector<cudaEvent_t> ev;
    double last_ms = 0.0;

    ~b200hnsw_sharded() {
        for (size_t s = 0; s < shards.size(); s++) {
            cudaSetDevice(devices[s]);
            if (s < dQ.size()) cudaFree(dQ[s]);
            if (s < local_block.size()) cudaFree(local_block[s]);
            if (s < st.size() && st[s]) cudaStreamDestroy(st[s]);
            if (s < ev.size() && ev[s]) cudaEventDestroy(ev[s]);
        }
        cudaSetDevice(root);
        cudaFree(blocks); cudaFree(outL); cudaFree(outD);
        if (root_stream) cudaStreamDestroy(root_stream);
        for (b200hnsw_index *h : shards) b200hnsw_destroy(h);
    }
};

using b200::set_error;

namespace {

int finish_setup(b200hnsw_sharded *g) {
    const size_t n = g->shards.size();
    g->root = g->devices[0];
    g->direct.assign(n, 0);
    g->dQ.assign(n, nullptr);
    g->local_block.assign(n, nullptr);
    g->st.assign(n, nullptr);
    g->ev.assign(n, nullptr);
    for (size_t s = 0; s < n; s++) {
        const int d = g->devices[s];
        B200_CUDA_OK(cudaSetDevice(d));
        B200_CUDA_OK(cudaStreamCreateWithFlags(&g->st[s], cudaStreamNonBlocking));
        B200_CUDA_OK(cudaEventCreateWithFlags(&g->ev[s], cudaEventDisableTiming));
        if (d == g->root) {
            g->direct[s] = 1;
            continue;
        }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, d, g->root) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->root, 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) g->direct[s] = 1;
            cudaGetLastError();
        }
        if (getenv("B200HNSW_SHARD_NO_P2P")) g->direct[s] = 0;
    }
    B200_CUDA_OK(cudaSetDevice(g->root));
    B200_CUDA_OK(cudaStreamCreateWithFlags(&g->root_stream, cudaStreamNonBlocking));
    return 0;
}

int ensure_buffers(b200hnsw_sharded *g, size_t nq, size_t k, size_t dim) {
    if (nq <= g->cap_q && k <= g->cap_k) return 0;
    const size_t q = std::max(nq, g->cap_q), kk = std::max(k, g->cap_k), n = g->shards.size();
    B200_CUDA_OK(cudaSetDevice(g->root));
    cudaFree(g->blocks); cudaFree(g->outL); cudaFree(g->outD);
    g->blocks = nullptr; g->outL = nullptr; g->outD = nullptr;
    g->cap_q = g->cap_k = 0;
    const size_t bb = (q * kk * 12 + 7) / 8 * 8;
    B200_CUDA_OK(cudaMalloc(&g->blocks, n * bb));
    B200_CUDA_OK(cudaMalloc(&g->outL, q * kk * 8));
    B200_CUDA_OK(cudaMalloc(&g->outD, q * kk * 4));
    for (size_t s = 0; s < n; s++) {
        B200_CUDA_OK(cudaSetDevice(g->devices[s]));
        cudaFree(g->dQ[s]); cudaFree(g->local_block[s]);
        g->dQ[s] = nullptr; g->local_block[s] = nullptr;
        B200_CUDA_OK(cudaMalloc(&g->dQ[s], q * dim * 4));
        if (!g->direct[s]) B200_CUDA_OK(cudaMalloc(&g->local_block[s], bb));
    }
    g->cap_q = q;
    g->cap_k = kk;
    return 0;
}

}  // namespace

#define B200_SH_BEGIN try {
#define B200_SH_END                                            \
    }                                                          \
    catch (const std::bad_alloc &) {                           \
        set_error("Not enough memory");                        \
        return B200HNSW_E_NOMEM;                               \
    }                                                          \
    catch (const std::exception &e) {                          \
        set_error(std::string("internal error: ") + e.what()); \
        return B200HNSW_E_STATE;                               \
    }

extern "C" {

int b200hnsw_sharded_create(const b200hnsw_params *params, const int *devices, size_t n_shards, b200hnsw_sharded **out) {
    B200_SH_BEGIN
    if (!out || !params || !devices || n_shards == 0) { set_error("null argument or no shards"); return B200HNSW_E_ARG; }
    *out = nullptr;
    std::unique_ptr<b200hnsw_sharded> g(new b200hnsw_sharded());
    for (size_t s = 0; s < n_shards; s++) {
        b200hnsw_params p = *params;
        p.device = devices[s];
        b200hnsw_index *h = nullptr;
        const int rc = b200hnsw_create(&p, &h);
        if (rc) return rc;
        g->shards.push_back(h);
        g->devices.push_back(h->ix.dev.device);
    }
    const int rc = finish_setup(g.get());
    if (rc) return rc;
    *out = g.release();
    return 0;
    B200_SH_END
}

int b200hnsw_sharded_load(const char *const *paths, const b200hnsw_params *params, const int *devices, size_t n_shards,
                          b200hnsw_sharded **out) {
    B200_SH_BEGIN
    if (!out || !params || !devices || !paths || n_shards == 0) { set_error("null argument or no shards"); return B200HNSW_E_ARG; }
    *out = nullptr;
    std::unique_ptr<b200hnsw_sharded> g(new b200hnsw_sharded());
    for (size_t s = 0; s < n_shards; s++) {
        b200hnsw_params p = *params;
        p.device = devices[s];
        b200hnsw_index *h = nullptr;
        const int rc = b200hnsw_load(paths[s], &p, &h);
        if (rc) return rc;
        g->shards.push_back(h);
        g->devices.push_back(h->ix.dev.device);
        g->next_label += h->ix.host.cur;
    }
    const int rc = finish_setup(g.get());
    if (rc) return rc;
    *out = g.release();
    return 0;
    B200_SH_END
}

void b200hnsw_sharded_destroy(b200hnsw_sharded *g) { delete g; }

int b200hnsw_sharded_num_shards(b200hnsw_sharded *g, size_t *n_out) {
    if (!g || !n_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    *n_out = g->shards.size();
    return 0;
}

int b200hnsw_sharded_get_shard(b200hnsw_sharded *g, size_t shard, b200hnsw_index **out) {
    if (!g || !out || shard >= g->shards.size()) { set_error("no such shard"); return B200HNSW_E_ARG; }
    *out = g->shards[shard];
    return 0;
}

int b200hnsw_sharded_count(b200hnsw_sharded *g, uint64_t *count_out) {
    if (!g || !count_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    uint64_t c = 0;
    for (b200hnsw_index *h : g->shards) {
        b200hnsw_info o;
        const int rc = b200hnsw_get_info(h, &o);
        if (rc) return rc;
        c += o.cur_element_count;
    }
    *count_out = c;
    return 0;
}

int b200hnsw_sharded_add_batch(b200hnsw_sharded *g, const float *X, const uint64_t *labels, size_t n) {
    B200_SH_BEGIN
    if (!g || (!X && n)) { set_error("null argument"); return B200HNSW_E_ARG; }
    std::lock_guard<std::mutex> lk(g->mu);
    const size_t ns = g->shards.size(), dim = g->shards[0]->ix.host.dim;
    std::vector<std::vector<float>> rows(ns);
    std::vector<std::vector<uint64_t>> labs(ns);
    for (size_t i = 0; i < n; i++) {
        const uint64_t lab = labels ? labels[i] : g->next_label + i;
        const size_t s = (size_t)(lab % ns);  // a label always lives on the same shard, so re-adding it updates it there
        rows[s].insert(rows[s].end(), X + i * dim, X + (i + 1) * dim);
        labs[s].push_back(lab);
    }
    if (!labels) g->next_label += n;
    for (size_t s = 0; s < ns; s++) {
        if (labs[s].empty()) continue;
        const int rc = b200hnsw_add_batch(g->shards[s], rows[s].data(), labs[s].data(), labs[s].size());
        if (rc) return rc;
    }
    return 0;
    B200_SH_END
}

int b200hnsw_sharded_flush(b200hnsw_sharded *g) {
    if (!g) { set_error("null handle"); return B200HNSW_E_ARG; }
    for (b200hnsw_index *h : g->shards) {
        const int rc = b200hnsw_flush(h);
        if (rc) return rc;
    }
    return 0;
}

int b200hnsw_sharded_save(b200hnsw_sharded *g, const char *const *paths) {
    if (!g || !paths) { set_error("null argument"); return B200HNSW_E_ARG; }
    for (size_t s = 0; s < g->shards.size(); s++) {
        const int rc = b200hnsw_save(g->shards[s], paths[s]);
        if (rc) return rc;
    }
    return 0;
}

int b200hnsw_sharded_search_batch(b200hnsw_sharded *g, const float *Q, size_t nq, size_t k, size_t ef,
                                  uint64_t *labels_out, float *dists_out, uint32_t *counts_out) {
    B200_SH_BEGIN
    if (!g) { set_error("null handle"); return B200HNSW_E_ARG; }
    if (nq == 0) return 0;
    if (!Q || !labels_out || !dists_out || k == 0) { set_error("search: null pointer or k == 0"); return B200HNSW_E_ARG; }
    std::lock_guard<std::mutex> lk(g->mu);
    const size_t ns = g->shards.size(), dim = g->shards[0]->ix.host.dim;
    int rc = ensure_buffers(g, nq, k, dim);
    if (rc) return rc;
    const size_t bb = (nq * k * 12 + 7) / 8 * 8;  // block size of THIS call (the merge reads blocks at this pitch)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    B200_CUDA_OK(cudaSetDevice(g->root));
    B200_CUDA_OK(cudaEventCreate(&t0));
    B200_CUDA_OK(cudaEventCreate(&t1));
    B200_CUDA_OK(cudaEventRecord(t0, g->root_stream));
    for (size_t s = 0; s < ns; s++) {
        B200_CUDA_OK(cudaSetDevice(g->devices[s]));
        B200_CUDA_OK(cudaStreamWaitEvent(g->st[s], t0, 0));
        B200_CUDA_OK(cudaMemcpyAsync(g->dQ[s], Q, nq * dim * 4, cudaMemcpyHostToDevice, g->st[s]));
        unsigned char *blk = g->direct[s] ? g->blocks + s * bb : g->local_block[s];
        rc = g->shards[s]->ix.search_device(g->dQ[s], nq, k, ef, (uint64_t *)blk, (float *)(blk + nq * k * 8), nullptr, nullptr,
                                            g->st[s]);
        if (rc) break;
        if (!g->direct[s])
            B200_CUDA_OK(cudaMemcpyPeerAsync(g->blocks + s * bb, g->root, blk, g->devices[s], bb, g->st[s]));
        B200_CUDA_OK(cudaEventRecord(g->ev[s], g->st[s]));
    }
    if (rc == 0) {
        B200_CUDA_OK(cudaSetDevice(g->root));
        for (size_t s = 0; s < ns; s++) B200_CUDA_OK(cudaStreamWaitEvent(g->root_stream, g->ev[s], 0));
        B200_CUDA_OK(b200::merge_level((const uint64_t *)g->blocks, (const float *)(g->blocks + nq * k * 8), bb / 8, bb / 4, ns, ns,
                                       nq, k, g->outL, g->outD, 0, 0, g->root_stream, /*sorted_rows=*/true));
        B200_CUDA_OK(cudaEventRecord(t1, g->root_stream));
        B200_CUDA_OK(cudaMemcpyAsync(labels_out, g->outL, nq * k * 8, cudaMemcpyDeviceToHost, g->root_stream));
        B200_CUDA_OK(cudaMemcpyAsync(dists_out, g->outD, nq * k * 4, cudaMemcpyDeviceToHost, g->root_stream));
        B200_CUDA_OK(cudaStreamSynchronize(g->root_stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        g->last_ms = ms;
        if (counts_out)
            for (size_t i = 0; i < nq; i++) {
                uint32_t c = 0;
                while (c < k && labels_out[i * k + c] != 0xFFFFFFFFFFFFFFFFull) c++;
                counts_out[i] = c;
            }
    } else {
        for (size_t s = 0; s < ns; s++) { cudaSetDevice(g->devices[s]); cudaStreamSynchronize(g->st[s]); }
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return rc;
    B200_SH_END
}

int b200hnsw_sharded_last_ms(b200hnsw_sharded *g, double *ms_out) {
    if (!g || !ms_out) { set_error("null argument"); return B200HNSW_E_ARG; }
    *ms_out = g->last_ms;
    return 0;
}

}  // extern "C"

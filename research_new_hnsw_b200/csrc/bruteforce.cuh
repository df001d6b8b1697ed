// bruteforce.cuh -- BruteforceSearch<float> replacement (reference: hnswlib/bruteforce.h:10-172).
//
// searchKnn there is a linear scan with a max-heap of (dist, label) pairs that keeps `dist <= lastdist`
// (bruteforce.h:106-135): its result is a pure function of the data -- the k lexicographically smallest
// (dist, label) pairs.  The exact-scan kernel below evaluates every distance in the
// summation order of the reference's shipped SSE kernels (space_l2.h:97-143, space_ip.h:255-303: four lane
// accumulators over elements i = l mod 4, separate multiply and add, ((T0+T1)+T2)+T3), so distances are
// bit-identical to the CPU build and ids match exactly, ties included.
#pragma once
#include <mutex>

#include "../../include/b200hnsw.h"
#include "common.cuh"
#include "host_image.hpp"

namespace b200 {

// Device state of the tensor-core candidate generator (bf_tensor.cu).
struct BruteTensor {
    void *xb = nullptr;        // bf16 [rows_pad][kp]
    float *xn2 = nullptr;      // [rows_pad] squared norms
    void *qb = nullptr;        // bf16 [q_cap][kp]
    float *qn2 = nullptr, *thr = nullptr, *tabB = nullptr, *tabT = nullptr;
    uint32_t *panelmin = nullptr, *cand = nullptr, *cand_cnt = nullptr, *overflow = nullptr;
    size_t kp = 0, rows_pad = 0, q_cap = 0, cap = 0, cap_floor = 0, pm_elems = 0, last_candidates = 0;
    void release();
};

struct BruteIndex {
    b200hnsw_params prm{};
    HostBrute host;
    int device = 0;
    size_t d4 = 0, cap = 0;
    float4 *dX = nullptr;       // [cap][d4] zero-padded rows
    uint64_t *dLabels = nullptr;
    std::mutex mu;
    b200hnsw_stats stats{};
    BruteTensor tz;
    // row mask of the filtered search in progress (BaseFilterFunctor verdicts, bruteforce.h:114,121); null otherwise
    uint8_t *dMask = nullptr;
    size_t mask_cap = 0;
    const uint8_t *cur_mask = nullptr;
    size_t cur_mask_rows = 0;
    int last_path = 0;  // 0 = tiled exact scan, 1 = tensor-core candidates + exact re-rank, 2 = streaming exact scan
    // scratch
    float *dQ = nullptr;
    uint64_t *dOutL = nullptr, *dPartL = nullptr, *dPart2L = nullptr;
    float *dOutD = nullptr, *dPartD = nullptr, *dPart2D = nullptr;
    uint32_t *dCounts = nullptr;
    size_t scratch_q = 0, scratch_k = 0, part_elems = 0, part2_elems = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    ~BruteIndex();
    int create(const b200hnsw_params &p);
    int load(const char *path, const b200hnsw_params &p);
    int init_device();
    int upload_rows(size_t first, size_t count);
    int add_batch(const float *X, const uint64_t *labels, size_t n);
    int remove(uint64_t label);
    int ensure_part(size_t elems);
    int search_device(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc, cudaStream_t st);
    int search_host(const float *Q, size_t nq, size_t k, uint64_t *labels, float *dists, uint32_t *counts,
                    const uint8_t *allowed = nullptr);
    // bf_tensor.cu
    int tensor_sync_rows(size_t first, size_t count);
    int search_tensor(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc, cudaStream_t st);
    int search_scan(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc, cudaStream_t st);
    // bf_stream.cu
    int search_stream(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc, cudaStream_t st);
    int ensure_part2(size_t elems);
};

}  // namespace b200

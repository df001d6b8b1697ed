// device_index.cuh -- the index as it lives in HBM (SoA), and its upload from the reference-layout host mirror.
//
// HBM layout (sized for max_elements so staged insertions need no reallocation):
//   vec      float4[cap][d4]     zero-padded rows, 16-byte aligned (reference rows sit at byte 4+4*maxM0 of a
//                                652/780-byte record and are only 4-byte aligned: hnswalg.h:120-124,202-204)
//   links0   u32[cap][maxM0]     level-0 neighbours, unused slots = kEmpty (the u16 count of the reference record is
//                                implied; 256 B per node at M=32, two aligned cache lines)
//   up_base  u32[cap]            index of the node's level-1 list in links_up, or kEmpty for level-0-only nodes
//   links_up u32[lists][maxM]    upper-level lists, node-major then level-major, unused slots = kEmpty
//   labels   u64[cap]            external labels (hnswalg.h:186-190)
#pragma once
#include "../../include/b200hnsw.h"
#include "common.cuh"
#include "host_image.hpp"

namespace b200 {

struct DeviceIndex {
    int device = 0;
    size_t cap = 0, n = 0, dim = 0, d4 = 0, maxM = 0, maxM0 = 0;
    size_t up_lists = 0, up_lists_cap = 0;
    float4 *vec = nullptr;
    uint32_t *links0 = nullptr, *up_base = nullptr, *links_up = nullptr;
    uint64_t *labels = nullptr;
    uint32_t *err_flag = nullptr;
    uint8_t *flags = nullptr;   // delete marks, allocated on first use
    uint4 *vec16 = nullptr;     // bf16 copy of the vectors [cap][d16] (storage variant B200HNSW_BF16)
    size_t d16 = 0;

    void release() {
        cudaFree(vec); cudaFree(links0); cudaFree(up_base); cudaFree(links_up); cudaFree(labels); cudaFree(err_flag); cudaFree(flags); cudaFree(vec16);
        vec = nullptr; links0 = up_base = links_up = nullptr; labels = nullptr; err_flag = nullptr; flags = nullptr; vec16 = nullptr;
        cap = n = 0;
    }
    size_t bytes() const {
        return cap * d4 * 16 + cap * maxM0 * 4 + cap * 4 + up_lists_cap * maxM * 4 + cap * 8;
    }
};

// De-interleave `count` reference records starting at element `first` (raw bytes in `raw`) into the SoA arrays.
// One warp per record; 4-byte granularity because records are only 4-byte aligned.
static __global__ void deinterleave_kernel(const uint32_t *__restrict__ raw, size_t rec_words, uint32_t first, uint32_t count,
                                    uint32_t maxM0, uint32_t dim, uint32_t d4, uint32_t n_valid, float *__restrict__ vec,
                                    uint32_t *__restrict__ links0, uint64_t *__restrict__ labels,
                                    uint32_t *__restrict__ err_flag) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= count) return;
    const uint32_t *r = raw + (size_t)w * rec_words;
    const uint32_t id = first + w;
    const uint32_t cnt = r[0] & 0xFFFFu;
    for (uint32_t j = lane; j < maxM0; j += 32) {
        uint32_t v = kEmpty;
        if (j < cnt) {
            v = r[1 + j];
            if (v >= n_valid) { atomicOr(err_flag, 1u); v = kEmpty; }  // "cand error" (hnswalg.h:1292)
        }
        links0[(size_t)id * maxM0 + j] = v;
    }
    if (cnt > maxM0 && lane == 0) atomicOr(err_flag, 2u);
    const uint32_t *rv = r + 1 + maxM0;
    for (uint32_t j = lane; j < d4 * 4; j += 32)
        vec[(size_t)id * d4 * 4 + j] = j < dim ? __uint_as_float(rv[j]) : 0.f;
    if (lane == 0) labels[id] = (uint64_t)rv[dim] | ((uint64_t)rv[dim + 1] << 32);
}

// fp32 rows -> bf16 rows (round to nearest even), 8 elements per 128-bit chunk, zero padded.  One thread per chunk.
static __global__ void rows_to_bf16_kernel(const float4 *__restrict__ vec, uint32_t d4, uint32_t d16, uint32_t first,
                                           uint32_t count, uint4 *__restrict__ vec16) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)count * d16) return;
    const uint32_t r = first + (uint32_t)(i / d16), c = (uint32_t)(i % d16);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (2 * c < d4) a = vec[(size_t)r * d4 + 2 * c];
    if (2 * c + 1 < d4) b = vec[(size_t)r * d4 + 2 * c + 1];
    auto pack = [](float lo, float hi) -> uint32_t {
        uint32_t l = __float_as_uint(lo), h = __float_as_uint(hi);
        l = (l + 0x7FFFu + ((l >> 16) & 1u)) >> 16;
        h = (h + 0x7FFFu + ((h >> 16) & 1u)) >> 16;
        return l | (h << 16);
    };
    vec16[(size_t)r * d16 + c] = make_uint4(pack(a.x, a.y), pack(a.z, a.w), pack(b.x, b.y), pack(b.z, b.w));
}

// updatePoint: new vectors / labels of `count` existing elements, staged as [count][d4] float4 rows, scattered to their
// places (and to the bf16 copy when there is one).  One warp per row.
static __global__ void scatter_rows_kernel(const float4 *__restrict__ rows, const uint64_t *__restrict__ row_labels,
                                           const uint32_t *__restrict__ ids, uint32_t count, uint32_t d4, uint32_t d16,
                                           float4 *__restrict__ vec, uint64_t *__restrict__ labels, uint4 *__restrict__ vec16,
                                           uint8_t *__restrict__ flags) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (w >= count) return;
    const uint32_t id = ids[w];
    for (uint32_t c = lane; c < d4; c += 32) vec[(size_t)id * d4 + c] = rows[(size_t)w * d4 + c];
    if (lane == 0) {
        labels[id] = row_labels[w];
        if (flags) flags[id] = 0;  // an updated element is live (unmarkDeletedInternal precedes updatePoint)
    }
    if (vec16) {
        auto pack = [](float lo, float hi) -> uint32_t {
            uint32_t l = __float_as_uint(lo), h = __float_as_uint(hi);
            l = (l + 0x7FFFu + ((l >> 16) & 1u)) >> 16;
            h = (h + 0x7FFFu + ((h >> 16) & 1u)) >> 16;
            return l | (h << 16);
        };
        for (uint32_t c = lane; c < d16; c += 32) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (2 * c < d4) a = rows[(size_t)w * d4 + 2 * c];
            if (2 * c + 1 < d4) b = rows[(size_t)w * d4 + 2 * c + 1];
            vec16[(size_t)id * d16 + c] = make_uint4(pack(a.x, a.y), pack(a.z, a.w), pack(b.x, b.y), pack(b.z, b.w));
        }
    }
}

}  // namespace b200

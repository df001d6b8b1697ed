// Micro-benchmark: what HBM bandwidth can random 512-byte row gathers reach on this GPU?  (practical ceiling of the
// graph-search gather, to put next to the STREAM-style copy peak in MEASURED_PEAKS.json)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/gather_bw.cu -o /tmp/gather_bw && /tmp/gather_bw
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// each group of 8 lanes reads ROWS rows of row_f4 float4 (4 per lane for 512 B), INFLIGHT rows at a time
template <int INFLIGHT>
__global__ void gather(const float4 *__restrict__ tab, const unsigned *__restrict__ idx, size_t n_idx, int row_f4, float *out) {
    const size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 8;
    const int sub = threadIdx.x & 7;
    const size_t groups = (size_t)gridDim.x * blockDim.x / 8;
    float acc = 0.f;
    for (size_t i = g * INFLIGHT; i + INFLIGHT <= n_idx; i += groups * INFLIGHT) {
        float4 v[INFLIGHT][4];
#pragma unroll
        for (int r = 0; r < INFLIGHT; r++) {
            const float4 *row = tab + (size_t)idx[i + r] * row_f4;
#pragma unroll
            for (int c = 0; c < 4; c++) v[r][c] = ldg_stream(row + sub + c * 8);
        }
#pragma unroll
        for (int r = 0; r < INFLIGHT; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc += v[r][c].x + v[r][c].y + v[r][c].z + v[r][c].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    const size_t rows = 1u << 21;  // 2M rows x 512 B = 1 GB  (>> 126 MB L2)
    const int row_f4 = 32;
    const size_t n_idx = 1u << 24;  // 16M gathers = 8 GB of traffic
    float4 *tab; unsigned *idx; float *out;
    cudaMalloc(&tab, rows * row_f4 * 16); cudaMemset(tab, 0, rows * row_f4 * 16);
    cudaMalloc(&idx, n_idx * 4); cudaMalloc(&out, 4);
    std::vector<unsigned> h(n_idx);
    unsigned long long s = 88172645463325252ull;
    for (size_t i = 0; i < n_idx; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (unsigned)(s % rows); }
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](auto kern, const char *name, int blocks_per_sm, int threads) {
        int grid = 148 * blocks_per_sm;
        for (int it = 0; it < 2; it++) kern<<<grid, threads>>>(tab, idx, n_idx, row_f4, out);
        cudaEventRecord(e0);
        for (int it = 0; it < 3; it++) kern<<<grid, threads>>>(tab, idx, n_idx, row_f4, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        printf("%-28s blocks/SM %2d x %4d thr: %.3f ms  %.0f GB/s\n", name, blocks_per_sm, threads, ms, n_idx * 512.0 / ms / 1e6);
    };
    run(gather<1>, "1 row/group in flight", 8, 256);
    run(gather<2>, "2 rows/group in flight", 8, 128);
    run(gather<2>, "2 rows/group in flight", 8, 256);
    run(gather<4>, "4 rows/group in flight", 8, 256);
    run(gather<4>, "4 rows/group in flight", 16, 128);
    run(gather<8>, "8 rows/group in flight", 4, 256);
    // sequential copy-like read of the same table for reference
    return 0;
}

// common.cuh -- shared device/host helpers for libb200hnsw (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace b200 {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;      // empty link slot / empty hash slot / "no upper lists"
constexpr uint32_t kExpanded = 0x80000000u;   // bit 31 of the id word of a candidate key: already expanded
constexpr uint32_t kIdMask = 0x7FFFFFFFu;
constexpr uint64_t kKeyMask = 0xFFFFFFFF7FFFFFFFull;  // key without the expanded flag

void set_error(const std::string &msg);  // thread-local last error (capi.cu)

#define B200_CUDA_OK(expr)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::b200::set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr); \
            return B200HNSW_E_CUDA;                                                                \
        }                                                                                          \
    } while (0)

// Monotone map float -> uint32 (total order incl. negatives: inner-product distances may be < 0).
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
    return ((uint64_t)f2ord(d) << 32) | (uint64_t)id;
}

#ifdef __CUDACC__
// 128-bit streaming load: vectors are gathered at random, caching them in L1 only evicts useful lines.
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
#endif

}  // namespace b200

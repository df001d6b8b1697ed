// bf_tensor.cu -- BruteforceSearch on the 5th-generation tensor cores: a hand-written tcgen05 GEMM with the top-k
// selection fused into its epilogue, followed by an exact fp32 re-rank in the reference's summation order.
//
// Reference replaced: hnswlib/bruteforce.h:106-135 (linear scan + max-heap); result = the k lexicographically smallest
// (dist, label) pairs.  Tensor cores cannot produce fp32-exact distances, so the GEMM only GENERATES CANDIDATES with a
// rigorous error bound, and bit-exactness comes from re-evaluating the candidates with the arithmetic of
// bf_scan_kernel (reference SSE order):
//
//   S = X_bf16 . Q_bf16^T            tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), 128 rows x 256 queries per
//                                    accumulator, K streamed in 64-element (128 B, SWIZZLE_128B) chunks by TMA
//   key(q,x)  = -S (inner product)   or   |x|^2 + |q|^2 - 2 S (L2)          approximate, to MINIMISE
//   E(q,x)    = c |q| |x| (+ delta (|x|^2+|q|^2) for L2),  c = 2^-8 + 2^-12  >= |key - true key|
//               (bf16 rounding is 2^-9 relative per operand; fp32 accumulation slack in the 2^-12)
//   pass 1    every `stride`-th 128-row panel: panelmin[panel][q] = min_x key + E.  Each panel holds a row whose TRUE
//             key is <= its panelmin, so U(q) = k-th smallest panelmin is an upper bound of the true k-th best key.
//   pass 2    all panels: emit row x for query q iff key - E <= U(q)  (a superset of the true top-k; ballot-compacted,
//             one atomicAdd per warp and query column).
//   re-rank   exact distances of the candidates, k smallest (dist, label), closest first.
//
// Roles in the 192-thread CTA (Blackwell playbook): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// MMA issuer, warps 2-5 = epilogue (tcgen05.ld of their 32-lane quarter).  Pipelines: 4-stage smem ring (full/empty
// mbarriers, tcgen05.commit frees a stage), 2 accumulator buffers of 256 TMEM columns (tmem_full / tmem_empty).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "bruteforce.cuh"

namespace b200 {

constexpr int kGM = 128;       // rows per tile (UMMA M)
constexpr int kGN = 256;       // queries per tile (UMMA N)
constexpr int kGK = 64;        // K elements per stage (128 bytes of bf16 = one swizzle atom row)
constexpr int kGStages = 4;
constexpr int kGThreads = 192;
constexpr uint32_t kStageA = kGM * kGK * 2;   // 16 KB
constexpr uint32_t kStageB = kGN * kGK * 2;   // 32 KB
constexpr uint32_t kStageBytes = kStageA + kStageB;
constexpr uint32_t kGemmSmem = kGStages * kStageBytes + 1024 /*align*/ + 8192 /*barriers + per-tile tables*/;
constexpr float kErrC = 0.00390625f + 0.000244140625f;   // 2^-8 + 2^-12
constexpr float kErrDelta = 4e-6f;

struct GemmArgs {
    const float *xn2;       // [rows] squared norms (fp32)
    const float *qn2;       // [nq_pad]
    const float *thr;       // [nq_pad] U(q)  (pass 2)
    uint32_t *panelmin;     // [sampled panels][nq_pad] ordered floats (pass 1)
    uint32_t *cand;         // [nq][cap] row ids (pass 2)
    uint32_t *cand_cnt;     // [nq]
    uint32_t n, nq, nq_pad, kchunks, panels, stride, qtiles, cap;
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
        "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 8-row atoms 1024 B apart (cute/arch/mma_sm100_desc.hpp
// SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout 2 <<61).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 at bit 17, M>>4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGN >> 3) << 17) | ((uint32_t)(kGM >> 4) << 24);

template <int METRIC, int MODE>
__global__ void __launch_bounds__(kGThreads, 1) bf_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const GemmArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *tail = smem + kGStages * kStageBytes;
    uint64_t *full = (uint64_t *)tail;            // [kGStages]
    uint64_t *empty = full + kGStages;            // [kGStages]
    uint64_t *tfull = empty + kGStages;           // [2]
    uint64_t *tempty = tfull + 2;                 // [2]
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);
    float *s_qn2 = (float *)(tail + 256);         // [kGN]
    float *s_cq = s_qn2 + kGN;                    // [kGN]  c * |q|
    float *s_thr = s_cq + kGN;                    // [kGN]
    uint32_t *s_colmin = (uint32_t *)(s_thr + kGN);  // [kGN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sampled = (a.panels + a.stride - 1) / a.stride;
    const uint32_t items = sampled * a.qtiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGStages; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 256-column accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t it = blockIdx.x; it < items; it += gridDim.x) {
                const uint32_t p = (it / a.qtiles) * a.stride, t = it % a.qtiles;
                for (uint32_t kc = 0; kc < a.kchunks; kc++) {
                    mbar_wait(empty + stage, phase ^ 1);
                    mbar_expect_tx(full + stage, kStageBytes);
                    unsigned char *sa = smem + stage * kStageBytes;
                    tma_load_2d(sa, &tmA, full + stage, (int)(kc * kGK), (int)(p * kGM));
                    tma_load_2d(sa + kStageA, &tmB, full + stage, (int)(kc * kGK), (int)(t * kGN));
                    if (++stage == kGStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, buf = 0, bphase = 0;
            for (uint32_t it = blockIdx.x; it < items; it += gridDim.x) {
                mbar_wait(tempty + buf, bphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * kGN;
                for (uint32_t kc = 0; kc < a.kchunks; kc++) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = umma_desc(sa), db = umma_desc(sa + kStageA);
#pragma unroll
                    for (uint32_t k4 = 0; k4 < kGK / 16; k4++)  // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
                        tc_mma_bf16(d_tmem, da + 2 * k4, db + 2 * k4, kIdesc, (kc | k4) != 0);
                    tc_commit(empty + stage);  // frees the smem stage when these MMAs retire
                    if (++stage == kGStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull + buf);  // accumulator complete
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int et = threadIdx.x - 64;  // 0..127
        uint32_t buf = 0, bphase = 0;
        for (uint32_t it = blockIdx.x; it < items; it += gridDim.x) {
            const uint32_t pi = it / a.qtiles, p = pi * a.stride, t = it % a.qtiles;
            // per-tile query tables
            for (int j = et; j < kGN; j += 128) {
                const uint32_t q = t * kGN + j;
                const float qn2 = a.qn2[q];
                s_qn2[j] = qn2;
                s_cq[j] = kErrC * sqrtf(qn2);
                if (MODE == 1) s_thr[j] = a.thr[q];
                if (MODE == 0) s_colmin[j] = 0xFFFFFFFFu;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const uint32_t r = p * kGM + quarter * 32 + lane;
            const bool rvalid = r < a.n;
            const float xn2 = rvalid ? a.xn2[r] : 0.f;
            const float xnorm = sqrtf(xn2);
            mbar_wait(tfull + buf, bphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < kGN / 32; c++) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kGN + c * 32, v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int col = c * 32 + j;
                    const float s = __uint_as_float(v[j]);
                    float key, err;
                    if (METRIC == 1) {
                        key = -s;
                        err = s_cq[col] * xnorm;
                    } else {
                        const float sum = xn2 + s_qn2[col];
                        key = fmaf(-2.f, s, sum);
                        err = fmaf(2.f * s_cq[col], xnorm, kErrDelta * sum);
                    }
                    if (MODE == 0) {
                        const uint32_t o = rvalid ? f2ord(key + err) : 0xFFFFFFFFu;
                        const uint32_t m = __reduce_min_sync(0xffffffffu, o);
                        if (lane == 0) atomicMin(&s_colmin[col], m);
                    } else {
                        const uint32_t q = t * kGN + col;
                        const bool hit = rvalid && q < a.nq && (key - err) <= s_thr[col];
                        const uint32_t m = __ballot_sync(0xffffffffu, hit);
                        if (m) {
                            uint32_t base = 0;
                            if (lane == 0) base = atomicAdd(a.cand_cnt + q, (uint32_t)__popc(m));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (hit) {
                                const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                                if (pos < a.cap) a.cand[(size_t)q * a.cap + pos] = r;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            if (lane == 0) mbar_arrive(tempty + buf);  // this warp is done with the accumulator
            if (++buf == 2) { buf = 0; bphase ^= 1; }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (MODE == 0)
                for (int j = et; j < kGN; j += 128) a.panelmin[(size_t)pi * a.nq_pad + t * kGN + j] = s_colmin[j];
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// fp32 rows -> bf16 rows padded to Kp (+ zero rows up to rows_pad) and squared norms.  One warp per row.
__global__ void bf_to_bf16_kernel(const float *__restrict__ src, size_t src_stride, uint32_t dim, uint32_t kp,
                                  uint32_t rows, uint32_t rows_pad, __nv_bfloat16 *__restrict__ dst,
                                  float *__restrict__ n2) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    if (r >= rows_pad) return;
    float acc = 0.f;
    for (uint32_t c = lane; c < kp; c += 32) {
        float v = 0.f;
        if (r < rows && c < dim) v = src[(size_t)r * src_stride + c];
        dst[(size_t)r * kp + c] = __float2bfloat16_rn(v);
        acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && n2) n2[r] = acc;
}

// U(q) = k-th smallest of the sampled panel minima (bitwise binary search over the ordered-float domain).
__global__ void bf_kth_kernel(const uint32_t *__restrict__ panelmin, uint32_t sampled, uint32_t nq_pad, uint32_t nq,
                              uint32_t k, float *__restrict__ thr) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_pad) return;
    if (q >= nq || sampled < k) { thr[q] = q >= nq ? -3.402823466e+38f : 3.402823466e+38f; return; }
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t cand = prefix | ((1u << bit) - 1u);  // largest value with this prefix and bit = 0
        uint32_t cnt = 0;
        for (uint32_t g = 0; g < sampled; g++) cnt += panelmin[(size_t)g * nq_pad + q] <= cand ? 1u : 0u;
        if (cnt < k) prefix |= 1u << bit;
    }
    thr[q] = ord2f(prefix);
}

// Exact distance in the reference's SSE order (same arithmetic as bf_scan_kernel): q in shared memory, x a padded row.
template <int METRIC>
__device__ __forceinline__ float exact_dist(const float *q, const float4 *__restrict__ x, uint32_t d4, uint32_t lane_chunks) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, t = 0.f;
    for (uint32_t c = 0; c < d4; c++) {
        const float4 xv = __ldg(x + c);
        const float4 qv = *(const float4 *)(q + 4 * c);
        float m0, m1, m2, m3;
        if (METRIC == 0) {
            const float a0 = __fsub_rn(qv.x, xv.x), a1 = __fsub_rn(qv.y, xv.y), a2 = __fsub_rn(qv.z, xv.z), a3 = __fsub_rn(qv.w, xv.w);
            m0 = __fmul_rn(a0, a0); m1 = __fmul_rn(a1, a1); m2 = __fmul_rn(a2, a2); m3 = __fmul_rn(a3, a3);
        } else {
            m0 = __fmul_rn(qv.x, xv.x); m1 = __fmul_rn(qv.y, xv.y); m2 = __fmul_rn(qv.z, xv.z); m3 = __fmul_rn(qv.w, xv.w);
        }
        if (c < lane_chunks) {
            s0 = __fadd_rn(s0, m0); s1 = __fadd_rn(s1, m1); s2 = __fadd_rn(s2, m2); s3 = __fadd_rn(s3, m3);
        } else {
            t = __fadd_rn(t, m0); t = __fadd_rn(t, m1); t = __fadd_rn(t, m2); t = __fadd_rn(t, m3);
        }
    }
    float r = __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
    r = __fadd_rn(r, t);
    if (METRIC == 1) r = __fsub_rn(1.0f, r);
    return r;
}

// One CTA per query: exact distances of its candidates, k smallest (dist, label) by rank, closest first.
template <int METRIC>
__global__ void __launch_bounds__(256) bf_rerank_kernel(const float4 *__restrict__ X, const uint64_t *__restrict__ labels,
                                                        const float *__restrict__ Q, uint32_t dim, uint32_t d4,
                                                        uint32_t lane_chunks, const uint32_t *__restrict__ cand,
                                                        const uint32_t *__restrict__ cand_cnt, uint32_t cap, uint32_t k,
                                                        uint32_t n, uint64_t *__restrict__ out_l, float *__restrict__ out_d,
                                                        uint32_t *__restrict__ out_c, uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char sm[];
    float *qs = (float *)sm;                        // [d4*4]
    float *cd = qs + d4 * 4;                        // [cap]
    uint64_t *cl = (uint64_t *)(cd + cap + (cap & 1));  // [cap]
    const uint32_t q = blockIdx.x;
    const uint32_t cnt_raw = cand_cnt[q];
    if (cnt_raw > cap) {  // candidate buffer overflowed: this batch is redone by the exact scan
        if (threadIdx.x == 0) atomicAdd(overflow, 1u);
        return;
    }
    const uint32_t cnt = cnt_raw;
    for (uint32_t i = threadIdx.x; i < d4 * 4; i += blockDim.x) qs[i] = i < dim ? Q[(size_t)q * dim + i] : 0.f;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint32_t r = cand[(size_t)q * cap + i];
        cd[i] = exact_dist<METRIC>(qs, X + (size_t)r * d4, d4, lane_chunks);
        cl[i] = labels[r];
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {  // padding for rows the ranks below do not fill
        if (j >= cnt) {
            out_l[(size_t)q * k + j] = 0xFFFFFFFFFFFFFFFFull;
            out_d[(size_t)q * k + j] = __int_as_float(0x7f800000);
        }
    }
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const float di = cd[i];
        const uint64_t li = cl[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            const float dj = cd[j];
            rank += (dj < di || (dj == di && cl[j] < li)) ? 1u : 0u;
        }
        if (rank < k) {
            out_l[(size_t)q * k + rank] = li;
            out_d[(size_t)q * k + rank] = di;
        }
    }
    if (threadIdx.x == 0 && out_c) out_c[q] = min(min(k, n), cnt);
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 tensor [rows][kp], K contiguous, box = 64 x box_rows, 128-byte swizzle.
static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t kp, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return B200HNSW_E_CUDA; }
    cuuint64_t dims[2] = {kp, rows};
    cuuint64_t strides[1] = {kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)kGK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed"); return B200HNSW_E_CUDA; }
    return 0;
}

void BruteTensor::release() {
    cudaFree(xb); cudaFree(xn2); cudaFree(qb); cudaFree(qn2); cudaFree(thr); cudaFree(panelmin); cudaFree(cand);
    cudaFree(cand_cnt); cudaFree(overflow);
    *this = BruteTensor();
}

// (re)build the bf16 copy + norms of rows [first, first+count)
int BruteIndex::tensor_sync_rows(size_t first, size_t count) {
    if (!count) return 0;
    const size_t kp = (host.dim + kGK - 1) / kGK * kGK;
    const size_t rows_pad = (std::max<size_t>(cap, 1) + kGM - 1) / kGM * kGM;
    if (!tz.xb) {
        B200_CUDA_OK(cudaMalloc(&tz.xb, rows_pad * kp * 2));
        B200_CUDA_OK(cudaMemset(tz.xb, 0, rows_pad * kp * 2));
        B200_CUDA_OK(cudaMalloc(&tz.xn2, rows_pad * 4));
        B200_CUDA_OK(cudaMemset(tz.xn2, 0, rows_pad * 4));
        tz.kp = kp;
        tz.rows_pad = rows_pad;
    }
    bf_to_bf16_kernel<<<(unsigned)((count * 32 + 255) / 256), 256>>>((const float *)(dX + first * d4), d4 * 4,
                                                                    (uint32_t)host.dim, (uint32_t)kp, (uint32_t)count,
                                                                    (uint32_t)count, (__nv_bfloat16 *)tz.xb + first * kp,
                                                                    tz.xn2 + first);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// returns 0 on success, 1 when the tensor path does not apply / overflowed (caller falls back to the exact scan)
int BruteIndex::search_tensor(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc,
                              cudaStream_t st) {
    const size_t n = host.cur;
    const size_t panels = (n + kGM - 1) / kGM;
    size_t stride = 4;
    if (const char *e = getenv("B200HNSW_BF_SAMPLE")) stride = std::max(1, atoi(e));
    while (stride > 1 && (panels + stride - 1) / stride < 2 * k) stride /= 2;
    if ((panels + stride - 1) / stride < k) return 1;
    int rc = 0;
    if (!tz.xb) {  // first use: bf16 copy + norms of everything stored so far (kept in sync by upload_rows afterwards)
        rc = tensor_sync_rows(0, n);
        if (rc) return rc;
        B200_CUDA_OK(cudaDeviceSynchronize());
    }
    const size_t kp = tz.kp;
    const size_t nq_pad = (nq + kGN - 1) / kGN * kGN;
    const size_t sampled = (panels + stride - 1) / stride;
    size_t cap_c = 4096;
    if (const char *e = getenv("B200HNSW_BF_CAP")) cap_c = std::max(256, atoi(e));
    if (nq_pad > tz.q_cap || cap_c != tz.cap) {
        cudaFree(tz.qb); cudaFree(tz.qn2); cudaFree(tz.thr); cudaFree(tz.cand); cudaFree(tz.cand_cnt);
        tz.qb = nullptr; tz.qn2 = tz.thr = nullptr; tz.cand = tz.cand_cnt = nullptr; tz.q_cap = 0;
        B200_CUDA_OK(cudaMalloc(&tz.qb, nq_pad * kp * 2));
        B200_CUDA_OK(cudaMalloc(&tz.qn2, nq_pad * 4));
        B200_CUDA_OK(cudaMalloc(&tz.thr, nq_pad * 4));
        B200_CUDA_OK(cudaMalloc(&tz.cand, nq_pad * cap_c * 4));
        B200_CUDA_OK(cudaMalloc(&tz.cand_cnt, nq_pad * 4));
        if (!tz.overflow) B200_CUDA_OK(cudaMalloc(&tz.overflow, 4));
        tz.q_cap = nq_pad;
        tz.cap = cap_c;
    }
    if (sampled * nq_pad > tz.pm_elems) {
        cudaFree(tz.panelmin);
        tz.panelmin = nullptr; tz.pm_elems = 0;
        B200_CUDA_OK(cudaMalloc(&tz.panelmin, sampled * nq_pad * 4));
        tz.pm_elems = sampled * nq_pad;
    }
    bf_to_bf16_kernel<<<(unsigned)((nq_pad * 32 + 255) / 256), 256, 0, st>>>(dQ_, host.dim, (uint32_t)host.dim, (uint32_t)kp,
                                                                           (uint32_t)nq, (uint32_t)nq_pad,
                                                                           (__nv_bfloat16 *)tz.qb, tz.qn2);
    B200_CUDA_OK(cudaMemsetAsync(tz.cand_cnt, 0, nq_pad * 4, st));
    B200_CUDA_OK(cudaMemsetAsync(tz.overflow, 0, 4, st));
    CUtensorMap mA, mB;
    rc = make_map(&mA, tz.xb, tz.rows_pad, kp, kGM);
    if (!rc) rc = make_map(&mB, tz.qb, nq_pad, kp, kGN);
    if (rc) return rc;
    GemmArgs a{};
    a.xn2 = tz.xn2; a.qn2 = tz.qn2; a.thr = tz.thr; a.panelmin = tz.panelmin; a.cand = tz.cand; a.cand_cnt = tz.cand_cnt;
    a.n = (uint32_t)n; a.nq = (uint32_t)nq; a.nq_pad = (uint32_t)nq_pad; a.kchunks = (uint32_t)(kp / kGK);
    a.panels = (uint32_t)panels; a.qtiles = (uint32_t)(nq_pad / kGN); a.cap = (uint32_t)cap_c;
    static bool configured[16] = {};
    if (device < 16 && !configured[device]) {
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_rerank_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_rerank_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[device] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const bool ip = prm.metric == B200HNSW_IP;
    // pass 1: bounds from every stride-th panel
    a.stride = (uint32_t)stride;
    unsigned grid = (unsigned)std::min<size_t>((size_t)sms, sampled * a.qtiles);
    if (ip) bf_gemm_kernel<1, 0><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a);
    else bf_gemm_kernel<0, 0><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a);
    bf_kth_kernel<<<(unsigned)((nq_pad + 127) / 128), 128, 0, st>>>(tz.panelmin, (uint32_t)sampled, (uint32_t)nq_pad,
                                                                   (uint32_t)nq, (uint32_t)k, tz.thr);
    // pass 2: candidates from all panels
    a.stride = 1;
    grid = (unsigned)std::min<size_t>((size_t)sms, panels * a.qtiles);
    if (ip) bf_gemm_kernel<1, 1><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a);
    else bf_gemm_kernel<0, 1><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a);
    // exact re-rank
    const size_t dim = host.dim;
    size_t lane_floats;
    if (dim % 4 == 0) lane_floats = dim;
    else if (dim > 16) lane_floats = dim >> 4 << 4;
    else if (dim > 4) lane_floats = dim >> 2 << 2;
    else lane_floats = 0;
    const size_t rsm = d4 * 16 + (cap_c + (cap_c & 1)) * 4 + cap_c * 8;
    if (ip)
        bf_rerank_kernel<1><<<(unsigned)nq, 256, rsm, st>>>(dX, dLabels, dQ_, (uint32_t)dim, (uint32_t)d4,
                                                           (uint32_t)(lane_floats / 4), tz.cand, tz.cand_cnt,
                                                           (uint32_t)cap_c, (uint32_t)k, (uint32_t)n, dl, dd, dc, tz.overflow);
    else
        bf_rerank_kernel<0><<<(unsigned)nq, 256, rsm, st>>>(dX, dLabels, dQ_, (uint32_t)dim, (uint32_t)d4,
                                                           (uint32_t)(lane_floats / 4), tz.cand, tz.cand_cnt,
                                                           (uint32_t)cap_c, (uint32_t)k, (uint32_t)n, dl, dd, dc, tz.overflow);
    B200_CUDA_OK(cudaGetLastError());
    uint32_t ov = 0;
    B200_CUDA_OK(cudaMemcpyAsync(&ov, tz.overflow, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
    stats.kernel_launches += 5;
    if (getenv("B200HNSW_BF_STATS")) {  // diagnostic: candidates generated per batch (costs a D2H copy)
        std::vector<uint32_t> c(nq);
        B200_CUDA_OK(cudaMemcpy(c.data(), tz.cand_cnt, nq * 4, cudaMemcpyDeviceToHost));
        size_t tot = 0;
        for (uint32_t v : c) tot += v;
        tz.last_candidates = tot;
    }
    return ov ? 1 : 0;
}

}  // namespace b200

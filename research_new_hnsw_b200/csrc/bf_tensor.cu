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
//   E(q,x)    = c |q| |x| (+ delta (|x|^2+|q|^2) for L2),  c = 2^-7 + 2^-12  >= |key - true key|
//               (bf16 keeps 8 significant bits: rounding to nearest is 2^-8 relative per operand, so a product is off
//               by at most (2^-7 + 2^-16)|q_i x_i| and the sum by that times |q||x| (Cauchy-Schwarz); the 2^-12 covers
//               the fp32 accumulation of up to 1024 terms)
//   pass 1    every `stride`-th 128-row panel: panelmin[panel][q] = min_x key + E.  Each panel holds a row whose TRUE
//             key is <= its panelmin, so U(q) = k-th smallest panelmin is an upper bound of the true k-th best key.
//   pass 2    all panels: emit (row x, key - E) for query q iff key - E <= U(q)  (a superset of the true top-k;
//             ballot-compacted, one atomicAdd per warp and query column).
//   re-rank   U2(q) = k-th smallest key + E over the candidates (every candidate's true key is below its key + E, so
//             U2 bounds the true k-th best key as well, and far more tightly than U: it sees all rows, not a sample);
//             survivors = candidates with key - E <= U2 (k plus the rows inside the error band, ~1.4 k); only those
//             get exact distances, then the k smallest (dist, label), closest first.
//
// Roles in the 320-thread CTA (Blackwell playbook): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// MMA issuer, warps 2-9 = epilogue (tcgen05.ld of their 32-lane quarter = warp % 4, half of the columns each).  Pipelines: 4-stage smem ring (full/empty
// mbarriers, tcgen05.commit frees a stage), 2 accumulator buffers of 256 TMEM columns (tmem_full / tmem_empty).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "bruteforce.cuh"
#include "mbarrier.cuh"

namespace b200 {

constexpr int kGM = 128;       // rows per tile (UMMA M)
constexpr int kGN = 256;       // queries per tile (UMMA N)
constexpr int kGK = 64;        // K elements per stage (128 bytes of bf16 = one swizzle atom row)
constexpr int kGStages = 4;
// Epilogue warps: 8 = two per TMEM lane quarter, each taking half of the tile's 256 query columns.  With one warp per
// quarter the epilogue of a tile (8 chunks of 32 columns per warp) took longer than the tile's MMAs and the tensor pipe
// waited for free accumulators: measured at C4 with the epilogue switched off 9.7 ms (9.0 ms as CTA pairs), with half of
// it 9.9 ms, with all of it on four warps 12.0 ms (gpurun_out/s2_bf_epi.log).
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kEpiCols = kGN / (kEpiWarps / 4);   // query columns per epilogue warp
constexpr int kGThreads = 64 + kEpiThreads;
constexpr uint32_t kStageA = kGM * kGK * 2;   // 16 KB
constexpr uint32_t kStageB = kGN * kGK * 2;   // 32 KB
constexpr uint32_t kStageBytes = kStageA + kStageB;
// CTA-pair variant (tcgen05 cta_group::2): the pair multiplies 256 rows x 256 queries per instruction; each CTA stages its
// own 128 rows of A and HALF of the query tile (128 queries), the tensor cores read the other half from the partner's
// shared memory -- a third fewer bytes per tile from L2 into each SM (the single-CTA mainloop pulls 14.6-16.8 TB/s
// through the L2->SM path, profiles/r02_bf_c4_kernels_ncu_full.txt), and the smaller stage buys a 6-deep ring.
constexpr int kGStages2 = 6;
constexpr uint32_t kStageB2 = (kGN / 2) * kGK * 2;   // 16 KB
constexpr uint32_t kStageBytes2 = kStageA + kStageB2;
constexpr int kQCap = 256;     // entries of one epilogue warp's hit queue (pass 2)
constexpr uint32_t kGemmSmem = kGStages * kStageBytes + 1024 /*align*/ + 8192 /*barriers + per-tile tables*/ +
                               kEpiWarps * 2 * kQCap * 4 /*hit queues*/;
static_assert(kGStages2 * kStageBytes2 == kGStages * kStageBytes, "both variants share one shared-memory carve-up");
constexpr float kErrC = 0.0078125f + 0.000244140625f;   // 2^-7 + 2^-12
constexpr float kErrDelta = 4e-6f;
#ifndef B200_BF_DEFAULT_CG
#define B200_BF_DEFAULT_CG 1
#endif
constexpr int kDefaultCG = B200_BF_DEFAULT_CG;

struct GemmArgs {
    const float *xn2;       // [rows] squared norms (fp32)
    const float *tabB;      // [nq_pad] per-query error slope  (bf_tables_kernel)
    const float *tabT;      // [nq_pad] per-query additive term / threshold
    uint32_t *panelmin;     // [sampled panels][nq_pad] ordered floats (pass 1)
    uint2 *cand;            // [nq][cap] (row id, bits of key - E without the per-query constant) (pass 2)
    uint32_t *cand_cnt;     // [nq]
    uint32_t n, nq, nq_pad, kchunks, panels, stride, qtiles, cap;
    const uint8_t *mask;    // [n] row filter of the call (bruteforce.h:114,121) or null
    uint32_t dbg_chunks;    // chunks of 32 columns every epilogue warp evaluates (kEpiCols / 32; fewer: timing experiments)
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// CTA-pair forms: the barrier operand is a shared::cluster address (the leader CTA's barrier for loads)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {  // same offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {  // arrives on the barrier at this offset in BOTH CTAs
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
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
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGN >> 3) << 17) | ((uint32_t)((2 * kGM) >> 4) << 24);

// Hit queue of one epilogue warp: entries (column of the tile << 8 | row of the tile, value), all of the CURRENT tile.
// 32 entries per step: their atomicAdds on the per-query counters are in flight together.
__device__ __forceinline__ void flush_queue(const GemmArgs &a, const uint32_t *qmeta, const float *qval, uint32_t qn,
                                            uint32_t p, uint32_t t, int lane) {
    __syncwarp();
    for (uint32_t e = lane; e < qn; e += 32) {
        const uint32_t m = qmeta[e];
        const uint32_t q = t * kGN + (m >> 8), r = p * kGM + (m & 0xFFu);
        const uint32_t pos = atomicAdd(a.cand_cnt + q, 1u);
        if (pos < a.cap) a.cand[(size_t)q * a.cap + pos] = make_uint2(r, __float_as_uint(qval[e]));
    }
    __syncwarp();
}

// CG = 1: one CTA per 128-row panel.  CG = 2: launched in clusters of two CTAs (one TPC); the pair works on two
// consecutive (sampled) panels and one query tile per item, rank 0 issues the cta_group::2 MMAs for both.
template <int METRIC, int MODE, int CG = 1>
__global__ void __launch_bounds__(kGThreads, 1) bf_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const GemmArgs a) {
    constexpr int kNS = CG == 2 ? kGStages2 : kGStages;
    constexpr uint32_t kSB = CG == 2 ? kStageBytes2 : kStageBytes;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B) as an OFFSET into the shared array: casting through an integer would lose the
    // address space and turn every table / queue access of the epilogue into a generic load
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *tail = smem + kGStages * kStageBytes;
    uint64_t *full = (uint64_t *)tail;            // [kNS] (max 6)
    uint64_t *empty = full + kGStages2;           // [kNS]
    uint64_t *tfull = empty + kGStages2;          // [2]
    uint64_t *tempty = tfull + 2;                 // [2]
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sampled = (a.panels + a.stride - 1) / a.stride;
    // item = (panel or panel pair, query tile); the CTAs of a pair walk the same items
    const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
    const uint32_t items = (CG == 2 ? (sampled + 1) / 2 : sampled) * a.qtiles;
    const uint32_t it0 = CG == 2 ? blockIdx.x / 2 : blockIdx.x, it_step = CG == 2 ? gridDim.x / 2 : gridDim.x;

    if (threadIdx.x == 0) {
        // CG = 2: `full` and `tempty` are only used in the leader (loads of both CTAs complete on the leader's barrier,
        // the epilogue warps of both CTAs release the accumulators there); `empty` / `tfull` exist in both and receive
        // the leader's multicast commits
        for (int s = 0; s < kNS; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull + b, 1); mbar_init(tempty + b, kEpiWarps * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 256-column accumulators)
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // the partner's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t it = it0; it < items; it += it_step) {
                const uint32_t p = ((it / a.qtiles) * CG + rank) * a.stride, t = it % a.qtiles;
                for (uint32_t kc = 0; kc < a.kchunks; kc++) {
                    mbar_wait(empty + stage, phase ^ 1);  // (CG = 2: the pair's MMAs on this stage have retired)
                    unsigned char *sa = smem + stage * kSB;
                    if (CG == 2) {
                        // both CTAs' boxes complete on the LEADER's barrier: 2 x (A half + B half) bytes per stage.  A panel
                        // beyond the last one (odd count) is an out-of-bounds box: zero fill, full byte count.
                        if (rank == 0) mbar_expect_tx(full + stage, 2 * kStageBytes2);
                        const uint32_t fb = map_to_cta(full + stage, 0);
                        tma_load_2d_pair(sa, &tmA, fb, (int)(kc * kGK), (int)(p * kGM));
                        tma_load_2d_pair(sa + kStageA, &tmB, fb, (int)(kc * kGK), (int)(t * kGN + rank * (kGN / 2)));
                    } else {
                        mbar_expect_tx(full + stage, kStageBytes);
                        tma_load_2d(sa, &tmA, full + stage, (int)(kc * kGK), (int)(p * kGM));
                        tma_load_2d(sa + kStageA, &tmB, full + stage, (int)(kc * kGK), (int)(t * kGN));
                    }
                    if (++stage == kNS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && rank == 0) {
            uint32_t stage = 0, phase = 0, buf = 0, bphase = 0;
            for (uint32_t it = it0; it < items; it += it_step) {
                mbar_wait(tempty + buf, bphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * kGN;
                for (uint32_t kc = 0; kc < a.kchunks; kc++) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kSB);
                    const uint64_t da = umma_desc(sa), db = umma_desc(sa + kStageA);
#pragma unroll
                    for (uint32_t k4 = 0; k4 < kGK / 16; k4++) {  // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
                        if (CG == 2) tc_mma_bf16_pair(d_tmem, da + 2 * k4, db + 2 * k4, kIdesc2, (kc | k4) != 0);
                        else tc_mma_bf16(d_tmem, da + 2 * k4, db + 2 * k4, kIdesc, (kc | k4) != 0);
                    }
                    // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                    if (CG == 2) tc_commit_pair(empty + stage); else tc_commit(empty + stage);
                    if (++stage == kNS) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) tc_commit_pair(tfull + buf); else tc_commit(tfull + buf);  // accumulator complete
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
        // Per element (row r = TMEM lane, query column j):  w = fma(+-B_j, |x_r|, alpha * S)  with alpha = -1 (IP) / -2 (L2),
        //   pass 1 (MODE 0): min over the 32 rows of a warp of  w + Ar  (+ T_j once per column), Ar = (1+delta)|x|^2 (L2) / 0
        //   pass 2 (MODE 1): hit iff  w <= T_j - Ar,                                   Ar = (1-delta)|x|^2 (L2) / 0
        // B_j, T_j come from bf_tables_kernel.  The 32 columns of a chunk are evaluated branch-free into a bit mask;
        // only chunks that contain a hit (rare) walk their set bits.  Tables of the next item are prefetched into
        // registers while this one is processed and live double-buffered in shared memory.
        const int quarter = warp & 3;
        const int et = threadIdx.x - 64;  // 0..kEpiThreads-1
        const int c_first = ((warp - 2) >> 2) * (kEpiCols / 32);  // this warp's chunks of 32 columns
        constexpr int kTabPer = kGN / kEpiThreads;  // table entries staged per thread
        float *s_B = (float *)(tail + 256);          // [2][kGN]
        float *s_T = s_B + 2 * kGN;                  // [2][kGN]
        uint32_t *s_min = (uint32_t *)(s_T + 2 * kGN);  // [2][kGN] (pass 1)
        uint32_t *qmeta = s_min + 2 * kGN + (warp - 2) * 2 * kQCap;  // this warp's hit queue (pass 2): [kQCap] meta
        float *qval = (float *)(qmeta + kQCap);                       // [kQCap] key - E (without the per-query constant)
        uint32_t qn = 0;
        uint32_t pend_q = 0, pend_r = 0, pend_pos = 0;  // this lane's queue entry whose counter atomic is in flight
        float pend_v = 0.f;
        bool pend = false;
        constexpr float alpha = METRIC == 1 ? -1.f : -2.f;
        uint32_t buf = 0, bphase = 0, tb = 0;
        float pB[kTabPer], pT[kTabPer];
        uint32_t it = it0;
        const uint32_t tempty_leader = CG == 2 ? map_to_cta(tempty, 0) : 0u;  // tempty[b] at + 8 * b
        if (it < items) {
            const uint32_t t = it % a.qtiles;
#pragma unroll
            for (int e = 0; e < kTabPer; e++) {
                const uint32_t q = t * kGN + et + e * kEpiThreads;
                s_B[et + e * kEpiThreads] = a.tabB[q];
                s_T[et + e * kEpiThreads] = a.tabT[q];
            }
        }
#pragma unroll
        for (int e = 0; e < kTabPer; e++) s_min[et + e * kEpiThreads] = s_min[kGN + et + e * kEpiThreads] = 0xFFFFFFFFu;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        uint32_t prev_pi = 0, prev_t = 0;
        bool have_prev = false;
        for (; it < items; it += it_step) {
            const uint32_t pi = (it / a.qtiles) * CG + rank, p = pi * a.stride, t = it % a.qtiles;
            const uint32_t nit = it + it_step;
            if (nit < items) {  // prefetch the next item's tables
                const uint32_t nt = nit % a.qtiles;
#pragma unroll
                for (int e = 0; e < kTabPer; e++) {
                    const uint32_t q = nt * kGN + et + e * kEpiThreads;
                    pB[e] = a.tabB[q];
                    pT[e] = a.tabT[q];
                }
            }
            if (MODE == 0 && have_prev) {  // flush the previous item's column minima (other table buffer)
#pragma unroll
                for (int e = 0; e < kTabPer; e++) {
                    const int j = et + e * kEpiThreads;
                    if (prev_pi < sampled) a.panelmin[(size_t)prev_pi * a.nq_pad + prev_t * kGN + j] = s_min[(tb ^ 1) * kGN + j];
                    s_min[(tb ^ 1) * kGN + j] = 0xFFFFFFFFu;
                }
            }
            const float *B = s_B + tb * kGN, *T = s_T + tb * kGN;
            const uint32_t r = p * kGM + quarter * 32 + lane;
            const bool rvalid = pi < sampled && r < a.n && (!a.mask || a.mask[r]);
            const float xn2 = rvalid ? a.xn2[r] : 0.f;
            const float xnorm = sqrtf(xn2);
            const float Ar = METRIC == 1 ? 0.f : (MODE == 0 ? (1.f + kErrDelta) : (1.f - kErrDelta)) * xn2;
            const float bx = MODE == 0 ? xnorm : -xnorm;
            mbar_wait(tfull + buf, bphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = c_first; c < c_first + (int)a.dbg_chunks; c++) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kGN + c * 32, v);
                float Bc[32], Tc[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++) {
                    const float4 b4 = *(const float4 *)(B + c * 32 + j4 * 4);
                    Bc[j4 * 4] = b4.x; Bc[j4 * 4 + 1] = b4.y; Bc[j4 * 4 + 2] = b4.z; Bc[j4 * 4 + 3] = b4.w;
                    if (MODE == 1) {
                        const float4 t4 = *(const float4 *)(T + c * 32 + j4 * 4);
                        Tc[j4 * 4] = t4.x; Tc[j4 * 4 + 1] = t4.y; Tc[j4 * 4 + 2] = t4.z; Tc[j4 * 4 + 3] = t4.w;
                    }
                }
                tc_wait_ld();
                if (MODE == 0) {
                    uint32_t mine = 0xFFFFFFFFu;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float w = fmaf(Bc[j], bx, alpha * __uint_as_float(v[j])) + Ar;
                        const uint32_t o = rvalid ? f2ord(w) : 0xFFFFFFFFu;
                        const uint32_t m = __reduce_min_sync(0xffffffffu, o);
                        mine = lane == j ? m : mine;
                    }
                    // lane j now holds the warp's minimum of column c*32+j; T_j is added once per column
                    if (mine != 0xFFFFFFFFu)
                        atomicMin(&s_min[tb * kGN + c * 32 + lane], f2ord(ord2f(mine) + T[c * 32 + lane]));
                } else {
                    uint32_t hits = 0;
                    const float rhs_r = -Ar;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float w = fmaf(Bc[j], bx, alpha * __uint_as_float(v[j]));
                        hits |= (w <= Tc[j] + rhs_r) ? (1u << j) : 0u;
                    }
                    if (!rvalid) hits = 0;
                    if (__any_sync(0xffffffffu, hits != 0)) {
                        // Hits go to this warp's shared-memory queue (no global round trip on the epilogue's critical
                        // path: one atomicAdd per hit column cost ~1 us each, more than the tile's whole mainloop);
                        // the queue is flushed 32 entries at a time, their atomics in flight together.
                        uint32_t mine = (uint32_t)__popc(hits), incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += up;
                        }
                        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                        if (qn + total > (uint32_t)kQCap) {
                            flush_queue(a, qmeta, qval, qn, p, t, lane);
                            qn = 0;
                        }
                        if (total <= (uint32_t)kQCap) {
                            uint32_t pos = qn + incl - mine;
                            uint32_t hb = hits;
                            while (hb) {
                                const int j = __ffs(hb) - 1;
                                hb &= hb - 1;
                                uint32_t vj = 0;  // v[] lives in registers: select, do not index
#pragma unroll
                                for (int jj = 0; jj < 32; jj++) vj = jj == j ? v[jj] : vj;
                                const float wj = fmaf(B[c * 32 + j], bx, alpha * __uint_as_float(vj));
                                qmeta[pos] = ((uint32_t)(c * 32 + j) << 8) | (uint32_t)(quarter * 32 + lane);
                                qval[pos] = wj + Ar;
                                pos++;
                            }
                            qn += total;
                        } else {  // a chunk with more hits than the queue holds (degenerate data): direct path
                            uint32_t any = __reduce_or_sync(0xffffffffu, hits);
                            while (any) {
                                const int j = __ffs(any) - 1;
                                any &= any - 1;
                                const uint32_t q = t * kGN + c * 32 + j;
                                const bool hit = (hits >> j) & 1u;
                                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                                uint32_t base = 0;
                                if (lane == 0) base = atomicAdd(a.cand_cnt + q, (uint32_t)__popc(m));
                                base = __shfl_sync(0xffffffffu, base, 0);
                                if (hit) {
                                    const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                                    if (pos < a.cap) {
                                        uint32_t vj = 0;
#pragma unroll
                                        for (int jj = 0; jj < 32; jj++) vj = jj == j ? v[jj] : vj;
                                        const float wj = fmaf(B[c * 32 + j], bx, alpha * __uint_as_float(vj));
                                        a.cand[(size_t)q * a.cap + pos] = make_uint2(r, __float_as_uint(wj + Ar));
                                    }
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            if (lane == 0) {  // this warp is done with the accumulator
                if (CG == 2) mbar_arrive_cluster(tempty_leader + 8 * buf); else mbar_arrive(tempty + buf);
            }
            if (++buf == 2) { buf = 0; bphase ^= 1; }
            if (MODE == 1) {
                // The queue's entries are relative to this tile, so it is emptied before the next one -- but nobody has to
                // wait for it: the counter atomics of the first 32 entries are ISSUED now, their candidate stores happen
                // at the end of the next tile (the round trip hides behind that tile's chunks).  More than 32 hits per
                // warp and tile are rare and go the synchronous way.
                if (pend && pend_pos < a.cap) a.cand[(size_t)pend_q * a.cap + pend_pos] = make_uint2(pend_r, __float_as_uint(pend_v));
                pend = false;
                if (qn) {
                    __syncwarp();
                    if ((uint32_t)lane < qn) {
                        const uint32_t m = qmeta[lane];
                        pend_q = t * kGN + (m >> 8);
                        pend_r = p * kGM + (m & 0xFFu);
                        pend_v = qval[lane];
                        pend_pos = atomicAdd(a.cand_cnt + pend_q, 1u);
                        pend = true;
                    }
                    if (qn > 32) flush_queue(a, qmeta + 32, qval + 32, qn - 32, p, t, lane);
                    __syncwarp();
                    qn = 0;
                }
            }
            if (nit < items) {
#pragma unroll
                for (int e = 0; e < kTabPer; e++) {
                    s_B[(tb ^ 1) * kGN + et + e * kEpiThreads] = pB[e];
                    s_T[(tb ^ 1) * kGN + et + e * kEpiThreads] = pT[e];
                }
            }
            prev_pi = pi; prev_t = t; have_prev = true;
            tb ^= 1;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        }
        if (MODE == 0 && have_prev && prev_pi < sampled) {
#pragma unroll
            for (int e = 0; e < kTabPer; e++) {
                const int j = et + e * kEpiThreads;
                a.panelmin[(size_t)prev_pi * a.nq_pad + prev_t * kGN + j] = s_min[(tb ^ 1) * kGN + j];
            }
        }
        if (MODE == 1 && pend && pend_pos < a.cap)
            a.cand[(size_t)pend_q * a.cap + pend_pos] = make_uint2(pend_r, __float_as_uint(pend_v));
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // nobody leaves while the partner may still read its shared memory or signal it
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
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

// Per-query epilogue tables.  pass 1: B = c|q| (IP) / 2c|q| (L2), T = 0 / (1+delta)|q|^2.
//                            pass 2: B as above,             T = U(q) / U(q) - (1-delta)|q|^2.
__global__ void bf_tables_kernel(const float *__restrict__ qn2, const float *__restrict__ thr, uint32_t nq_pad, int metric,
                                 int mode, float *__restrict__ tabB, float *__restrict__ tabT) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_pad) return;
    const float n2 = qn2[q], cq = kErrC * sqrtf(n2);
    if (metric == 1) {
        tabB[q] = cq;
        tabT[q] = mode == 0 ? 0.f : thr[q];
    } else {
        tabB[q] = 2.f * cq;
        tabT[q] = mode == 0 ? (1.f + kErrDelta) * n2 : thr[q] - (1.f - kErrDelta) * n2;
    }
}

// U(q) = k-th smallest of the sampled panel minima (bitwise binary search over the ordered-float domain).
// Block = 32 queries x 8 slices of the panel range; loads are coalesced over queries.  Reads the minima 32 times from
// L2: only used when the staged version below does not fit in shared memory.
__global__ void __launch_bounds__(256) bf_kth_kernel(const uint32_t *__restrict__ panelmin, uint32_t sampled,
                                                     uint32_t nq_pad, uint32_t nq, uint32_t k, float *__restrict__ thr) {
    __shared__ uint32_t part[8][32];
    const uint32_t x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * 32 + x;  // nq_pad is a multiple of 256, so q < nq_pad
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t cand = prefix | ((1u << bit) - 1u);  // largest value with this prefix and bit = 0
        uint32_t cnt = 0;
        for (uint32_t g = y; g < sampled; g += 8) cnt += panelmin[(size_t)g * nq_pad + q] <= cand ? 1u : 0u;
        part[y][x] = cnt;
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) tot += part[i][x];
        __syncthreads();
        if (tot < k) prefix |= 1u << bit;
    }
    if (y == 0) thr[q] = q >= nq ? -3.402823466e+38f : (sampled < k ? 3.402823466e+38f : ord2f(prefix));
}

// Same result with the minima of kKthQ queries staged ONCE in shared memory ([sampled][kKthQ] words, one 32-byte sector
// per panel row), the 32 counting rounds then run out of shared memory: 1.1 ms -> tens of microseconds at C4.
constexpr int kKthQ = 8;
__global__ void __launch_bounds__(256) bf_kth_smem_kernel(const uint32_t *__restrict__ panelmin, uint32_t sampled,
                                                          uint32_t nq_pad, uint32_t nq, uint32_t k, float *__restrict__ thr) {
    extern __shared__ uint32_t s_pm[];  // [sampled][kKthQ]
    __shared__ uint32_t s_tot[2][kKthQ];
    const uint32_t qi = threadIdx.x % kKthQ, slice = threadIdx.x / kKthQ;  // 32 slices of the panel range
    const uint32_t q0 = blockIdx.x * kKthQ;
    for (uint32_t i = threadIdx.x; i < sampled * kKthQ; i += 256) s_pm[i] = panelmin[(size_t)(i / kKthQ) * nq_pad + q0 + i % kKthQ];
    if (threadIdx.x < 2 * kKthQ) s_tot[threadIdx.x / kKthQ][threadIdx.x % kKthQ] = 0;
    __syncthreads();
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t cand = prefix | ((1u << bit) - 1u);
        uint32_t cnt = 0;
        for (uint32_t g = slice; g < sampled; g += 256 / kKthQ) cnt += s_pm[g * kKthQ + qi] <= cand ? 1u : 0u;
        const int par = bit & 1;
        if (cnt) atomicAdd(&s_tot[par][qi], cnt);
        __syncthreads();
        if (s_tot[par][qi] < k) prefix |= 1u << bit;
        if (threadIdx.x < kKthQ) s_tot[par ^ 1][threadIdx.x] = 0;  // the other parity is read again two rounds later
        __syncthreads();
    }
    const uint32_t q = q0 + qi;
    if (slice == 0) thr[q] = q >= nq ? -3.402823466e+38f : (sampled < k ? 3.402823466e+38f : ord2f(prefix));
}

// Re-rank, one CTA per query, in three steps:
//  A. candidates (row, key - E) from pass 2; E is recomputed from the two norms, so key + E = (key - E) + 2E;
//  B. bound: U2 = k-th smallest key + E over the candidates (bitwise binary search over ordered floats); survivors are
//     the candidates with key - E <= U2 -- the true top-k plus the rows inside the error band;
//  C. exact distances of the survivors in the reference's SSE order (same arithmetic as bf_scan_kernel: four lane
//     accumulators over the first lane_chunks 128-bit chunks, sequential tail, separate multiply and add), rows staged
//     through shared memory 32 floats at a time so global reads stay coalesced while every thread sums ITS
//     candidate strictly in index order; then the k smallest (dist, label) by rank, closest first.
constexpr int kRrThreads = 256;
constexpr int kRrRows = 64;     // survivor rows staged per round of step C (4 threads per row)
constexpr int kRrChunk4 = 32;   // 128-bit chunks of every row per stage (512 bytes)
// Staged layout of step C: tile[row][lane accumulator l][chunk] -- thread (row, l) finds the 32 floats it sums in one
// stage (elements 4c + l of chunks c0..c0+31) CONTIGUOUS, so it reads them with 8 LDS.128 instead of 32 scalar loads
// with address arithmetic (ncu of the previous layout: 1.23 G warp instructions per 10 k queries, 43 % issue-active --
// the kernel was instruction-bound).  36 floats per (row, l): 16-byte reads of 8 consecutive threads fall into 8
// different bank groups, and the staging stores of a warp (one chunk per lane) into 32 different banks.
constexpr int kRrLane = 36;
constexpr int kRrStride = 4 * kRrLane;  // floats per staged row

template <int METRIC>
__global__ void __launch_bounds__(kRrThreads, 4) bf_rerank_kernel(const float4 *__restrict__ X, const uint64_t *__restrict__ labels,
                                                              const float *__restrict__ xn2, const float *__restrict__ qn2,
                                                              const float *__restrict__ Q, uint32_t dim, uint32_t d4,
                                                              uint32_t lane_chunks, const uint2 *__restrict__ cand,
                                                              const uint32_t *__restrict__ cand_cnt, uint32_t cap, uint32_t scap,
                                                              uint32_t k, uint32_t n, uint64_t *__restrict__ out_l,
                                                              float *__restrict__ out_d, uint32_t *__restrict__ out_c,
                                                              uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char sm[];
    float *qs = (float *)sm;                              // [4][qstr] query, transposed: qs[l * qstr + c] = Q[4c + l]
    const uint32_t qstr = ((d4 + 3) & ~3u) + 4;           // (+4: the four l-rows start in different bank groups)
    // Per-candidate arrays are sized by cap, per-SURVIVOR arrays by scap (<= cap; k plus the rows inside the error band
    // are a small fraction of the candidates): 54 KB instead of 73 KB at cap 2048, i.e. four resident CTAs per SM.
    float *fd = qs + 4 * qstr;                            // [scap] survivor staging list, then their exact distances
    uint32_t *ids = (uint32_t *)(fd + scap);              // [cap] candidate rows, later survivor rows
    uint64_t *cl = (uint64_t *)(ids + cap + (cap & 1));   // [scap] survivor labels
    float *tile = (float *)(cl + scap);                   // [max(kRrRows*kRrStride, 2*cap)]: (key+E, key-E) of every
                                                          // candidate in steps A/B, the staged rows in step C
    __shared__ uint32_t s_part[kRrThreads / 32];
    __shared__ uint32_t s_hist[kRrThreads];
    __shared__ uint32_t s_cnt;
    static_assert(kRrThreads == 256, "step B uses one histogram bin per thread");
    const uint32_t q = blockIdx.x;
    const uint32_t cnt = cand_cnt[q];
    if (cnt > cap) {  // candidate buffer overflowed: this batch is redone by the exact scan
        if (threadIdx.x == 0) atomicAdd(overflow, 1u);
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < d4 * 4; i += kRrThreads) qs[(i & 3) * qstr + (i >> 2)] = i < dim ? Q[(size_t)q * dim + i] : 0.f;
    // ---- A: candidates and their error bounds ----
    const float q2 = qn2[q], qn = sqrtf(q2);
    float2 *he = (float2 *)tile;  // (key + E, key - E); tile[] proper is only written in step C
    for (uint32_t i = tid; i < cnt; i += kRrThreads) {
        const uint2 c = cand[(size_t)q * cap + i];
        ids[i] = c.x;
        const float x2 = xn2[c.x];
        float lo = __uint_as_float(c.y), e;
        if (METRIC == 1) {
            e = kErrC * qn * sqrtf(x2);
        } else {
            lo += (1.f - kErrDelta) * q2;
            e = 2.f * kErrC * qn * sqrtf(x2) + kErrDelta * (x2 + q2);
        }
        // lo was rounded a few times on its way here: widen both sides by a relative 2^-20 of the magnitudes involved
        const float slack = 9.5367431640625e-07f * (fabsf(lo) + e + (METRIC == 0 ? x2 + q2 : 0.f));
        he[i] = make_float2(lo + 2.f * e + slack, lo - slack);
    }
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    // ---- B: U2 = k-th smallest of key + E: radix select over the ordered-float bits, 8 bits per pass (4 passes of one
    // shared-memory histogram each; the bit-by-bit search it replaces took 32 rounds of two block barriers) ----
    uint32_t prefix = 0;
    if (cnt > k) {
        uint32_t mask = 0, kk = k;  // kk-th smallest among the keys that match `prefix` under `mask`
        for (int shift = 24; shift >= 0; shift -= 8) {
            s_hist[tid] = 0;  // kRrThreads == 256 bins
            __syncthreads();
            for (uint32_t i = tid; i < cnt; i += kRrThreads) {
                const uint32_t u = f2ord(he[i].x);
                if ((u & mask) == prefix) atomicAdd(&s_hist[(u >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (warp == 0) {  // lane l scans bins 8l .. 8l+7
                uint32_t h[8], sum = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) { h[b] = s_hist[lane * 8 + b]; sum += h[b]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const uint32_t before = incl - sum;
                if (before < kk && kk <= incl) {  // exactly one lane
                    uint32_t acc = before;
                    int bin = 0;
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        if (acc < kk && kk <= acc + h[b]) { bin = b; break; }
                        acc += h[b];
                    }
                    s_part[0] = (uint32_t)(lane * 8 + bin);
                    s_part[1] = kk - acc;
                }
            }
            __syncthreads();
            prefix |= s_part[0] << shift;
            mask |= 255u << shift;
            kk = s_part[1];
        }
    } else {
        prefix = 0xFFFFFFFFu;
    }
    const float u2 = ord2f(prefix);
    // survivors, compacted in place (fd[] is reused as a staging list first)
    uint32_t *surv = (uint32_t *)fd;
    for (uint32_t b0 = 0; b0 < cnt; b0 += kRrThreads) {
        const uint32_t i = b0 + tid;
        const bool keep = i < cnt && (cnt <= k || he[i].y <= u2);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_cnt, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < scap) surv[pos] = ids[i];
    }
    __syncthreads();
    const uint32_t ns = s_cnt;
    if (ns > scap) {  // more rows inside the error band than the survivor arrays hold: handled like a full candidate buffer
        if (tid == 0) atomicAdd(overflow, 1u);
        return;
    }
    for (uint32_t i = tid; i < ns; i += kRrThreads) ids[i] = surv[i];
    __syncthreads();
    // ---- C: exact distances of the survivors in reference order ----
    // kRrRows rows per round, 128 floats of every row per stage (coalesced 512-byte reads, 8 independent 128-bit loads per
    // thread in flight); thread (row, l) owns the reference's lane accumulator l of its row -- elements 4j + l in index
    // order -- and thread (row, 0) the sequential tail; ((s0+s1)+s2)+s3 + tail is formed over the four lanes at the end.
    const uint32_t rl = (uint32_t)tid >> 2, al = (uint32_t)tid & 3;
    constexpr int kLd = kRrRows * kRrChunk4 / kRrThreads;  // 128-bit loads per thread and stage
    // stage (base, c0) -> registers; the loads of the NEXT stage are issued before the current one is consumed, so their
    // latency hides behind the arithmetic (ncu of the single-buffered version: 25 % of the samples on the staging stores
    // waiting for their loads, and prefetching all survivor rows to L2 up front thrashed it: 1.65x the DRAM bytes)
    float4 v[kLd];
    auto issue = [&](uint32_t base, uint32_t c0) {
        const uint32_t nbb = min((uint32_t)kRrRows, ns - base);
#pragma unroll
        for (int i = 0; i < kLd; i++) {
            const uint32_t f = tid + kRrThreads * i;  // float4 slot: row f / 32, part f % 32
            const uint32_t cj = f / kRrChunk4, part = f % kRrChunk4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cj < nbb && c0 + part < d4) v[i] = __ldg(X + (size_t)ids[base + cj] * d4 + c0 + part);
        }
    };
    if (ns) issue(0, 0);
    for (uint32_t base = 0; base < ns; base += kRrRows) {
        const uint32_t nb = min((uint32_t)kRrRows, ns - base);
        float acc = 0.f, t = 0.f;
        for (uint32_t c0 = 0; c0 < d4; c0 += kRrChunk4) {
            __syncthreads();  // the previous stage has been consumed
#pragma unroll
            for (int i = 0; i < kLd; i++) {
                const uint32_t f = tid + kRrThreads * i;
                float *dst = tile + (f / kRrChunk4) * kRrStride + (f % kRrChunk4);
                dst[0] = v[i].x; dst[kRrLane] = v[i].y; dst[2 * kRrLane] = v[i].z; dst[3 * kRrLane] = v[i].w;
            }
            __syncthreads();
            if (c0 + kRrChunk4 < d4) issue(base, c0 + kRrChunk4);
            else if (base + kRrRows < ns) issue(base + kRrRows, 0);
            if (rl < nb) {
                const float *row = tile + rl * kRrStride;
                const uint32_t pend = min((uint32_t)kRrChunk4, d4 - c0);
                if (c0 + kRrChunk4 <= lane_chunks) {
                    // the whole stage belongs to the lane accumulators: 32 terms, strictly in index order
                    const float4 *x4 = (const float4 *)(row + al * kRrLane);
                    const float4 *q4 = (const float4 *)(qs + al * qstr + c0);
#pragma unroll
                    for (int j = 0; j < kRrChunk4 / 4; j++) {
                        const float4 xv = x4[j], qv = q4[j];
                        if (METRIC == 0) {
                            const float a0 = __fsub_rn(qv.x, xv.x), a1 = __fsub_rn(qv.y, xv.y), a2 = __fsub_rn(qv.z, xv.z),
                                        a3 = __fsub_rn(qv.w, xv.w);
                            acc = __fadd_rn(acc, __fmul_rn(a0, a0));
                            acc = __fadd_rn(acc, __fmul_rn(a1, a1));
                            acc = __fadd_rn(acc, __fmul_rn(a2, a2));
                            acc = __fadd_rn(acc, __fmul_rn(a3, a3));
                        } else {
                            acc = __fadd_rn(acc, __fmul_rn(qv.x, xv.x));
                            acc = __fadd_rn(acc, __fmul_rn(qv.y, xv.y));
                            acc = __fadd_rn(acc, __fmul_rn(qv.z, xv.z));
                            acc = __fadd_rn(acc, __fmul_rn(qv.w, xv.w));
                        }
                    }
                } else
                for (uint32_t part = 0; part < pend; part++) {
                    const uint32_t c = c0 + part;
                    if (c < lane_chunks) {
                        const float x = row[al * kRrLane + part], qv = qs[al * qstr + c];
                        float m;
                        if (METRIC == 0) {
                            const float a0 = __fsub_rn(qv, x);
                            m = __fmul_rn(a0, a0);
                        } else {
                            m = __fmul_rn(qv, x);
                        }
                        acc = __fadd_rn(acc, m);
                    } else if (al == 0) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float x = row[e * kRrLane + part], qv = qs[e * qstr + c];
                            float m;
                            if (METRIC == 0) {
                                const float a0 = __fsub_rn(qv, x);
                                m = __fmul_rn(a0, a0);
                            } else {
                                m = __fmul_rn(qv, x);
                            }
                            t = __fadd_rn(t, m);
                        }
                    }
                }
            }
        }
        // the four lane accumulators of a row sit in four consecutive lanes
        const float s1 = __shfl_down_sync(0xffffffffu, acc, 1), s2 = __shfl_down_sync(0xffffffffu, acc, 2),
                    s3 = __shfl_down_sync(0xffffffffu, acc, 3);
        if (rl < nb && al == 0) {
            float r = __fadd_rn(__fadd_rn(__fadd_rn(acc, s1), s2), s3);
            r = __fadd_rn(r, t);
            if (METRIC == 1) r = __fsub_rn(1.0f, r);
            fd[base + rl] = r;
            cl[base + rl] = labels[ids[base + rl]];
        }
        __syncthreads();
    }
    __syncthreads();
    for (uint32_t j = tid; j < k; j += kRrThreads) {  // padding for rows the ranks below do not fill
        if (j >= ns) {
            out_l[(size_t)q * k + j] = 0xFFFFFFFFFFFFFFFFull;
            out_d[(size_t)q * k + j] = __int_as_float(0x7f800000);
        }
    }
    for (uint32_t i = tid; i < ns; i += kRrThreads) {
        const float di = fd[i];
        const uint64_t li = cl[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < ns; j++) {
            const float dj = fd[j];
            rank += (dj < di || (dj == di && cl[j] < li)) ? 1u : 0u;
        }
        if (rank < k) {
            out_l[(size_t)q * k + rank] = li;
            out_d[(size_t)q * k + rank] = di;
        }
    }
    if (tid == 0 && out_c) out_c[q] = min(min(k, n), ns);  // n = rows that pass the call's filter
}

template <int METRIC>
static void launch_rerank(unsigned grid, size_t smem, cudaStream_t st, const float4 *X, const uint64_t *labels,
                          const float *xn2, const float *qn2, const float *Q, uint32_t dim, uint32_t d4, uint32_t lane_chunks,
                          const uint2 *cand, const uint32_t *cand_cnt, uint32_t cap, uint32_t scap, uint32_t k, uint32_t n,
                          uint64_t *out_l, float *out_d, uint32_t *out_c, uint32_t *overflow) {
    static bool cfg[16] = {};  // function attributes are per device
    int dv = 0;
    cudaGetDevice(&dv);
    if (dv >= 16 || !cfg[dv]) {
        cudaFuncSetAttribute(bf_rerank_kernel<METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (dv < 16) cfg[dv] = true;
    }
    bf_rerank_kernel<METRIC><<<grid, kRrThreads, smem, st>>>(X, labels, xn2, qn2, Q, dim, d4, lane_chunks, cand, cand_cnt, cap,
                                                            scap, k, n, out_l, out_d, out_c, overflow);
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
    cudaFree(xb); cudaFree(xn2); cudaFree(qb); cudaFree(qn2); cudaFree(thr); cudaFree(tabB); cudaFree(tabT);
    cudaFree(panelmin); cudaFree(cand);
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
    // Sampled bound pass: every stride-th panel.  The bound is the k-th smallest of the sampled panel minima, i.e. about the
    // k * stride-th smallest key overall: a sparser sample costs candidates (680 -> 1350 -> 2750 per query at stride 4 /
    // 8 / 16, C4) but saves 1/stride of a full GEMM.  Measured at C4 (gpurun_out/s2_bf_sample.log): 17.8 / 16.8 / 17.2 ms.
    size_t stride = 8;
    if (const char *e = getenv("B200HNSW_BF_SAMPLE")) stride = std::max(1, atoi(e));
    while (stride > 1 && (panels + stride - 1) / stride < 2 * k) stride /= 2;
    if ((panels + stride - 1) / stride < k) return 1;
    if (d4 > 256) return 1;  // dim > 1024: exact scan
    int rc = 0;
    if (!tz.xb) {  // first use: bf16 copy + norms of everything stored so far (kept in sync by upload_rows afterwards)
        rc = tensor_sync_rows(0, n);
        if (rc) return rc;
        B200_CUDA_OK(cudaDeviceSynchronize());
    }
    const size_t kp = tz.kp;
    const size_t nq_pad = (nq + kGN - 1) / kGN * kGN;
    const size_t sampled = (panels + stride - 1) / stride;
    // candidate slots per query: k/f + the error band is the expectation; an overflowing batch is retried with 4x
    size_t cap_c = std::max<size_t>(2048, (stride >= 8 ? 20 : 12) * k);
    if (const char *e = getenv("B200HNSW_BF_CAP")) cap_c = std::max(256, atoi(e));
    cap_c = (cap_c + 3) & ~(size_t)3;
    cap_c = std::max(cap_c, tz.cap_floor);
    cap_c = (cap_c + 3) / 4 * 4;  // the re-rank kernel's shared arrays stay 16-byte aligned
    if (nq_pad > tz.q_cap || cap_c != tz.cap) {
        cudaFree(tz.qb); cudaFree(tz.qn2); cudaFree(tz.thr); cudaFree(tz.cand); cudaFree(tz.cand_cnt);
        tz.qb = nullptr; tz.qn2 = tz.thr = nullptr; tz.cand = tz.cand_cnt = nullptr; tz.q_cap = 0;
        B200_CUDA_OK(cudaMalloc(&tz.qb, nq_pad * kp * 2));
        B200_CUDA_OK(cudaMalloc(&tz.qn2, nq_pad * 4));
        B200_CUDA_OK(cudaMalloc(&tz.thr, nq_pad * 4));
        cudaFree(tz.tabB); cudaFree(tz.tabT);
        tz.tabB = tz.tabT = nullptr;
        B200_CUDA_OK(cudaMalloc(&tz.tabB, nq_pad * 4));
        B200_CUDA_OK(cudaMalloc(&tz.tabT, nq_pad * 4));
        B200_CUDA_OK(cudaMalloc(&tz.cand, nq_pad * cap_c * 8));
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
    // B200HNSW_BF_PROFILE=1: CUDA events between the stages, printed per call (diagnostic only)
    static const bool prof = getenv("B200HNSW_BF_PROFILE") && atoi(getenv("B200HNSW_BF_PROFILE")) != 0;
    cudaEvent_t pe[6] = {};
    int pn = 0;
    auto mark = [&]() {
        if (prof && pn < 6) { cudaEventCreate(&pe[pn]); cudaEventRecord(pe[pn], st); pn++; }
    };
    mark();
    bf_to_bf16_kernel<<<(unsigned)((nq_pad * 32 + 255) / 256), 256, 0, st>>>(dQ_, host.dim, (uint32_t)host.dim, (uint32_t)kp,
                                                                           (uint32_t)nq, (uint32_t)nq_pad,
                                                                           (__nv_bfloat16 *)tz.qb, tz.qn2);
    B200_CUDA_OK(cudaMemsetAsync(tz.cand_cnt, 0, nq_pad * 4, st));
    B200_CUDA_OK(cudaMemsetAsync(tz.overflow, 0, 4, st));
    // CTA pairs (cta_group::2) unless B200HNSW_BF_CG=1
    const int cg = getenv("B200HNSW_BF_CG") ? atoi(getenv("B200HNSW_BF_CG")) : kDefaultCG;
    const bool pair = cg == 2;
    CUtensorMap mA, mB;
    rc = make_map(&mA, tz.xb, tz.rows_pad, kp, kGM);
    if (!rc) rc = make_map(&mB, tz.qb, nq_pad, kp, pair ? kGN / 2 : kGN);
    if (rc) return rc;
    GemmArgs a{};
    a.xn2 = tz.xn2; a.tabB = tz.tabB; a.tabT = tz.tabT; a.panelmin = tz.panelmin; a.cand = (uint2 *)tz.cand;
    a.cand_cnt = tz.cand_cnt;
    a.n = (uint32_t)n; a.nq = (uint32_t)nq; a.nq_pad = (uint32_t)nq_pad; a.kchunks = (uint32_t)(kp / kGK);
    a.panels = (uint32_t)panels; a.qtiles = (uint32_t)(nq_pad / kGN); a.cap = (uint32_t)cap_c;
    a.mask = cur_mask;
    a.dbg_chunks = kEpiCols / 32;
    if (const char *e = getenv("B200HNSW_BF_DEBUG_CHUNKS")) a.dbg_chunks = (uint32_t)std::min(kEpiCols / 32, std::max(0, atoi(e)));  // WRONG RESULTS: timing only
    static bool configured[16] = {};
    if (device < 16 && !configured[device]) {
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<0, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<0, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        B200_CUDA_OK(cudaFuncSetAttribute(bf_gemm_kernel<1, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        configured[device] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const bool ip = prm.metric == B200HNSW_IP;
    // pass 1: bounds from every stride-th panel
    const unsigned tgrid = (unsigned)((nq_pad + 255) / 256);
    bf_tables_kernel<<<tgrid, 256, 0, st>>>(tz.qn2, tz.thr, (uint32_t)nq_pad, ip ? 1 : 0, 0, tz.tabB, tz.tabT);
    a.stride = (uint32_t)stride;
    // one persistent CTA per SM; pairs: one cluster of two per TPC, both CTAs walk the same items
    auto launch_gemm = [&](int mode, size_t npanels) -> int {
        if (!pair) {
            const unsigned grid = (unsigned)std::min<size_t>((size_t)sms, npanels * a.qtiles);
            if (ip) { if (mode) bf_gemm_kernel<1, 1><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a); else bf_gemm_kernel<1, 0><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a); }
            else { if (mode) bf_gemm_kernel<0, 1><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a); else bf_gemm_kernel<0, 0><<<grid, kGThreads, kGemmSmem, st>>>(mA, mB, a); }
            B200_CUDA_OK(cudaGetLastError());
            return 0;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (unsigned)std::min<size_t>((size_t)sms / 2, (npanels + 1) / 2 * a.qtiles));
        cfg.blockDim = dim3(kGThreads);
        cfg.dynamicSmemBytes = kGemmSmem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        auto *fn = ip ? (mode ? bf_gemm_kernel<1, 1, 2> : bf_gemm_kernel<1, 0, 2>) : (mode ? bf_gemm_kernel<0, 1, 2> : bf_gemm_kernel<0, 0, 2>);
        B200_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, mA, mB, a));
        return 0;
    };
    rc = launch_gemm(0, sampled);
    if (rc) return rc;
    mark();
    const size_t kth_smem = sampled * kKthQ * 4;
    if (kth_smem <= 160 * 1024) {
        static bool kcfg[16] = {};
        if (device >= 16 || !kcfg[device]) {
            B200_CUDA_OK(cudaFuncSetAttribute(bf_kth_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            if (device < 16) kcfg[device] = true;
        }
        bf_kth_smem_kernel<<<(unsigned)(nq_pad / kKthQ), 256, kth_smem, st>>>(tz.panelmin, (uint32_t)sampled, (uint32_t)nq_pad,
                                                                             (uint32_t)nq, (uint32_t)k, tz.thr);
    } else {
        bf_kth_kernel<<<(unsigned)(nq_pad / 32), 256, 0, st>>>(tz.panelmin, (uint32_t)sampled, (uint32_t)nq_pad,
                                                              (uint32_t)nq, (uint32_t)k, tz.thr);
    }
    mark();
    // pass 2: candidates from all panels
    bf_tables_kernel<<<tgrid, 256, 0, st>>>(tz.qn2, tz.thr, (uint32_t)nq_pad, ip ? 1 : 0, 1, tz.tabB, tz.tabT);
    a.stride = 1;
    rc = launch_gemm(1, panels);
    if (rc) return rc;
    mark();
    // exact re-rank
    const size_t dim = host.dim;
    size_t lane_floats;
    if (dim % 4 == 0) lane_floats = dim;
    else if (dim > 16) lane_floats = dim >> 4 << 4;
    else if (dim > 4) lane_floats = dim >> 2 << 2;
    else lane_floats = 0;
    // survivors (k + the rows inside the error band) are a small fraction of the candidates; a query with more of them
    // than scap counts as an overflow and takes the retry / exact-scan path below
    const size_t scap = (std::min(cap_c, std::max(cap_c / 4, 4 * k)) + 3) & ~(size_t)3;
    const size_t rsm = ((((size_t)d4 + 3) & ~(size_t)3) + 4) * 16 + scap * 4 + (cap_c + (cap_c & 1)) * 4 + scap * 8 +
                       std::max<size_t>((size_t)kRrRows * kRrStride, 2 * cap_c) * 4;
    if (ip)
        launch_rerank<1>((unsigned)nq, rsm, st, dX, dLabels, tz.xn2, tz.qn2, dQ_, (uint32_t)dim, (uint32_t)d4,
                         (uint32_t)(lane_floats / 4), (const uint2 *)tz.cand, tz.cand_cnt, (uint32_t)cap_c, (uint32_t)scap,
                         (uint32_t)k, (uint32_t)(cur_mask ? cur_mask_rows : n), dl, dd, dc, tz.overflow);
    else
        launch_rerank<0>((unsigned)nq, rsm, st, dX, dLabels, tz.xn2, tz.qn2, dQ_, (uint32_t)dim, (uint32_t)d4,
                         (uint32_t)(lane_floats / 4), (const uint2 *)tz.cand, tz.cand_cnt, (uint32_t)cap_c, (uint32_t)scap,
                         (uint32_t)k, (uint32_t)(cur_mask ? cur_mask_rows : n), dl, dd, dc, tz.overflow);
    mark();
    B200_CUDA_OK(cudaGetLastError());
    uint32_t ov = 0;
    B200_CUDA_OK(cudaMemcpyAsync(&ov, tz.overflow, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
    if (prof && pn == 5) {
        float t[4];
        for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], pe[i], pe[i + 1]);
        std::vector<uint32_t> cc(nq);
        cudaMemcpy(cc.data(), tz.cand_cnt, nq * 4, cudaMemcpyDeviceToHost);
        size_t tot = 0, mx = 0;
        for (uint32_t v : cc) { tot += v; mx = std::max<size_t>(mx, v); }
        fprintf(stderr, "[b200bf profile] nq %zu: pass1 %.3f ms, kth %.3f ms, pass2 %.3f ms, rerank %.3f ms; candidates/query "
                        "%.0f (max %zu, cap %zu)\n", nq, t[0], t[1], t[2], t[3], (double)tot / nq, mx, cap_c);
    }
    for (int i = 0; i < pn; i++) cudaEventDestroy(pe[i]);
    stats.kernel_launches += 8;
    if (getenv("B200HNSW_BF_STATS")) {  // diagnostic: candidates generated per batch (costs a D2H copy)
        std::vector<uint32_t> c(nq);
        B200_CUDA_OK(cudaMemcpy(c.data(), tz.cand_cnt, nq * 4, cudaMemcpyDeviceToHost));
        size_t tot = 0;
        for (uint32_t v : c) tot += v;
        tz.last_candidates = tot;
    }
    if (ov && cap_c < 16384) {  // rare: give every query 4x the slots and redo the batch on the tensor path
        tz.cap_floor = cap_c * 4;
        return search_tensor(dQ_, nq, k, dl, dd, dc, st);
    }
    return ov ? 1 : 0;
}

}  // namespace b200

// bf_stream.cu -- exact brute-force scan for SMALL query batches (reference: hnswlib/bruteforce.h:106-135).
//
// With a handful of queries BruteforceSearch::searchKnn is one pass over the stored rows: HBM-bound up to ~8 queries
// (3 KB of row per 1.5 k non-fused multiply/add pairs per query), FP32-issue-bound from there until the tensor-core
// path (bf_tensor.cu) takes over.  The tiled scan in bruteforce.cu is a GEMM-shaped kernel (64 queries x slices of
// rows, at most 64 CTAs for one query tile) and cannot stream: this kernel is the streaming one.
//
//   * persistent grid, one CTA per SM, CTA b walks row tiles b, b+grid, ... of 64 rows;
//   * a producer warp moves 64 rows x 512 B per stage with 16-byte cp.async copies (one warp instruction = 512
//     contiguous bytes of one row, completion on an mbarrier via cp.async.mbarrier.arrive) into a ring of 32 KB stages laid out
//     like a 128-byte-swizzled TMA tile: the eight 16 B chunks of a 128 B row segment are XOR-ed with the row number,
//     so lane = row reads are conflict-free without padding.  Measured alternatives on 1M x 768, one query: one 512 B
//     1-D bulk copy (cp.async.bulk) per row 4.5 ms (64 small TMA operations per stage); four 64-row x 128 B swizzled
//     2-D TMA boxes per stage 0.84 ms; this LDGSTS producer 0.86 ms -- the last two are bound by the compute warps,
//     not by the copy engine;
//   * compute thread = (row of the tile, query slot): it owns the four lane accumulators + sequential tail of the
//     reference's SSE kernels for P queries (space_l2.h:97-143, space_ip.h:255-303: separate multiply and add,
//     ((s0+s1)+s2)+s3 + tail) -- query chunks are broadcast reads, so distances are bit-identical to the CPU;
//   * per row tile, distances not worse than the query's current k-th are pushed to a small shared queue and one
//     warp per query folds the queue into that query's unsorted k-best list (replace-worst, ordered by (dist,label));
//   * every CTA writes its k best per query; merge_tree() (merge_launch.cuh) reduces the per-CTA lists.
#include <algorithm>
#include <cstring>

#include "bruteforce.cuh"
#include "bf_topk.cuh"
#include "mbarrier.cuh"
#include "merge_launch.cuh"

namespace b200 {

constexpr int kStRows = 64;                     // rows per tile
constexpr int kStKC4 = 32;                      // 128-bit chunks per row per stage (512 B)
constexpr int kStMaxStages = 8;                 // ring depth is chosen at launch: as many 32 KB stages as fit (<= 6 in 227 KB)
constexpr int kStBoxBytes = kStRows * 128;      // one TMA box: 64 rows x 32 floats
constexpr int kStStageBytes = kStRows * kStKC4 * 16;

struct StreamSmem {
    uint32_t off_stage, off_q, off_topl, off_topd, off_ql, off_qd, off_meta, total;
    __host__ __device__ StreamSmem(uint32_t nqt, uint32_t d4, uint32_t k, uint32_t stages) {
        uint32_t o = 0;
        off_stage = o; o += stages * kStStageBytes + 1024;  // + slack: the ring is aligned to 1024 B at run time
        off_q = o;     o += nqt * d4 * 16;
        off_topl = o;  o += nqt * k * 8;
        off_ql = o;    o += nqt * kStRows * 8;
        off_topd = o;  o += nqt * k * 4;
        off_qd = o;    o += nqt * kStRows * 4;
        off_meta = o;  o += nqt * 4 * 4;  // cnt, worst position, queue length, threshold (float bits)
        total = o;
    }
};

template <int METRIC>
__device__ __forceinline__ float st_term(float q, float x) {
    if (METRIC == 0) {
        const float df = __fsub_rn(q, x);
        return __fmul_rn(df, df);
    }
    return __fmul_rn(q, x);
}

// QS query slots x P queries per slot = queries per pass; 64*QS compute threads + one producer warp.
template <int METRIC, int QS, int P>
__global__ void __launch_bounds__(64 * QS + 32, 1)
    bf_stream_kernel(const float4 *__restrict__ X, const uint64_t *__restrict__ labels, uint32_t n, uint32_t d4,
                     uint32_t lane_chunks, uint32_t dim, const float *__restrict__ Q, uint32_t nq, uint32_t k,
                     uint32_t stages, float *__restrict__ part_d, uint64_t *__restrict__ part_l,
                     const uint8_t *__restrict__ mask) {
    constexpr int NQT = QS * P, NCOMP = 64 * QS, NCW = NCOMP / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[kStMaxStages], empty[kStMaxStages];
    const StreamSmem L(NQT, d4, k, stages);
    unsigned char *ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // 128-byte swizzle needs 1024 B alignment
    float4 *sQ = (float4 *)(smem + L.off_q);
    uint64_t *topl = (uint64_t *)(smem + L.off_topl);
    float *topd = (float *)(smem + L.off_topd);
    uint64_t *ql = (uint64_t *)(smem + L.off_ql);
    float *qd = (float *)(smem + L.off_qd);
    int *meta = (int *)(smem + L.off_meta);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t q0 = blockIdx.y * NQT;
    const uint32_t ntiles = (n + kStRows - 1) / kStRows;
    const uint32_t nkc = (d4 + kStKC4 - 1) / kStKC4;

    if (tid == 0) {
        for (uint32_t s = 0; s < stages; s++) {
            mbar_init(full + s, 32);  // one deferred arrival per producer lane
            mbar_init(empty + s, NCW);
        }
        mbar_init_fence();
    }
    {   // queries of this pass, zero-padded to d4 chunks (rows of Q are only 4-byte aligned in general)
        float *sQf = (float *)sQ;
        const uint32_t per = d4 * 4;
        for (uint32_t i = tid; i < NQT * per; i += blockDim.x) {
            const uint32_t qq = i / per, c = i % per;
            sQf[i] = (q0 + qq < nq && c < dim) ? __ldg(Q + (size_t)(q0 + qq) * dim + c) : 0.f;
        }
        for (int i = tid; i < NQT; i += blockDim.x) {
            meta[i * 4 + 0] = 0;
            meta[i * 4 + 1] = 0;
            meta[i * 4 + 2] = 0;
            meta[i * 4 + 3] = 0x7f800000;  // +inf: everything is admitted until the list is full
        }
    }
    __syncthreads();

    if (warp == NCW) {
        // ---- producer warp: lane = 16-byte chunk of the stage's 512-byte row segment
        uint32_t s = 0, ph = 0;
        const uint32_t slot = (uint32_t)(lane >> 3) * kStBoxBytes;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const uint32_t rows = min((uint32_t)kStRows, n - tile * kStRows);
            for (uint32_t kc = 0; kc < nkc; kc++, s = (s + 1 == stages ? 0 : s + 1), ph ^= (s == 0)) {
                const uint32_t ci = kc * kStKC4 + lane;
                mbar_wait(empty + s, ph ^ 1);
                if (ci < d4) {  // rows / chunks past the end are never read by the compute threads
                    unsigned char *dst = ring + s * kStStageBytes + slot;
                    const float4 *src = X + (size_t)tile * kStRows * d4 + ci;
#pragma unroll 8
                    for (uint32_t r = 0; r < rows; r++)
                        cp_async16(dst + r * 128 + (((uint32_t)(lane & 7) ^ (r & 7)) << 4), src + (size_t)r * d4);
                }
                cp_async_arrive(full + s);
            }
        }
        return;
    }

    // ---- compute threads
    const int r = tid & (kStRows - 1), qs = tid / kStRows;
    const uint32_t sw = (uint32_t)(r & 7);
    uint32_t xo[8];
#pragma unroll
    for (int j = 0; j < 8; j++) xo[j] = ((uint32_t)j ^ sw) << 4;
    uint32_t s = 0, ph = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float acc[P][4], tail[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            tail[p] = 0.f;
#pragma unroll
            for (int l = 0; l < 4; l++) acc[p][l] = 0.f;
        }
        for (uint32_t kc = 0; kc < nkc; kc++, s = (s + 1 == stages ? 0 : s + 1), ph ^= (s == 0)) {
            const uint32_t base = kc * kStKC4;
            const uint32_t c4 = min((uint32_t)kStKC4, d4 - base);
            const uint32_t nl = base < lane_chunks ? min(c4, lane_chunks - base) : 0u;
            mbar_wait(full + s, ph);
            // chunk c of row r lives in box c/8 at 16-byte slot (c%8) ^ (r%8) of the row's 128-byte segment
            const unsigned char *sx = ring + s * kStStageBytes + r * 128;
            const float4 *sq = sQ + (size_t)(qs * P) * d4 + base;
            // whole boxes inside the lane part: the eight slot offsets are per-thread constants (xo[]), so a chunk costs
            // two LDS.128 + 8 FP instructions per query and no address arithmetic
            const uint32_t nfull = nl >> 3;
            for (uint32_t b = 0; b < nfull; b++) {
                const unsigned char *bx = sx + b * kStBoxBytes;
                const float4 *bq = sq + b * 8;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 x = *(const float4 *)(bx + xo[j]);
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        const float4 q = bq[(size_t)p * d4 + j];
                        acc[p][0] = __fadd_rn(acc[p][0], st_term<METRIC>(q.x, x.x));
                        acc[p][1] = __fadd_rn(acc[p][1], st_term<METRIC>(q.y, x.y));
                        acc[p][2] = __fadd_rn(acc[p][2], st_term<METRIC>(q.z, x.z));
                        acc[p][3] = __fadd_rn(acc[p][3], st_term<METRIC>(q.w, x.w));
                    }
                }
            }
            for (uint32_t c = nfull * 8; c < nl; c++) {
                const float4 x = *(const float4 *)(sx + (c >> 3) * kStBoxBytes + (((c & 7) ^ sw) << 4));
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const float4 q = sq[(size_t)p * d4 + c];
                    acc[p][0] = __fadd_rn(acc[p][0], st_term<METRIC>(q.x, x.x));
                    acc[p][1] = __fadd_rn(acc[p][1], st_term<METRIC>(q.y, x.y));
                    acc[p][2] = __fadd_rn(acc[p][2], st_term<METRIC>(q.z, x.z));
                    acc[p][3] = __fadd_rn(acc[p][3], st_term<METRIC>(q.w, x.w));
                }
            }
            for (uint32_t c = nl; c < c4; c++) {  // sequential tail of the reference's residual kernels
                const float4 x = *(const float4 *)(sx + (c >> 3) * kStBoxBytes + (((c & 7) ^ sw) << 4));
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const float4 q = sq[(size_t)p * d4 + c];
                    tail[p] = __fadd_rn(tail[p], st_term<METRIC>(q.x, x.x));
                    tail[p] = __fadd_rn(tail[p], st_term<METRIC>(q.y, x.y));
                    tail[p] = __fadd_rn(tail[p], st_term<METRIC>(q.z, x.z));
                    tail[p] = __fadd_rn(tail[p], st_term<METRIC>(q.w, x.w));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        // distances of this row -> candidate queues
        const uint32_t g = tile * kStRows + r;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const uint32_t qq = qs * P + p;
            float d = __fadd_rn(__fadd_rn(__fadd_rn(acc[p][0], acc[p][1]), acc[p][2]), acc[p][3]);
            d = __fadd_rn(d, tail[p]);
            if (METRIC == 1) d = __fsub_rn(1.0f, d);
            if (g < n && q0 + qq < nq && d <= __int_as_float(meta[qq * 4 + 3]) && (!mask || mask[g])) {
                const int pos = atomicAdd(&meta[qq * 4 + 2], 1);
                qd[qq * kStRows + pos] = d;
                ql[qq * kStRows + pos] = __ldg(labels + g);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCOMP) : "memory");
        // one warp per query folds its queue into the k-best list
        for (uint32_t qq = warp; qq < (uint32_t)NQT; qq += NCW) {
            const int m = meta[qq * 4 + 2];
            if (m == 0) continue;
            float *td = topd + (size_t)qq * k;
            uint64_t *tl = topl + (size_t)qq * k;
            int cnt = meta[qq * 4 + 0], wpos = meta[qq * 4 + 1];
            float wd = cnt == (int)k ? td[wpos] : 0.f;
            uint64_t wl = cnt == (int)k ? tl[wpos] : 0;
            for (int e = 0; e < m; e++)
                topk_insert(td, tl, (int)k, cnt, wpos, wd, wl, qd[qq * kStRows + e], ql[qq * kStRows + e], lane);
            if (lane == 0) {
                meta[qq * 4 + 0] = cnt;
                meta[qq * 4 + 1] = wpos;
                meta[qq * 4 + 2] = 0;
                meta[qq * 4 + 3] = cnt == (int)k ? __float_as_int(wd) : 0x7f800000;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCOMP) : "memory");
    }
    // this CTA's lists -> global partials, padded with (+inf, UINT64_MAX)
    for (uint32_t e = tid; e < NQT * k; e += NCOMP) {
        const uint32_t qq = e / k, j = e % k;
        if (q0 + qq >= nq) continue;
        const int cnt = meta[qq * 4 + 0];
        const size_t o = ((size_t)blockIdx.x * nq + q0 + qq) * k + j;
        if ((int)j < cnt) {
            part_d[o] = topd[(size_t)qq * k + j];
            part_l[o] = topl[(size_t)qq * k + j];
        } else {
            part_d[o] = __int_as_float(0x7f800000);
            part_l[o] = 0xFFFFFFFFFFFFFFFFull;
        }
    }
}

__global__ void bf_stream_counts_kernel(uint32_t *counts, uint32_t nq, uint32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) counts[i] = v;
}

template <int METRIC, int QS, int P>
static cudaError_t stream_launch(dim3 grid, uint32_t smem_bytes, cudaStream_t st, const float4 *X, const uint64_t *labels,
                                 uint32_t n, uint32_t d4, uint32_t lane_chunks, uint32_t dim, const float *Q, uint32_t nq,
                                 uint32_t k, uint32_t stages, float *pd, uint64_t *pl, const uint8_t *mask) {
    cudaError_t e = cudaFuncSetAttribute(bf_stream_kernel<METRIC, QS, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
    if (e != cudaSuccess) return e;
    bf_stream_kernel<METRIC, QS, P><<<grid, 64 * QS + 32, smem_bytes, st>>>(X, labels, n, d4, lane_chunks, dim, Q, nq, k,
                                                                            stages, pd, pl, mask);
    return cudaGetLastError();
}

// returns 1 when the shape does not fit this kernel (the caller then uses the tiled scan)
int BruteIndex::search_stream(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc,
                              cudaStream_t st) {
    const size_t n = host.cur;
    if (n == 0) return 1;
    static int sm_count[16] = {};
    static int smem_optin[16] = {};
    const int di = device < 16 ? device : 0;
    if (!sm_count[di]) {
        B200_CUDA_OK(cudaDeviceGetAttribute(&smem_optin[di], cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        B200_CUDA_OK(cudaDeviceGetAttribute(&sm_count[di], cudaDevAttrMultiProcessorCount, device));
    }
    // static shared memory of the kernel: 2 * kStMaxStages mbarriers
    const size_t budget = (size_t)smem_optin[di] - 2 * kStMaxStages * 8 - 64;
    // queries per pass: the largest of {16, 8, 4, 2, 1} that is useful and fits.  A row chunk read from shared memory
    // serves the P queries of its thread, so P grows before the number of query slots does (with 8 slots x 1 query
    // the rows were re-read 8 times and shared-memory bandwidth, not HBM, set the pace).  The ring takes what is left
    // (measured at one query: 2 / 4 / 6 stages = 0.895 / 0.772 / 0.769 ms, so depth beyond 4 is not the limiter).
    static const int cfg_qs[5] = {4, 4, 4, 2, 1}, cfg_p[5] = {4, 2, 1, 1, 1};
    const char *es = getenv("B200HNSW_BF_STAGES");
    const uint32_t max_stages = es ? (uint32_t)std::min(6, std::max(2, atoi(es))) : 6u;
    int pick = -1;
    uint32_t stages = 0;
    for (int c = 0; c < 5 && pick < 0; c++) {
        const size_t nqt = (size_t)cfg_qs[c] * cfg_p[c];
        if (c + 1 < 5 && nqt / 2 >= nq) continue;  // a smaller pass already covers nq
        for (uint32_t t = max_stages; t >= 3 || (t >= 2 && es); t--)
            if (StreamSmem((uint32_t)nqt, (uint32_t)d4, (uint32_t)k, t).total <= budget) { pick = c; stages = t; break; }
    }
    if (pick < 0) return 1;
    const size_t nqt = (size_t)cfg_qs[pick] * cfg_p[pick];
    const uint32_t smem_bytes = StreamSmem((uint32_t)nqt, (uint32_t)d4, (uint32_t)k, stages).total;
    const size_t ntiles = (n + kStRows - 1) / kStRows;
    const size_t slices = std::min<size_t>(ntiles, (size_t)sm_count[di]);
    const size_t passes = (nq + nqt - 1) / nqt;
    int rc = ensure_part(slices * nq * k);
    if (rc) return rc;
    rc = ensure_part2(merge_tree_scratch(slices, nq, k));
    if (rc) return rc;
    const size_t dim = host.dim;
    size_t lane_floats;  // lane part / sequential tail split of the reference's dispatch ladder (space_l2.h:214-238)
    if (dim % 4 == 0) lane_floats = dim;
    else if (dim > 16) lane_floats = dim >> 4 << 4;
    else if (dim > 4) lane_floats = dim >> 2 << 2;
    else lane_floats = 0;
    const dim3 grid((unsigned)slices, (unsigned)passes);
    const int m = prm.metric == B200HNSW_L2 ? 0 : 1;

    cudaError_t e;
#define B200_STREAM_CASE(QS_, P_)                                                                                       \
    e = m == 0 ? stream_launch<0, QS_, P_>(grid, smem_bytes, st, dX, dLabels, (uint32_t)n, (uint32_t)d4,                \
                                           (uint32_t)(lane_floats / 4), (uint32_t)dim, dQ_, (uint32_t)nq, (uint32_t)k,  \
                                           stages, dPartD, dPartL, cur_mask)                                            \
               : stream_launch<1, QS_, P_>(grid, smem_bytes, st, dX, dLabels, (uint32_t)n, (uint32_t)d4,                \
                                           (uint32_t)(lane_floats / 4), (uint32_t)dim, dQ_, (uint32_t)nq, (uint32_t)k,  \
                                           stages, dPartD, dPartL, cur_mask)
    switch (pick) {
        case 0: B200_STREAM_CASE(4, 4); break;
        case 1: B200_STREAM_CASE(4, 2); break;
        case 2: B200_STREAM_CASE(4, 1); break;
        case 3: B200_STREAM_CASE(2, 1); break;
        default: B200_STREAM_CASE(1, 1); break;
    }
#undef B200_STREAM_CASE
    B200_CUDA_OK(e);
    unsigned merges = 0;
    B200_CUDA_OK(merge_tree(dPartL, dPartD, slices, nq, k, dPart2L, dPart2D, dl, dd, st, &merges));
    if (dc)
        bf_stream_counts_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(
            dc, (uint32_t)nq, (uint32_t)std::min(k, cur_mask ? cur_mask_rows : n));
    stats.kernel_launches += 1 + merges;
    return 0;
}

}  // namespace b200

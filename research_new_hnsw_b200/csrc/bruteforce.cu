// bruteforce.cu -- exact batched brute-force scan (reference: hnswlib/bruteforce.h:106-135).
#include <algorithm>
#include <cstring>
#include <vector>

#include "bruteforce.cuh"
#include "bf_topk.cuh"
#include "merge_launch.cuh"

namespace b200 {

constexpr int kBfQT = 64;       // queries per CTA tile
constexpr int kBfRT = 64;       // rows per tile
constexpr int kBfKC = 16;       // floats of K per stage (4 x 128-bit)
constexpr int kBfStride = 20;   // padded smem row stride in floats: 80 B steps keep 128-bit reads conflict-free
constexpr int kBfThreads = 256;

struct BfSmem {
    uint32_t off_q, off_x, off_dt, off_topd, off_topl, off_meta, total;
    __host__ __device__ explicit BfSmem(uint32_t k) {
        uint32_t o = 0;
        off_q = o;    o += kBfQT * kBfStride * 4;
        off_x = o;    o += kBfRT * kBfStride * 4;
        off_dt = o;   o += kBfQT * (kBfRT + 1) * 4;
        o = (o + 7) & ~7u;
        off_topl = o; o += kBfQT * k * 8;
        off_topd = o; o += kBfQT * k * 4;
        off_meta = o; o += kBfQT * 2 * 4;  // cnt, worst position
        total = o;
    }
};

// One CTA: 64 queries x one slice of rows.  Distances are accumulated exactly like the reference's SSE kernels:
// lane_chunks 128-bit chunks go to four per-lane accumulators, the remaining (< 16) elements to a sequential
// tail, final = ((s0+s1)+s2)+s3 + tail (space_l2.h:149-160 for the residual split).  Each query keeps its k best
// (dist, label) pairs of the slice unsorted in shared memory next to the position of the worst one.
template <int METRIC>
__global__ void __launch_bounds__(kBfThreads) bf_scan_kernel(const float4 *__restrict__ X,
                                                             const uint64_t *__restrict__ labels, uint32_t n,
                                                             uint32_t d4, uint32_t lane_chunks, uint32_t dim,
                                                             const float *__restrict__ Q, uint32_t nq, uint32_t k,
                                                             uint32_t rows_per_slice, float *__restrict__ part_d,
                                                             uint64_t *__restrict__ part_l,
                                                             const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem[];
    const BfSmem L(k);
    float *sQ = (float *)(smem + L.off_q);
    float *sX = (float *)(smem + L.off_x);
    float *dt = (float *)(smem + L.off_dt);
    uint64_t *topl = (uint64_t *)(smem + L.off_topl);
    float *topd = (float *)(smem + L.off_topd);
    int *meta = (int *)(smem + L.off_meta);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const uint32_t slice = blockIdx.x, q0 = blockIdx.y * kBfQT;
    const uint32_t r_begin = slice * rows_per_slice;
    const uint32_t r_end = min(n, r_begin + rows_per_slice);

    if (tid < kBfQT) { meta[tid * 2] = 0; meta[tid * 2 + 1] = 0; }
    __syncthreads();

    for (uint32_t r0 = r_begin; r0 < r_end; r0 += kBfRT) {
        float s[4][4][4];
        float t[4][4];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                t[j][i] = 0.f;
#pragma unroll
                for (int l = 0; l < 4; l++) s[j][i][l] = 0.f;
            }
        for (uint32_t c0 = 0; c0 < d4; c0 += kBfKC / 4) {
            __syncthreads();
            {   // X tile: one 128-bit load per thread
                const int r = tid >> 2, c = tid & 3;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + r < r_end && c0 + c < d4) v = __ldg(X + (size_t)(r0 + r) * d4 + c0 + c);
                *(float4 *)(sX + r * kBfStride + c * 4) = v;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {  // Q tile: rows of Q are only 4-byte aligned in general
                const int idx = tid + e * kBfThreads;
                const int r = idx >> 4, c = idx & 15;
                const uint32_t col = c0 * 4 + c;
                float v = 0.f;
                if (q0 + r < nq && col < dim) v = __ldg(Q + (size_t)(q0 + r) * dim + col);
                sQ[r * kBfStride + c] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kBfKC / 4; kk++) {
                if (c0 + kk >= d4) break;
                float4 qv[4], xv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) qv[j] = *(const float4 *)(sQ + (ty * 4 + j) * kBfStride + kk * 4);
#pragma unroll
                for (int i = 0; i < 4; i++) xv[i] = *(const float4 *)(sX + (tx + 16 * i) * kBfStride + kk * 4);
                if (c0 + kk < lane_chunks) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const float a[4] = {qv[j].x, qv[j].y, qv[j].z, qv[j].w};
                            const float b[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                            for (int l = 0; l < 4; l++) {
                                float m;
                                if (METRIC == 0) {
                                    const float df = __fsub_rn(a[l], b[l]);
                                    m = __fmul_rn(df, df);
                                } else {
                                    m = __fmul_rn(a[l], b[l]);
                                }
                                s[j][i][l] = __fadd_rn(s[j][i][l], m);
                            }
                        }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const float a[4] = {qv[j].x, qv[j].y, qv[j].z, qv[j].w};
                            const float b[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                            for (int l = 0; l < 4; l++) {
                                float m;
                                if (METRIC == 0) {
                                    const float df = __fsub_rn(a[l], b[l]);
                                    m = __fmul_rn(df, df);
                                } else {
                                    m = __fmul_rn(a[l], b[l]);
                                }
                                t[j][i] = __fadd_rn(t[j][i], m);
                            }
                        }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float r = __fadd_rn(__fadd_rn(__fadd_rn(s[j][i][0], s[j][i][1]), s[j][i][2]), s[j][i][3]);
                r = __fadd_rn(r, t[j][i]);
                if (METRIC == 1) r = __fsub_rn(1.0f, r);
                dt[(ty * 4 + j) * (kBfRT + 1) + tx + 16 * i] = r;
            }
        __syncthreads();
        // selection: warp w owns queries w*8 .. w*8+7 of the tile
        for (int qq = warp * 8; qq < warp * 8 + 8; qq++) {
            if (q0 + qq >= nq) break;
            float *td = topd + (size_t)qq * k;
            uint64_t *tl = topl + (size_t)qq * k;
            int cnt = meta[qq * 2], wpos = meta[qq * 2 + 1];
            float wd = cnt == (int)k ? td[wpos] : 0.f;
            uint64_t wl = cnt == (int)k ? tl[wpos] : 0;
            for (int half = 0; half < 2; half++) {
                const int r = lane + half * 32;
                const uint32_t g = r0 + r;
                // mask: the caller's BaseFilterFunctor verdict per row (bruteforce.h:114,121); null = no filter
                const bool valid = g < r_end && (!mask || mask[g]);
                const float d = valid ? dt[qq * (kBfRT + 1) + r] : 0.f;
                uint64_t lab = 0;
                bool pend = valid && (cnt < (int)k || d <= wd);
                if (pend) lab = __ldg(labels + g);
                uint32_t m = __ballot_sync(0xffffffffu, pend);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float cd = __shfl_sync(0xffffffffu, d, src);
                    const uint64_t cl = __shfl_sync(0xffffffffu, lab, src);
                    topk_insert(td, tl, (int)k, cnt, wpos, wd, wl, cd, cl, lane);
                }
            }
            if (lane == 0) { meta[qq * 2] = cnt; meta[qq * 2 + 1] = wpos; }
        }
    }
    __syncthreads();
    // slice result -> global partials, padded with (+inf, UINT64_MAX)
    for (uint32_t e = tid; e < kBfQT * k; e += kBfThreads) {
        const uint32_t qq = e / k, j = e % k;
        if (q0 + qq >= nq) continue;
        const int cnt = meta[qq * 2];
        const size_t o = ((size_t)slice * nq + q0 + qq) * k + j;
        if ((int)j < cnt) {
            part_d[o] = topd[(size_t)qq * k + j];
            part_l[o] = topl[(size_t)qq * k + j];
        } else {
            part_d[o] = __int_as_float(0x7f800000);
            part_l[o] = 0xFFFFFFFFFFFFFFFFull;
        }
    }
}

__global__ void bf_counts_kernel(uint32_t *counts, uint32_t nq, uint32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) counts[i] = v;
}

__global__ void bf_pad_rows_kernel(const float *__restrict__ X, uint32_t dim, uint32_t d4, size_t rows,
                                   float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d4 * 4) return;
    const size_t r = i / (d4 * 4), c = i % (d4 * 4);
    out[i] = c < dim ? X[r * dim + c] : 0.f;
}

BruteIndex::~BruteIndex() {
    tz.release();
    cudaFree(dX); cudaFree(dLabels); cudaFree(dQ); cudaFree(dOutL); cudaFree(dOutD); cudaFree(dPartL);
    cudaFree(dPartD); cudaFree(dPart2L); cudaFree(dPart2D); cudaFree(dCounts); cudaFree(dMask);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
}

int BruteIndex::init_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
        return B200HNSW_E_CUDA;
    }
    int d = prm.device;
    if (d < 0) B200_CUDA_OK(cudaGetDevice(&d));
    if (d >= count) { set_error("device ordinal out of range"); return B200HNSW_E_ARG; }
    device = d;
    B200_CUDA_OK(cudaSetDevice(d));
    B200_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    B200_CUDA_OK(cudaEventCreate(&ev0));
    B200_CUDA_OK(cudaEventCreate(&ev1));
    d4 = (host.dim + 3) / 4;
    cap = host.maxel;
    const size_t c = cap ? cap : 1;
    B200_CUDA_OK(cudaMalloc(&dX, c * d4 * 16));
    B200_CUDA_OK(cudaMalloc(&dLabels, c * 8));
    return 0;
}

int BruteIndex::create(const b200hnsw_params &p) {
    prm = p;
    host.init(p.dim, p.max_elements);
    return init_device();
}

int BruteIndex::load(const char *path, const b200hnsw_params &p) {
    prm = p;
    const int r = host.load(path, p.dim);
    if (r == -1) { set_error("Cannot open file"); return B200HNSW_E_OPEN; }
    if (r) { set_error("Index seems to be corrupted or unsupported"); return B200HNSW_E_CORRUPT; }
    int rc = init_device();
    if (rc) return rc;
    return upload_rows(0, host.cur);
}

// host rows [first, first+count) -> padded device rows + labels
int BruteIndex::upload_rows(size_t first, size_t count) {
    if (!count) return 0;
    B200_CUDA_OK(cudaSetDevice(device));
    const size_t dim = host.dim;
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(count, (size_t)(128u << 20) / (dim * 4)));
    std::vector<float> tmp(chunk * dim);
    std::vector<uint64_t> labs(chunk);
    float *draw = nullptr;
    B200_CUDA_OK(cudaMalloc(&draw, chunk * dim * 4));
    for (size_t s = 0; s < count; s += chunk) {
        const size_t c = std::min(chunk, count - s);
        for (size_t i = 0; i < c; i++) {
            const char *row = host.data.data() + (first + s + i) * host.row;
            memcpy(tmp.data() + i * dim, row, dim * 4);
            memcpy(&labs[i], row + dim * 4, 8);
        }
        cudaError_t e = cudaMemcpy(draw, tmp.data(), c * dim * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            const size_t tot = c * d4 * 4;
            bf_pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(draw, (uint32_t)dim, (uint32_t)d4, c,
                                                                      (float *)(dX + (first + s) * d4));
            e = cudaMemcpy(dLabels + first + s, labs.data(), c * 8, cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            cudaFree(draw);
            set_error(std::string("CUDA error during upload: ") + cudaGetErrorString(e));
            return B200HNSW_E_CUDA;
        }
    }
    cudaFree(draw);
    if (tz.xb) {  // keep the bf16 copy of the tensor path in step
        int rc = tensor_sync_rows(first, count);
        if (rc) return rc;
        B200_CUDA_OK(cudaDeviceSynchronize());
    }
    return 0;
}

// bruteforce.h:64-83
int BruteIndex::add_batch(const float *X, const uint64_t *labels, size_t n) {
    std::lock_guard<std::mutex> g(mu);
    size_t lo = (size_t)-1, hi = 0;
    for (size_t i = 0; i < n; i++) {
        const uint64_t lab = labels ? labels[i] : host.cur;
        size_t idx;
        auto it = host.lookup.find(lab);
        if (it != host.lookup.end()) {
            idx = it->second;
        } else {
            if (host.cur >= host.maxel) {
                set_error("The number of elements exceeds the specified limit\n");
                if (lo != (size_t)-1) upload_rows(lo, hi - lo + 1);
                return B200HNSW_E_CAPACITY;
            }
            idx = host.cur++;
            host.lookup[lab] = idx;
        }
        char *row = host.data.data() + idx * host.row;
        memcpy(row + host.dim * 4, &lab, 8);
        memcpy(row, X + i * host.dim, host.dim * 4);
        lo = std::min(lo, idx);
        hi = std::max(hi, idx);
    }
    if (lo == (size_t)-1) return 0;
    return upload_rows(lo, hi - lo + 1);
}

// bruteforce.h:86-103
int BruteIndex::remove(uint64_t label) {
    std::lock_guard<std::mutex> g(mu);
    auto it = host.lookup.find(label);
    if (it == host.lookup.end()) return 0;
    const size_t c = it->second;
    host.lookup.erase(it);
    uint64_t last;
    memcpy(&last, host.data.data() + host.row * (host.cur - 1) + host.dim * 4, 8);
    host.lookup[last] = c;
    memmove(host.data.data() + host.row * c, host.data.data() + host.row * (host.cur - 1), host.row);
    host.cur--;
    if (c < host.cur) return upload_rows(c, 1);
    return 0;
}

int BruteIndex::ensure_part2(size_t elems) {
    if (elems <= part2_elems) return 0;
    cudaFree(dPart2L); cudaFree(dPart2D);
    dPart2L = nullptr; dPart2D = nullptr; part2_elems = 0;
    B200_CUDA_OK(cudaMalloc(&dPart2L, elems * 8));
    B200_CUDA_OK(cudaMalloc(&dPart2D, elems * 4));
    part2_elems = elems;
    return 0;
}

int BruteIndex::ensure_part(size_t elems) {
    if (elems <= part_elems) return 0;
    cudaFree(dPartL); cudaFree(dPartD);
    dPartL = nullptr; dPartD = nullptr; part_elems = 0;
    B200_CUDA_OK(cudaMalloc(&dPartL, elems * 8));
    B200_CUDA_OK(cudaMalloc(&dPartD, elems * 4));
    part_elems = elems;
    return 0;
}

// Path choice: the tensor-core candidate generator pays off once the scan is a real GEMM (rows x queries large);
// a few queries over many rows stream the rows once (bf_stream.cu); everything else, and anything those two cannot
// take, goes to the tiled exact scan.  B200HNSW_BF_PATH=scan|tensor|stream forces one.
int BruteIndex::search_device(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc,
                              cudaStream_t st) {
    if (nq == 0) return 0;
    if (!dQ_ || !dl || !dd || k == 0) { set_error("search: null pointer or k == 0"); return B200HNSW_E_ARG; }
    B200_CUDA_OK(cudaSetDevice(device));
    const char *force = getenv("B200HNSW_BF_PATH");
    const bool want_scan = force && !strcmp(force, "scan");
    const bool want_tensor = force && !strcmp(force, "tensor");
    const bool want_stream = force && !strcmp(force, "stream");
    const size_t n = host.cur;
    // measured on 1M x 768 (scripts/probe_bf_small.py, round 2): the tensor path reads the bf16 copy of the rows (half the
    // bytes of the fp32 rows the streaming scan walks) and wins at EVERY batch size on a large index -- 0.47 ms against
    // 0.76 ms for one query, 0.49 against 1.47 ms for eight; on small indexes its fixed cost (a dozen launches) does not pay
    const bool tensor_pays = (double)n * (double)nq >= 6.7e7 || n >= 131072;
    if (!want_scan && !want_stream && k <= (cur_mask ? cur_mask_rows : n) && (want_tensor || tensor_pays)) {
        const int rc = search_tensor(dQ_, nq, k, dl, dd, dc, st);
        if (rc <= 0) { last_path = 1; return rc; }  // done, or a real error; rc == 1 -> fall through to the scan
    }
    if (!want_scan && (want_stream || (nq <= 64 && n >= 16384))) {
        const int rc = search_stream(dQ_, nq, k, dl, dd, dc, st);
        if (rc <= 0) { last_path = 2; return rc; }  // rc == 1: shape does not fit the streaming kernel
    }
    last_path = 0;
    return search_scan(dQ_, nq, k, dl, dd, dc, st);
}

int BruteIndex::search_scan(const float *dQ_, size_t nq, size_t k, uint64_t *dl, float *dd, uint32_t *dc,
                            cudaStream_t st) {
    const size_t n = host.cur;
    const BfSmem L((uint32_t)k);
    if (L.total > 227 * 1024) { set_error("k too large for the brute-force kernel"); return B200HNSW_E_UNSUPPORTED; }
    const size_t qtiles = (nq + kBfQT - 1) / kBfQT;
    size_t slices = std::max<size_t>(1, (592 + qtiles - 1) / qtiles);
    slices = std::min(slices, std::max<size_t>(1, n / 4096));
    slices = std::min<size_t>(slices, 64);
    size_t rows_per_slice = n ? (n + slices - 1) / slices : 1;
    rows_per_slice = (rows_per_slice + kBfRT - 1) / kBfRT * kBfRT;
    slices = n ? (n + rows_per_slice - 1) / rows_per_slice : 1;
    int rc = ensure_part(slices * nq * k);
    if (rc) return rc;
    // lane part / sequential tail split of the reference's dispatch ladder (space_l2.h:214-238)
    const size_t dim = host.dim;
    size_t lane_floats;
    if (dim % 4 == 0) lane_floats = dim;
    else if (dim > 16) lane_floats = dim >> 4 << 4;
    else if (dim > 4) lane_floats = dim >> 2 << 2;
    else lane_floats = 0;
    dim3 grid((unsigned)slices, (unsigned)qtiles);
    static bool configured[2][16] = {};
    const int m = prm.metric == B200HNSW_L2 ? 0 : 1;
    if (device < 16 && !configured[m][device]) {
        if (m == 0)
            B200_CUDA_OK(cudaFuncSetAttribute(bf_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else
            B200_CUDA_OK(cudaFuncSetAttribute(bf_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[m][device] = true;
    }
    if (m == 0)
        bf_scan_kernel<0><<<grid, kBfThreads, L.total, st>>>(dX, dLabels, (uint32_t)n, (uint32_t)d4,
                                                             (uint32_t)(lane_floats / 4), (uint32_t)dim, dQ_,
                                                             (uint32_t)nq, (uint32_t)k, (uint32_t)rows_per_slice,
                                                             dPartD, dPartL, cur_mask);
    else
        bf_scan_kernel<1><<<grid, kBfThreads, L.total, st>>>(dX, dLabels, (uint32_t)n, (uint32_t)d4,
                                                             (uint32_t)(lane_floats / 4), (uint32_t)dim, dQ_,
                                                             (uint32_t)nq, (uint32_t)k, (uint32_t)rows_per_slice,
                                                             dPartD, dPartL, cur_mask);
    B200_CUDA_OK(cudaGetLastError());
    rc = ensure_part2(merge_tree_scratch(slices, nq, k));
    if (rc) return rc;
    unsigned merges = 0;
    B200_CUDA_OK(merge_tree(dPartL, dPartD, slices, nq, k, dPart2L, dPart2D, dl, dd, st, &merges));
    if (dc)
        bf_counts_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(dc, (uint32_t)nq,
                                                                      (uint32_t)std::min(k, cur_mask ? cur_mask_rows : n));
    stats.kernel_launches += 1 + merges;
    return 0;
}

int BruteIndex::search_host(const float *Q, size_t nq, size_t k, uint64_t *labels, float *dists, uint32_t *counts,
                            const uint8_t *allowed) {
    if (nq == 0) return 0;
    if (!Q || !labels || !dists || k == 0) { set_error("search: null pointer or k == 0"); return B200HNSW_E_ARG; }
    std::lock_guard<std::mutex> g(mu);
    B200_CUDA_OK(cudaSetDevice(device));
    struct MaskScope {  // the row mask of a filtered call lives exactly as long as the call
        BruteIndex &ix;
        ~MaskScope() { ix.cur_mask = nullptr; ix.cur_mask_rows = 0; }
    } mask_scope{*this};
    if (allowed) {
        const size_t n = host.cur;
        if (mask_cap < std::max<size_t>(n, 1)) {
            cudaFree(dMask);
            dMask = nullptr;
            mask_cap = 0;
            B200_CUDA_OK(cudaMalloc(&dMask, std::max<size_t>(cap, 1)));
            mask_cap = std::max<size_t>(cap, 1);
        }
        size_t rows = 0;
        for (size_t i = 0; i < n; i++) rows += allowed[i] ? 1 : 0;
        if (n) B200_CUDA_OK(cudaMemcpyAsync(dMask, allowed, n, cudaMemcpyHostToDevice, stream));
        cur_mask = dMask;
        cur_mask_rows = rows;
    }
    if (nq > scratch_q || k > scratch_k) {
        const size_t q = std::max(nq, scratch_q), kk = std::max(k, scratch_k);
        cudaFree(dQ); cudaFree(dOutL); cudaFree(dOutD); cudaFree(dCounts);
        dQ = nullptr; dOutL = nullptr; dOutD = nullptr; dCounts = nullptr;
        scratch_q = scratch_k = 0;
        B200_CUDA_OK(cudaMalloc(&dQ, q * host.dim * 4));
        B200_CUDA_OK(cudaMalloc(&dOutL, q * kk * 8));
        B200_CUDA_OK(cudaMalloc(&dOutD, q * kk * 4));
        B200_CUDA_OK(cudaMalloc(&dCounts, q * 4));
        scratch_q = q;
        scratch_k = kk;
    }
    B200_CUDA_OK(cudaMemcpyAsync(dQ, Q, nq * host.dim * 4, cudaMemcpyHostToDevice, stream));
    B200_CUDA_OK(cudaEventRecord(ev0, stream));
    int rc = search_device(dQ, nq, k, dOutL, dOutD, dCounts, stream);
    if (rc) return rc;
    B200_CUDA_OK(cudaEventRecord(ev1, stream));
    B200_CUDA_OK(cudaMemcpyAsync(labels, dOutL, nq * k * 8, cudaMemcpyDeviceToHost, stream));
    B200_CUDA_OK(cudaMemcpyAsync(dists, dOutD, nq * k * 4, cudaMemcpyDeviceToHost, stream));
    if (counts) B200_CUDA_OK(cudaMemcpyAsync(counts, dCounts, nq * 4, cudaMemcpyDeviceToHost, stream));
    B200_CUDA_OK(cudaStreamSynchronize(stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    stats.last_kernel_ms = ms;
    stats.queries = nq;
    stats.dist_evals = (uint64_t)nq * host.cur;
    stats.hops_base = (uint64_t)last_path;  // 1: tcgen05 candidate GEMM + exact re-rank, 0: exact scan
    stats.hops_upper = tz.last_candidates;  // candidates re-ranked (only with B200HNSW_BF_STATS)
    return 0;
}

}  // namespace b200

// bf_topk.cuh -- the k-best list of the exact brute-force kernels (bruteforce.cu, bf_stream.cu).
//
// BruteforceSearch::searchKnn keeps a max-heap of (dist, label) and replaces the top while `dist <= lastdist`
// (bruteforce.h:106-135): the result is the k lexicographically smallest (dist, label) pairs.  On the GPU one warp
// owns an UNSORTED list of k pairs in shared memory plus the position of its worst entry; a candidate either fills
// the list or replaces the worst, after which the worst is recomputed by the whole warp.
#pragma once
#include <cstdint>

namespace b200 {

__device__ __forceinline__ bool pair_less(float d1, uint64_t l1, float d2, uint64_t l2) {
    return d1 < d2 || (d1 == d2 && l1 < l2);
}

// Warp-synchronous (all 32 lanes, identical arguments).  cnt / wpos / wd / wl are the warp-uniform list state:
// element count, position + value of the worst entry (valid once cnt == k).
__device__ __forceinline__ void topk_insert(float *td, uint64_t *tl, int k, int &cnt, int &wpos, float &wd, uint64_t &wl,
                                            float cd, uint64_t cl, int lane) {
    if (cnt < k) {
        if (lane == 0) { td[cnt] = cd; tl[cnt] = cl; }
        cnt++;
        if (cnt < k) return;
    } else {
        if (!pair_less(cd, cl, wd, wl)) return;
        if (lane == 0) { td[wpos] = cd; tl[wpos] = cl; }
    }
    __syncwarp();
    float bd = -3.402823466e+38f;  // recompute the worst (largest (dist, label)) entry
    uint64_t bl = 0;
    int bp = -1;
    for (int e = lane; e < k; e += 32) {
        const float ed = td[e];
        const uint64_t el = tl[e];
        if (bp < 0 || pair_less(bd, bl, ed, el)) { bd = ed; bl = el; bp = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const uint64_t ol = __shfl_xor_sync(0xffffffffu, bl, o);
        const int op = __shfl_xor_sync(0xffffffffu, bp, o);
        if (op >= 0 && (bp < 0 || pair_less(bd, bl, od, ol))) { bd = od; bl = ol; bp = op; }
    }
    wd = bd; wl = bl; wpos = bp;
    __syncwarp();
}

}  // namespace b200

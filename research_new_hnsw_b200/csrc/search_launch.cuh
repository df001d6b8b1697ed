// search_launch.cuh -- launch templates of hnsw_search_kernel.  The instantiations are split over four translation
// units (search_inst_*.cu: metric x {bare-bone f32, deleted-elements / bf16 variants}) so that they compile in parallel;
// hnsw_index.cu dispatches through the four functions declared at the end.
#pragma once
#include <cstdlib>

#include "hnsw_index.cuh"

namespace b200 {

template <int TEAM, int LPV, int CPL, int METRIC, bool NB, int STORE, bool FULL>
static int launch_one(const SearchArgs &a, size_t smem, cudaStream_t st) {
    static bool configured[16] = {};  // per device; set once (benign race: idempotent)
    int d = 0;
    cudaGetDevice(&d);
    if (d < 16 && !configured[d]) {
        cudaFuncAttributes fa;
        B200_CUDA_OK(cudaFuncGetAttributes(&fa, hnsw_search_kernel<TEAM, LPV, CPL, METRIC, NB, STORE, FULL>));
        int optin = 0;
        B200_CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d));
        B200_CUDA_OK(cudaFuncSetAttribute(hnsw_search_kernel<TEAM, LPV, CPL, METRIC, NB, STORE, FULL>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          optin - (int)fa.sharedSizeBytes));
        configured[d] = true;
    }
    // programmatic stream serialization: this grid may start while the previous kernel of the stream drains (the kernel
    // orders its own output writes behind the previous grid with griddepcontrol.wait)
    static const bool pdl = !(getenv("B200HNSW_PDL") && atoi(getenv("B200HNSW_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.nq);
    cfg.blockDim = dim3(TEAM);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, hnsw_search_kernel<TEAM, LPV, CPL, METRIC, NB, STORE, FULL>, a));
    return 0;
}

// Rows of exactly LPV * CPL chunks take the FULL instantiation on the throughput path (bare-bone search, f32 or bf16
// rows, 64- and 128-thread teams); everything else the generic one.
template <int TEAM, int LPV, int CPL, int METRIC, bool NB, int STORE>
static int launch_shape(const SearchArgs &a, size_t smem, cudaStream_t st) {
    static const bool full_on = !(getenv("B200HNSW_FULL") && atoi(getenv("B200HNSW_FULL")) == 0);
    if constexpr (!NB && TEAM >= 64) {
        if (full_on && a.d4 == (uint32_t)(LPV * CPL) && (STORE == 0 || a.d16 * 2 == a.d4))
            return launch_one<TEAM, LPV, CPL, METRIC, NB, STORE, true>(a, smem, st);
    }
    return launch_one<TEAM, LPV, CPL, METRIC, NB, STORE, false>(a, smem, st);
}

template <int TEAM, int METRIC, bool NB = false, int STORE = 0>
static int launch_team(const SearchArgs &a, size_t smem, cudaStream_t st) {
    const uint32_t d4 = a.d4;
    if (d4 <= 8) return launch_shape<TEAM, 8, 1, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 16) return launch_shape<TEAM, 8, 2, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 24) return launch_shape<TEAM, 8, 3, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 32) return launch_shape<TEAM, 8, 4, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 48) return launch_shape<TEAM, 16, 3, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 64) return launch_shape<TEAM, 16, 4, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 96) return launch_shape<TEAM, 32, 3, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 128) return launch_shape<TEAM, 32, 4, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 192) return launch_shape<TEAM, 32, 6, METRIC, NB, STORE>(a, smem, st);
    if (d4 <= 256) return launch_shape<TEAM, 32, 8, METRIC, NB, STORE>(a, smem, st);
    set_error("dimension > 1024 is not supported by the search kernel");
    return B200HNSW_E_UNSUPPORTED;
}

template <int METRIC>
static int launch_metric(const SearchArgs &a, size_t smem, int team, cudaStream_t st) {
    if (team == 32) return launch_team<32, METRIC>(a, smem, st);
    if (team == 64) return launch_team<64, METRIC>(a, smem, st);
    return launch_team<128, METRIC>(a, smem, st);
}

// implemented in search_inst_{l2,ip}_{bare,var}.cu
int search_launch_l2_bare(const SearchArgs &a, size_t smem, int team, cudaStream_t st);
int search_launch_ip_bare(const SearchArgs &a, size_t smem, int team, cudaStream_t st);
// variant: 1 = deleted elements / filter (128-thread teams), 2 = bf16 rows with 64-thread teams, 3 = bf16 rows with 128
int search_launch_l2_var(const SearchArgs &a, size_t smem, int variant, cudaStream_t st);
int search_launch_ip_var(const SearchArgs &a, size_t smem, int variant, cudaStream_t st);

}  // namespace b200

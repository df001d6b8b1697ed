// build_update.cu -- one family of build-kernel instantiations (build_kernels.cuh); the families compile in parallel.
#include "build_kernels.cuh"

namespace b200 {

int build_run_batch_update(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st) {
    return metric == B200HNSW_L2 ? run_batch_metric<0, true, false>(a, smem_search, smem_link, st)
                                 : run_batch_metric<1, true, false>(a, smem_search, smem_link, st);
}

int build_run_update_phase1(int metric, const BuildArgs &a, uint32_t *newlists, cudaStream_t st) {
    return metric == B200HNSW_L2 ? run_update_phase1_metric<0>(a, newlists, st) : run_update_phase1_metric<1>(a, newlists, st);
}

}  // namespace b200

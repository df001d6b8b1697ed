// build_nb.cu -- one family of build-kernel instantiations (build_kernels.cuh); the families compile in parallel.
#include "build_kernels.cuh"

namespace b200 {

int build_run_batch_insert_nb(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st) {
    return metric == B200HNSW_L2 ? run_batch_metric<0, false, true>(a, smem_search, smem_link, st)
                                 : run_batch_metric<1, false, true>(a, smem_search, smem_link, st);
}

int build_run_batch_update_nb(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st) {
    return metric == B200HNSW_L2 ? run_batch_metric<0, true, true>(a, smem_search, smem_link, st)
                                 : run_batch_metric<1, true, true>(a, smem_search, smem_link, st);
}

}  // namespace b200

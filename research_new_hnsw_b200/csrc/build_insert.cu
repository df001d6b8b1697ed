// build_insert.cu -- one family of build-kernel instantiations (build_kernels.cuh); the families compile in parallel.
#include "build_kernels.cuh"

namespace b200 {

int build_run_batch_insert(int metric, const BuildArgs &a, size_t smem_search, size_t smem_link, cudaStream_t st) {
    return metric == B200HNSW_L2 ? run_batch_metric<0, false, false>(a, smem_search, smem_link, st)
                                 : run_batch_metric<1, false, false>(a, smem_search, smem_link, st);
}

}  // namespace b200

// search_inst_l2_bare.cu -- one family of hnsw_search_kernel instantiations (search_launch.cuh); the families compile in parallel.
#include "search_launch.cuh"

namespace b200 {

int search_launch_l2_bare(const SearchArgs &a, size_t smem, int team, cudaStream_t st) {
    return launch_metric<0>(a, smem, team, st);
}

}  // namespace b200

// search_inst_l2_var.cu -- one family of hnsw_search_kernel instantiations (search_launch.cuh); the families compile in parallel.
#include "search_launch.cuh"

namespace b200 {

int search_launch_l2_var(const SearchArgs &a, size_t smem, int variant, cudaStream_t st) {
    if (variant == 1) return launch_team<128, 0, true>(a, smem, st);
    if (variant == 2) return launch_team<64, 0, false, 1>(a, smem, st);
    return launch_team<128, 0, false, 1>(a, smem, st);
}

}  // namespace b200

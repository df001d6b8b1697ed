// build.cu -- batched GPU graph construction behind addPoint (reference: hnswalg.h:1153-1267).
#include "hnsw_index.cuh"

namespace b200 {

int HnswIndex::add_batch(const float *X, const uint64_t *labels, size_t n) {
    (void)X; (void)labels;
    if (n == 0) return 0;
    set_error("GPU graph build is not available in this build");
    return B200HNSW_E_UNSUPPORTED;
}

int HnswIndex::flush() { return 0; }

}  // namespace b200

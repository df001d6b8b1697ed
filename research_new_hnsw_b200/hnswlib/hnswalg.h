// hnswalg.h -- the reference splits its engine over several headers (/root/reference/hnswlib/hnswalg.h); in the GPU drop-in
// everything lives in hnswlib.h, this file only keeps direct includes of "hnswalg.h" compiling.
#pragma once
#include "hnswlib.h"

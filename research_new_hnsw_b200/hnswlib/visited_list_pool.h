// visited_list_pool.h -- the reference splits its engine over several headers (/root/reference/hnswlib/visited_list_pool.h); in the GPU drop-in
// everything lives in hnswlib.h, this file only keeps direct includes of "visited_list_pool.h" compiling.
#pragma once
#include "hnswlib.h"

// space_l2.h -- the reference splits its engine over several headers (/root/reference/hnswlib/space_l2.h); in the GPU drop-in
// everything lives in hnswlib.h, this file only keeps direct includes of "space_l2.h" compiling.
#pragma once
#include "hnswlib.h"

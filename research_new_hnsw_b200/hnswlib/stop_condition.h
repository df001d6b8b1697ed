// stop_condition.h -- the reference splits its engine over several headers (/root/reference/hnswlib/stop_condition.h); in the GPU drop-in
// everything lives in hnswlib.h, this file only keeps direct includes of "stop_condition.h" compiling.
#pragma once
#include "hnswlib.h"

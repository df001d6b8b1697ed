// bruteforce.h -- the reference splits its engine over several headers (/root/reference/hnswlib/bruteforce.h); in the GPU drop-in
// everything lives in hnswlib.h, this file only keeps direct includes of "bruteforce.h" compiling.
#pragma once
#include "hnswlib.h"

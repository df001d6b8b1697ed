/* rlimit_shim.c -> libb200_norlimit.so (LD_PRELOAD for the reference's hnsw_service).
 *
 * hnsw_service/main.cpp:19-22 caps the process at RLIMIT_AS = 2 GB before it constructs the index: that cap is the
 * reference's memory-experiment fault injector for its CPU engine.  A CUDA context reserves far more virtual address
 * space than 2 GB, so with the cap in place no GPU engine can initialise.  Preloading this library turns exactly that
 * one request (RLIMIT_AS) into a no-op and forwards every other setrlimit call unchanged; main.cpp itself stays
 * unmodified. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <sys/resource.h>

typedef int (*setrlimit_fn)(int, const struct rlimit *);

int setrlimit(__rlimit_resource_t resource, const struct rlimit *rlim) {
    static setrlimit_fn real = 0;
    if (!real) real = (setrlimit_fn)dlsym(RTLD_NEXT, "setrlimit");
    if (resource == RLIMIT_AS) {
        fprintf(stderr, "[b200hnsw] RLIMIT_AS request ignored: a CUDA context needs more address space than a CPU engine\n");
        return 0;
    }
    return real ? real((int)resource, rlim) : -1;
}

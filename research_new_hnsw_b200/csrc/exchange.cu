// exchange.cu -- per-shard result exchange between PROCESSES (one per GPU) without a collective kernel.
//
// The processes-per-GPU form of the sharded search has every rank hand its [nq][k] result block to every
// other rank before the merge.  An NCCL all_gather does that with a kernel that occupies SMs on every GPU and spins until
// the slowest rank arrives -- on SMs the next batch's search kernel wants (round 1: 20 % of the step at N = 8).  Here the
// exchange uses no SM at all:
//   * every rank owns a receive area [kDepth][world][block] + flags [kDepth][world] in its HBM, exported with CUDA IPC
//     handles (one node, NVLink/NVSwitch peers);
//   * step s: the rank's search kernel has written its block straight into its own slot; the rank then PUSHES that block
//     into slot [s % kDepth][rank] of every peer with cudaMemcpyAsync (copy engines over NVLink) and, behind each copy in
//     stream order, writes the step number into the peer's flag with cuStreamWriteValue32;
//   * it then enqueues cuStreamWaitValue32(flag[s % kDepth][r] >= s) for every peer r: the stream -- not an SM -- waits
//     until all blocks of step s have landed; the merge kernel behind it reads them in place.
// kDepth = 3 receive areas: a peer can be at most one step ahead of the slowest rank's merge (it needs that rank's
// flag of step s-1 before it can finish its own exchange s-1), so a block is never overwritten while it is being merged.
#include <cuda.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200hnsw.h"
#include "common.cuh"

namespace {
constexpr int kDepth = 3;

typedef CUresult (*WriteValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

bool driver_fn(const char *name, void **out) {
    cudaDriverEntryPointQueryResult qr;
    return cudaGetDriverEntryPoint(name, out, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess && *out;
}

struct IpcDesc {  // what a rank publishes: 2 x 64-byte IPC handles
    cudaIpcMemHandle_t area, flags;
};
static_assert(sizeof(IpcDesc) == B200HNSW_EXCHANGE_DESC_BYTES, "descriptor size is part of the ABI");
}  // namespace

struct b200hnsw_exchange {
    int device = 0;
    size_t world = 0, rank = 0, block = 0;
    unsigned char *area = nullptr;   // [kDepth][world][block] (mine)
    uint32_t *flags = nullptr;       // [kDepth][world] (mine)
    std::vector<unsigned char *> peer_area;  // peers' areas mapped into this process (nullptr for self)
    std::vector<uint32_t *> peer_flags;
    WriteValueFn write_value = nullptr;
    WaitValueFn wait_value = nullptr;
    bool connected = false;
    ~b200hnsw_exchange() {
        cudaSetDevice(device);
        for (size_t r = 0; r < peer_area.size(); r++) {
            if (peer_area[r]) cudaIpcCloseMemHandle(peer_area[r]);
            if (peer_flags[r]) cudaIpcCloseMemHandle(peer_flags[r]);
        }
        cudaFree(area);
        cudaFree(flags);
    }
};

using b200::set_error;

extern "C" {

int b200hnsw_exchange_create(int device, size_t world, size_t rank, size_t block_bytes, b200hnsw_exchange **out,
                             void *desc_out) {
    if (!out || !desc_out || world == 0 || rank >= world || block_bytes == 0 || block_bytes % 8) {
        set_error("exchange_create: bad argument (block_bytes must be a positive multiple of 8)");
        return B200HNSW_E_ARG;
    }
    *out = nullptr;
    b200hnsw_exchange *x = new b200hnsw_exchange();
    x->device = device; x->world = world; x->rank = rank; x->block = block_bytes;
    auto fail = [&](const char *what, cudaError_t e) {
        set_error(std::string(what) + ": " + cudaGetErrorString(e));
        delete x;
        return B200HNSW_E_CUDA;
    };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail("cudaSetDevice", e);
    if (!driver_fn("cuStreamWriteValue32", (void **)&x->write_value) || !driver_fn("cuStreamWaitValue32", (void **)&x->wait_value)) {
        set_error("cuStreamWriteValue32 / cuStreamWaitValue32 are not available from this driver");
        delete x;
        return B200HNSW_E_UNSUPPORTED;
    }
    if ((e = cudaMalloc(&x->area, (size_t)kDepth * world * block_bytes)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&x->flags, (size_t)kDepth * world * 4)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemset(x->flags, 0, (size_t)kDepth * world * 4)) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail("cudaDeviceSynchronize", e);
    IpcDesc d;
    if ((e = cudaIpcGetMemHandle(&d.area, x->area)) != cudaSuccess) return fail("cudaIpcGetMemHandle", e);
    if ((e = cudaIpcGetMemHandle(&d.flags, x->flags)) != cudaSuccess) return fail("cudaIpcGetMemHandle", e);
    memcpy(desc_out, &d, sizeof(d));
    *out = x;
    return 0;
}

int b200hnsw_exchange_connect(b200hnsw_exchange *x, const void *all_descs) {
    if (!x || !all_descs) { set_error("null argument"); return B200HNSW_E_ARG; }
    B200_CUDA_OK(cudaSetDevice(x->device));
    x->peer_area.assign(x->world, nullptr);
    x->peer_flags.assign(x->world, nullptr);
    for (size_t r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        IpcDesc d;
        memcpy(&d, (const unsigned char *)all_descs + r * sizeof(IpcDesc), sizeof(d));
        void *pa = nullptr, *pf = nullptr;
        B200_CUDA_OK(cudaIpcOpenMemHandle(&pa, d.area, cudaIpcMemLazyEnablePeerAccess));
        x->peer_area[r] = (unsigned char *)pa;
        B200_CUDA_OK(cudaIpcOpenMemHandle(&pf, d.flags, cudaIpcMemLazyEnablePeerAccess));
        x->peer_flags[r] = (uint32_t *)pf;
    }
    x->connected = true;
    return 0;
}

void b200hnsw_exchange_destroy(b200hnsw_exchange *x) { delete x; }

int b200hnsw_exchange_slot(b200hnsw_exchange *x, uint32_t step, void **my_block_out, void **all_blocks_out) {
    if (!x) { set_error("null handle"); return B200HNSW_E_ARG; }
    unsigned char *base = x->area + (size_t)(step % kDepth) * x->world * x->block;
    if (my_block_out) *my_block_out = base + x->rank * x->block;
    if (all_blocks_out) *all_blocks_out = base;
    return 0;
}

int b200hnsw_exchange_step(b200hnsw_exchange *x, uint32_t step, void *cuda_stream) {
    if (!x || !x->connected) { set_error("exchange is not connected"); return B200HNSW_E_STATE; }
    if (step == 0) { set_error("steps are numbered from 1 (flags start at 0)"); return B200HNSW_E_ARG; }
    B200_CUDA_OK(cudaSetDevice(x->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t par = step % kDepth;
    const unsigned char *mine = x->area + (par * x->world + x->rank) * x->block;
    for (size_t i = 1; i < x->world; i++) {  // start with the next rank so that the pushes of all ranks fan out
        const size_t r = (x->rank + i) % x->world;
        B200_CUDA_OK(cudaMemcpyAsync(x->peer_area[r] + (par * x->world + x->rank) * x->block, mine, x->block,
                                     cudaMemcpyDeviceToDevice, st));
        if (x->write_value((CUstream)st, (CUdeviceptr)(x->peer_flags[r] + par * x->world + x->rank), step, 0) != CUDA_SUCCESS) {
            set_error("cuStreamWriteValue32 failed");
            return B200HNSW_E_CUDA;
        }
    }
    for (size_t r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        if (x->wait_value((CUstream)st, (CUdeviceptr)(x->flags + par * x->world + r), step, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
            set_error("cuStreamWaitValue32 failed");
            return B200HNSW_E_CUDA;
        }
    }
    return 0;
}

}  // extern "C"

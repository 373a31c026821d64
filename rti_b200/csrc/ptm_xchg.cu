// Cross-GPU merge of the 12 extrema keys over NVLink peer memory (one process per GPU).
//
// The fit path has exactly one exchange step: every GPU needs the global
// per-coefficient extrema before it can quantise.  The payload is 48 bytes, so what matters is
// latency, not bandwidth.  Instead of a collective library call this is two one-block kernels
// around plain peer stores:
//
//   publish : writes this GPU's keys into slot [parity][rank] of EVERY peer's mailbox (P2P
//             st.global over NVLink), fences system-wide, then raises flag [parity][rank] = epoch
//             on every peer.
//   gather  : on each GPU, one lane per rank polls its LOCAL flag (peers wrote it into this GPU's
//             memory), then the block max-reduces the 8 x 12 slots into the caller's keys.
//
// Mailboxes are double buffered by epoch parity: a rank can run at most one exchange ahead of the
// slowest one (its next publish needs its own gather, which needs everybody's publish), so two
// buffers never collide.  The poll is bounded (globaltimer) and reports a timeout instead of
// hanging the GPU if a peer died.
//
// Peer mappings come from CUDA IPC handles that the host processes exchange out of band
// (rti_b200/encoder.py uses torch.distributed's object all-gather for that).
#include <cstring>

#include "../../include/ptm_b200.h"
#include "ptm_device.cuh"
#include "ptm_internal.h"
#include "ptm_kernels.h"

namespace ptm {

struct XchgPeers {
    int *mailbox[kMaxRanks];
};

__global__ void xchg_publish_kernel(XchgPeers peers, int world, int rank, unsigned epoch, const int *keys) {
    const int parity = epoch & 1;
    const int peer = threadIdx.x / kSlotInts, i = threadIdx.x % kSlotInts;
    if (peer < world && i < 2 * PTM_NCOEF) {
        int *slots = peers.mailbox[peer] + XchgLayout::slots_off / sizeof(int);
        slots[(parity * kMaxRanks + rank) * kSlotInts + i] = keys[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        volatile int *flags = peers.mailbox[threadIdx.x] + XchgLayout::flags_off / sizeof(int);
        __threadfence_system();
        flags[parity * kMaxRanks + rank] = (int) epoch;
    }
}

__global__ void xchg_gather_kernel(int *mailbox, int world, unsigned epoch, int *keys, unsigned long long timeout_ns) {
    __shared__ int ok_s;
    const int parity = epoch & 1;
    volatile int *flags = mailbox + XchgLayout::flags_off / sizeof(int);
    int *status = mailbox + XchgLayout::status_off / sizeof(int);
    if (threadIdx.x == 0)
        ok_s = 1;
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (flags[parity * kMaxRanks + threadIdx.x] != (int) epoch) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) {
                ok_s = 0;
                break;
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < 2 * PTM_NCOEF) {
        const volatile int *slots = mailbox + XchgLayout::slots_off / sizeof(int);
        int m = slots[(parity * kMaxRanks + 0) * kSlotInts + threadIdx.x];
        for (int r = 1; r < world; ++r)
            m = max(m, slots[(parity * kMaxRanks + r) * kSlotInts + threadIdx.x]);
        keys[threadIdx.x] = m;
    }
    if (threadIdx.x == 0 && !ok_s)
        status[0] = 1;  // sticky: a peer never showed up
}

}  // namespace ptm

extern "C" int ptmb_internal_fail(const char *what, const char *detail);
extern "C" int ptmb_internal_device(const ptmb_ctx_t *c);
extern "C" void *ptmb_internal_stream(const ptmb_ctx_t *c);

#define CKX(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess)                                             \
            return ptmb_internal_fail(#call, cudaGetErrorString(e_));      \
    } while (0)

extern "C" {

int ptmb_xchg_create(ptmb_ctx_t *c, int rank, int world, ptmb_xchg_t **out, void *handle_out) {
    if (!c || !out || !handle_out || world < 1 || world > ptm::kMaxRanks || rank < 0 || rank >= world)
        return ptmb_internal_fail("ptmb_xchg_create", "bad arguments (world must be 1..16)");
    static_assert(sizeof(cudaIpcMemHandle_t) == PTMB_XCHG_HANDLE_BYTES, "handle size");
    CKX(cudaSetDevice(ptmb_internal_device(c)));
    ptmb_xchg *x = new ptmb_xchg;
    x->device = ptmb_internal_device(c);
    x->rank = rank;
    x->world = world;
    cudaError_t e = cudaMalloc(&x->local, ptm::XchgLayout::bytes);
    if (e == cudaSuccess)
        e = cudaMemset(x->local, 0, ptm::XchgLayout::bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess)
        e = cudaIpcGetMemHandle(&h, x->local);
    if (e != cudaSuccess) {
        cudaFree(x->local);
        delete x;
        return ptmb_internal_fail("ptmb_xchg_create", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, sizeof(h));
    x->mailbox[rank] = x->local;
    *out = x;
    return 0;
}

int ptmb_xchg_connect(ptmb_xchg_t *x, const void *handles) {
    if (!x || !handles)
        return ptmb_internal_fail("ptmb_xchg_connect", "NULL argument");
    CKX(cudaSetDevice(x->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank)
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *) handles + (size_t) r * sizeof(h), sizeof(h));
        void *p = nullptr;
        CKX(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->mailbox[r] = (int *) p;
        x->opened[r] = true;
    }
    return 0;
}

/* keys_dev (this GPU's extrema) -> every peer; then keys_dev <- max over all ranks.  Stream ordered,
 * no host synchronisation.  All ranks must call it the same number of times. */
int ptmb_xchg_allreduce_max(ptmb_ctx_t *c, ptmb_xchg_t *x, int32_t *keys_dev) {
    if (!c || !x || !keys_dev)
        return ptmb_internal_fail("ptmb_xchg_allreduce_max", "NULL argument");
    CKX(cudaSetDevice(x->device));
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const unsigned epoch = ++x->epoch;
    ptm::XchgPeers peers{};
    for (int r = 0; r < x->world; ++r)
        peers.mailbox[r] = x->mailbox[r];
    ptm::xchg_publish_kernel<<<1, ptm::kMaxRanks * ptm::kSlotInts, 0, st>>>(peers, x->world, x->rank, epoch, keys_dev);
    CKX(cudaGetLastError());
    ptm::xchg_gather_kernel<<<1, 32, 0, st>>>(x->local, x->world, epoch, keys_dev, 5000000000ull);
    CKX(cudaGetLastError());
    return 0;
}

/* 0 = every exchange so far completed; 1 = a peer timed out (results after that are undefined).
 * Synchronises the stream. */
int ptmb_xchg_status(ptmb_ctx_t *c, ptmb_xchg_t *x, int *status) {
    if (!c || !x || !status)
        return ptmb_internal_fail("ptmb_xchg_status", "NULL argument");
    CKX(cudaSetDevice(x->device));
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    CKX(cudaMemcpyAsync(status, (char *) x->local + ptm::XchgLayout::status_off, sizeof(int), cudaMemcpyDeviceToHost, st));
    CKX(cudaStreamSynchronize(st));
    return 0;
}

void ptmb_xchg_destroy(ptmb_xchg_t *x) {
    if (!x)
        return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (x->opened[r])
            cudaIpcCloseMemHandle(x->mailbox[r]);
    cudaFree(x->local);
    delete x;
}

}  // extern "C"

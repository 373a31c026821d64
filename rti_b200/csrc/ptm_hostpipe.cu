// Pageable host memory <-> device at PCIe rate: a pinned staging ring per direction, filled and drained
// by a pool of host threads (see ptm_hostpipe.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ptm_hostpipe.h"

namespace ptm {

// ---------------------------------------------------------------------------------------- CopyPool

CopyPool::CopyPool(int threads) {
    for (int i = 1; i < threads; ++i)   // the calling thread is worker 0
        workers_.emplace_back([this] { worker(); });
}

CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        quit_ = true;
    }
    cv_.notify_all();
    for (std::thread &t : workers_)
        t.join();
}

void CopyPool::worker() {
    size_t seen = 0;
    for (;;) {
        const std::vector<Piece> *job;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return quit_ || generation_ != seen; });
            if (quit_)
                return;
            seen = generation_;
            job = job_;
        }
        for (size_t i; (i = next_.fetch_add(1)) < job->size();)
            memcpy((*job)[i].dst, (*job)[i].src, (*job)[i].bytes);
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (--active_ == 0)
                done_cv_.notify_all();
        }
    }
}

void CopyPool::run(std::vector<Piece> &pieces) {
    if (pieces.empty())
        return;
    size_t total = 0;
    for (const Piece &p : pieces)
        total += p.bytes;
    if (workers_.empty() || total < (1u << 20)) {
        for (const Piece &p : pieces)
            memcpy(p.dst, p.src, p.bytes);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_ = &pieces;
        next_.store(0);
        active_ = workers_.size();
        ++generation_;
    }
    cv_.notify_all();
    for (size_t i; (i = next_.fetch_add(1)) < pieces.size();)
        memcpy(pieces[i].dst, pieces[i].src, pieces[i].bytes);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return active_ == 0; });
}

void CopyPool::copy(void *dst, const void *src, size_t bytes) { copy2d(dst, bytes, src, bytes, bytes, 1); }

void CopyPool::copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t rows) {
    // pieces of about total / (4 x threads), at least 256 KB, never across a row boundary
    const size_t total = row_bytes * rows;
    size_t piece = std::max<size_t>(256u << 10, total / (size_t) (4 * threads()) + 1);
    std::vector<Piece> pieces;
    for (size_t r = 0; r < rows; ++r)
        for (size_t o = 0; o < row_bytes; o += piece)
            pieces.push_back(Piece{(uint8_t *) dst + r * dpitch + o, (const uint8_t *) src + r * spitch + o,
                                   std::min(piece, row_bytes - o)});
    run(pieces);
}

// ---------------------------------------------------------------------------------------- HostPipe

bool HostPipe::is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

cudaError_t HostPipe::init(int dev, int threads) {
    device = dev;
    if (threads <= 0)
        threads = (int) std::thread::hardware_concurrency();
    if (const char *v = getenv("PTM_HOST_THREADS"))
        threads = atoi(v);
    threads = std::max(1, std::min(threads, 32));
    pool = new CopyPool(threads);
    cudaError_t e = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking);
    for (int i = 0; i < kSlots && e == cudaSuccess; ++i) {
        e = cudaHostAlloc((void **) &up_buf[i], kChunk, cudaHostAllocDefault);
        if (e == cudaSuccess)
            e = cudaHostAlloc((void **) &down_buf[i], kChunk, cudaHostAllocDefault);
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(&up_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(&down_done[i], cudaEventDisableTiming);
    }
    return e;
}

void HostPipe::destroy() {
    finish();
    for (int i = 0; i < kSlots; ++i) {
        if (up_buf[i])
            cudaFreeHost(up_buf[i]);
        if (down_buf[i])
            cudaFreeHost(down_buf[i]);
        if (up_done[i])
            cudaEventDestroy(up_done[i]);
        if (down_done[i])
            cudaEventDestroy(down_done[i]);
        up_buf[i] = down_buf[i] = nullptr;
        up_done[i] = down_done[i] = nullptr;
    }
    if (up)
        cudaStreamDestroy(up);
    if (down)
        cudaStreamDestroy(down);
    up = down = nullptr;
    delete pool;
    pool = nullptr;
}

cudaError_t HostPipe::h2d(void *dev, const void *host, size_t bytes, cudaStream_t st) {
    return h2d_2d(dev, bytes, host, bytes, bytes, 1, st);
}

cudaError_t HostPipe::h2d_2d(void *dev, size_t dpitch, const void *host, size_t spitch, size_t row_bytes, size_t rows,
                             cudaStream_t st) {
    if (row_bytes == 0 || rows == 0)
        return cudaSuccess;
    if (is_pinned(host)) {
        if (rows == 1)
            return cudaMemcpyAsync(dev, host, row_bytes, cudaMemcpyHostToDevice, st);
        return cudaMemcpy2DAsync(dev, dpitch, host, spitch, row_bytes, rows, cudaMemcpyHostToDevice, st);
    }
    cudaError_t e;
    // whole rows per chunk while they fit, otherwise pieces of one row
    size_t r = 0;
    while (r < rows) {
        const unsigned slot = up_next++ % kSlots;
        if ((e = cudaEventSynchronize(up_done[slot])) != cudaSuccess)   // the DMA that last read this slot
            return e;
        if (row_bytes <= kChunk) {
            const size_t nr = std::min(rows - r, std::max<size_t>(1, kChunk / row_bytes));
            pool->copy2d(up_buf[slot], row_bytes, (const uint8_t *) host + r * spitch, spitch, row_bytes, nr);
            e = cudaMemcpy2DAsync((uint8_t *) dev + r * dpitch, dpitch, up_buf[slot], row_bytes, row_bytes, nr,
                                  cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess)
                e = cudaEventRecord(up_done[slot], st);
            if (e != cudaSuccess)
                return e;
            r += nr;
        } else {
            unsigned s = slot;
            for (size_t o = 0; o < row_bytes; o += kChunk) {
                if (o != 0) {
                    s = up_next++ % kSlots;
                    if ((e = cudaEventSynchronize(up_done[s])) != cudaSuccess)
                        return e;
                }
                const size_t n = std::min(kChunk, row_bytes - o);
                pool->copy(up_buf[s], (const uint8_t *) host + r * spitch + o, n);
                e = cudaMemcpyAsync((uint8_t *) dev + r * dpitch + o, up_buf[s], n, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess)
                    e = cudaEventRecord(up_done[s], st);
                if (e != cudaSuccess)
                    return e;
            }
            ++r;
        }
    }
    return cudaSuccess;
}

cudaError_t HostPipe::drain_one() {
    const unsigned slot = down_head % kSlots;
    cudaError_t e = cudaEventSynchronize(down_done[slot]);
    if (e != cudaSuccess)
        return e;
    const Pending &p = pending[slot];
    pool->copy2d(p.dst, p.dpitch, down_buf[slot], p.row_bytes, p.row_bytes, p.rows);
    ++down_head;
    --down_count;
    return cudaSuccess;
}

cudaError_t HostPipe::d2h(void *host, const void *dev, size_t bytes, cudaStream_t st) {
    if (bytes == 0)
        return cudaSuccess;
    if (is_pinned(host))
        return cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st);
    cudaError_t e;
    for (size_t o = 0; o < bytes; o += kChunk) {
        if (down_count == kSlots && (e = drain_one()) != cudaSuccess)
            return e;
        const unsigned slot = (down_head + down_count) % kSlots;
        const size_t n = std::min(kChunk, bytes - o);
        e = cudaMemcpyAsync(down_buf[slot], (const uint8_t *) dev + o, n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaEventRecord(down_done[slot], st);
        if (e != cudaSuccess)
            return e;
        pending[slot] = Pending{(uint8_t *) host + o, n, n, 1};
        ++down_count;
    }
    return cudaSuccess;
}

cudaError_t HostPipe::finish() {
    while (down_count) {
        const cudaError_t e = drain_one();
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------- HostState

cudaError_t HostState::reserve(int i, size_t need) {
    if (scratch_cap[i] >= need)
        return cudaSuccess;
    cudaDeviceSynchronize();
    cudaFree(scratch[i]);
    scratch[i] = nullptr;
    scratch_cap[i] = 0;
    const cudaError_t e = cudaMalloc(&scratch[i], need);
    if (e == cudaSuccess)
        scratch_cap[i] = need;
    return e;
}

void HostState::release() {
    if (pipe_ok)
        pipe.destroy();
    pipe_ok = false;
    spec_valid = false;
    for (int i = 0; i < 6; ++i) {
        cudaFree(scratch[i]);
        scratch[i] = nullptr;
        scratch_cap[i] = 0;
    }
    cudaFree(stack_dev);
    stack_dev = nullptr;
    stack_cap = 0;
    valid = false;
    for (cudaEvent_t &e : ev) {
        if (e)
            cudaEventDestroy(e);
        e = nullptr;
    }
}

}  // namespace ptm

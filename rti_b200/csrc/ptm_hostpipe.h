// Internal: moving PAGEABLE host memory to and from the device at PCIe rate, and the per-context state
// of the step-by-step host entry points (ptm_api_host.cu).
//
// A cudaMemcpy from pageable memory runs at ~10 GB/s on the B200 boxes (the driver stages it through
// its own small pinned buffer on one thread); a pinned copy runs at 55 GB/s.  The callers of the
// ptmlib.h drop-in hand over malloc'd buffers (rti-builder/ptm-encoder.c:245, :270-318), so the
// library stages them itself: a ring of pinned chunks per direction, filled / drained by a pool of
// host threads (16 threads copy 59 GB/s here, tools/ubench_h2d.py) while the DMA engine moves the
// previous chunk.  Buffers that already are pinned (cudaHostAlloc / cudaHostRegister) are copied
// directly.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <mutex>
#include <thread>
#include <vector>

struct ptmb_ctx;

namespace ptm {

// memcpy split over persistent worker threads
class CopyPool {
  public:
    explicit CopyPool(int threads);
    ~CopyPool();
    void copy(void *dst, const void *src, size_t bytes);
    // rows x row_bytes, source and destination pitches in bytes
    void copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t rows);
    int threads() const { return (int) workers_.size() + 1; }

  private:
    struct Piece {
        uint8_t *dst;
        const uint8_t *src;
        size_t bytes;
    };
    void run(std::vector<Piece> &pieces);
    void worker();
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::vector<Piece> *job_ = nullptr;
    std::atomic<size_t> next_{0};
    size_t generation_ = 0, active_ = 0;
    bool quit_ = false;
};

struct HostPipe {
    static constexpr int kSlots = 4;
    static constexpr size_t kChunk = 16u << 20;

    int device = 0;
    CopyPool *pool = nullptr;
    cudaStream_t up = nullptr, down = nullptr;     // copy streams: uploads and downloads overlap each other and the kernels
    uint8_t *up_buf[kSlots] = {}, *down_buf[kSlots] = {};
    cudaEvent_t up_done[kSlots] = {}, down_done[kSlots] = {};
    unsigned up_next = 0;
    struct Pending {
        void *dst;
        size_t dpitch, row_bytes, rows;   // rows == 1: plain copy of row_bytes
    } pending[kSlots];
    unsigned down_head = 0, down_count = 0;

    cudaError_t init(int device, int threads /* 0 = all hardware threads, at most 32 */);
    void destroy();
    static bool is_pinned(const void *p);

    // host -> device on `st` (stream ordered); pageable sources are staged, the call returns when the last chunk
    // has been handed to the DMA engine.  2-D form: `rows` rows of row_bytes.
    cudaError_t h2d(void *dev, const void *host, size_t bytes, cudaStream_t st);
    cudaError_t h2d_2d(void *dev, size_t dpitch, const void *host, size_t spitch, size_t row_bytes, size_t rows,
                       cudaStream_t st);
    // device -> host on `st`; pageable destinations are completed lazily: call finish() before the data is used
    cudaError_t d2h(void *host, const void *dev, size_t bytes, cudaStream_t st);
    cudaError_t finish();

  private:
    cudaError_t drain_one();
};

// What the step-by-step entries keep between calls (see ptm_api_host.cu).
struct HostState {
    HostPipe pipe;
    bool pipe_ok = false;
    cudaEvent_t ev[8] = {};
    // device mirror of the caller's stack: byte for byte the host range [base, base + bytes)
    uint8_t *stack_dev = nullptr;
    size_t stack_cap = 0;
    const uint8_t *base = nullptr;
    size_t bytes = 0;            // mirrored bytes
    size_t mirror_image_bytes = 0;
    unsigned mirror_lights = 0;
    unsigned served_fits = 0;    // bit c: channel offset c has been fitted from this mirror
    bool served_avg = false;
    bool valid = false;
    // scratch
    void *scratch[6] = {};
    size_t scratch_cap[6] = {};
    // what ptm_lrgb_finish left on the device for the ptm_scale_coefficients that normally follows (ptm_api_host.cu):
    // its float results (scratch[4]) and their quantised bytes + scale / bias (scratch[5]), for host buffer spec_host
    const float *spec_host = nullptr;
    size_t spec_pixels = 0;
    bool spec_valid = false;
    cudaError_t reserve(int i, size_t bytes);
    void release();
};

// the context's host state with its staging pipe ready, or nullptr if pinned memory could not be had
HostState *host_state(struct ::ptmb_ctx *c);

}  // namespace ptm

// One process, several GPUs: the row-slab sharding of the fit path behind the C ABI.
//
// The reference parallelises ONE encode over the host's cores with `#pragma omp parallel for` over
// image rows (rti-builder/ptmlib.c:704,739,774,829,869); here the same rows are split into one
// contiguous slab per GPU -- rows [floor(i*H/n), floor((i+1)*H/n)) for GPU i, the formula of
// rti_b200/encoder.py:slab_rows -- and the only cross-GPU step, the merge of the 12 extrema keys,
// goes through the mailbox kernels of ptm_xchg.cu / ptm_device.cuh.  Inside one process the
// mailboxes need no IPC handles: with peer access enabled every GPU stores straight into the
// others' mailbox allocations (unified addressing).
//
// Each GPU has a persistent host thread that issues its work (cudaSetDevice is per thread, the
// host-buffer path blocks on its copies, and at 8 GPUs a device-resident fit is 0.13 ms -- issuing
// 8 x 2 launches from one thread would leave the GPUs waiting for the host).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ptm_b200.h"
#include "ptm_internal.h"
#include "ptm_kernels.h"

extern "C" int ptmb_internal_fail(const char *what, const char *detail);
extern "C" int32_t *ptmb_internal_keys(ptmb_ctx_t *c);
extern "C" void ptmb_internal_set_host_threads(ptmb_ctx_t *c, int n);
extern "C" float ptmb_internal_last_copy_ms(ptmb_ctx_t *c);
extern "C" int ptmb_internal_encode_host_fit(ptmb_ctx_t *c, const void *stack_host, size_t host_image_bytes,
                                             int sample_type, size_t width, size_t height, unsigned n_lights,
                                             const float *M, int format_planes, int32_t *keys_dev);
extern "C" int ptmb_internal_encode_host_finish(ptmb_ctx_t *c, const int32_t *keys_dev, uint8_t *const *blocks_host,
                                                float *coeffs_host, size_t coeffs_plane_floats,
                                                ptmb_scale_bias_t *sb_host);

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool pending = false, quit = false;
    int rc = 0;
    std::string err;
};

}  // namespace

struct ptmb_multi {
    int n = 0;
    std::vector<int> device;
    std::vector<ptmb_ctx_t *> ctx;
    std::vector<ptmb_xchg *> xchg;
    std::vector<cudaStream_t> stream;
    std::vector<Worker *> worker;
    bool peer_mailboxes = false;
    // share of an image's rows each GPU takes in the HOST-buffer encode (equal until calibrated): the slabs
    // arrive over PCIe, and the GPUs of one box do not all get the same host bandwidth when they copy at once
    std::vector<double> weight;
    bool calibrated = false;
    bool explicit_weights = false;   // set by ptmb_multi_set_host_weights: never adjusted automatically
    int refinements = 0;             // automatic adjustments made from the copy times of real encodes
    int host_encodes = 0;            // large host encodes since the last calibration (the first one also pays for
                                     // allocations and first launches: its copy times are not used)

    // rows [cut(i), cut(i+1)) belong to GPU i
    size_t cut(size_t height, int i) const {
        if (i <= 0)
            return 0;
        if (i >= n)
            return height;
        double total = 0, upto = 0;
        for (int j = 0; j < n; ++j) {
            total += weight[j];
            if (j < i)
                upto += weight[j];
        }
        const size_t r = (size_t) ((double) height * (upto / total) + 0.5);
        return std::min(r, height);
    }

    // run fn(i) on every GPU's thread; returns the first failure (message moved to the caller's thread)
    int run_all(const std::function<int(int)> &fn) {
        for (int i = 0; i < n; ++i) {
            Worker &w = *worker[i];
            std::lock_guard<std::mutex> lk(w.mu);
            w.job = [i, &fn] { return fn(i); };
            w.pending = true;
            w.cv.notify_all();
        }
        int rc = 0;
        for (int i = 0; i < n; ++i) {
            Worker &w = *worker[i];
            std::unique_lock<std::mutex> lk(w.mu);
            w.cv.wait(lk, [&w] { return !w.pending; });
            if (w.rc && !rc) {
                rc = w.rc;
                ptmb_internal_fail("GPU worker", w.err.c_str());
            }
        }
        return rc;
    }
};

namespace {

void worker_main(Worker *w) {
    for (;;) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [w] { return w->pending || w->quit; });
        if (w->quit)
            return;
        std::function<int()> job = w->job;
        lk.unlock();
        const int rc = job();
        std::string err = rc ? ptmb_last_error() : "";
        lk.lock();
        w->rc = rc;
        w->err.swap(err);
        w->pending = false;
        w->cv.notify_all();
    }
}

}  // namespace

extern "C" {

void ptmb_multi_destroy(ptmb_multi_t *m) {
    if (!m)
        return;
    for (Worker *w : m->worker) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable())
            w->th.join();
        delete w;
    }
    for (int i = 0; i < m->n; ++i) {
        if (i < (int) m->ctx.size() && m->ctx[i])
            ptmb_synchronize(m->ctx[i]);
    }
    for (int i = 0; i < (int) m->xchg.size(); ++i) {
        if (m->xchg[i]) {
            cudaSetDevice(m->device[i]);
            cudaFree(m->xchg[i]->local);
            delete m->xchg[i];
        }
    }
    for (int i = 0; i < (int) m->ctx.size(); ++i) {
        if (m->ctx[i])
            ptmb_destroy(m->ctx[i]);
        if (i < (int) m->stream.size() && m->stream[i]) {
            cudaSetDevice(m->device[i]);
            cudaStreamDestroy(m->stream[i]);
        }
    }
    delete m;
}

int ptmb_multi_create(const int *devices, int n_dev, ptmb_multi_t **out) {
    if (!out)
        return ptmb_internal_fail("ptmb_multi_create", "out pointer is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return ptmb_internal_fail("ptmb_multi_create", "no CUDA device (this library has no CPU path)");
    if (n_dev <= 0)
        n_dev = count;   // all visible GPUs
    if (n_dev > ptm::kMaxRanks)
        return ptmb_internal_fail("ptmb_multi_create", "at most 16 GPUs");
    ptmb_multi *m = new ptmb_multi;
    m->n = n_dev;
    for (int i = 0; i < n_dev; ++i) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= count) {
            ptmb_multi_destroy(m);
            return ptmb_internal_fail("ptmb_multi_create", "device index out of range");
        }
        for (int j = 0; j < i; ++j)
            if (m->device[j] == d) {
                ptmb_multi_destroy(m);
                return ptmb_internal_fail("ptmb_multi_create", "a device is listed twice");
            }
        m->device.push_back(d);
    }
    m->weight.assign(n_dev, 1.0);
    m->ctx.assign(n_dev, nullptr);
    m->xchg.assign(n_dev, nullptr);
    m->stream.assign(n_dev, nullptr);
    for (int i = 0; i < n_dev; ++i) {
        cudaError_t e = cudaSetDevice(m->device[i]);
        if (e == cudaSuccess)
            e = cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking);
        if (e != cudaSuccess || ptmb_create(m->device[i], m->stream[i], &m->ctx[i])) {
            if (e != cudaSuccess)
                ptmb_internal_fail("ptmb_multi_create", cudaGetErrorString(e));
            ptmb_multi_destroy(m);
            return 1;
        }
    }
    // pageable host buffers are staged by host threads: share the cores between the GPUs
    {
        const int hw = (int) std::thread::hardware_concurrency();
        for (int i = 0; i < n_dev; ++i)
            ptmb_internal_set_host_threads(m->ctx[i], std::max(2, hw / n_dev));
    }
    // peer access between every pair
    bool peers = n_dev > 1;
    for (int i = 0; i < n_dev && peers; ++i)
        for (int j = 0; j < n_dev && peers; ++j) {
            if (i == j)
                continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->device[i], m->device[j]) != cudaSuccess || !can)
                peers = false;
        }
    if (peers) {
        for (int i = 0; i < n_dev; ++i) {
            cudaSetDevice(m->device[i]);
            for (int j = 0; j < n_dev; ++j) {
                if (i == j)
                    continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(m->device[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled)
                    cudaGetLastError();   // clear
                else if (e != cudaSuccess)
                    peers = false;
            }
        }
    }
    if (n_dev > 1 && !peers) {
        ptmb_multi_destroy(m);
        return ptmb_internal_fail("ptmb_multi_create", "the GPUs cannot map each other's memory (no peer access)");
    }
    if (n_dev > 1) {
        for (int i = 0; i < n_dev; ++i) {
            ptmb_xchg *x = new ptmb_xchg;
            x->device = m->device[i];
            x->rank = i;
            x->world = n_dev;
            m->xchg[i] = x;
            cudaSetDevice(m->device[i]);
            cudaError_t e = cudaMalloc(&x->local, ptm::XchgLayout::bytes);
            if (e == cudaSuccess)
                e = cudaMemset(x->local, 0, ptm::XchgLayout::bytes);
            if (e == cudaSuccess)
                e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                ptmb_internal_fail("ptmb_multi_create", cudaGetErrorString(e));
                ptmb_multi_destroy(m);
                return 1;
            }
        }
        for (int i = 0; i < n_dev; ++i)
            for (int r = 0; r < n_dev; ++r)
                m->xchg[i]->mailbox[r] = m->xchg[r]->local;   // plain peer pointers (same process)
        m->peer_mailboxes = true;
    }
    for (int i = 0; i < n_dev; ++i) {
        Worker *w = new Worker;
        m->worker.push_back(w);
        w->th = std::thread(worker_main, w);
    }
    // every worker binds its GPU once
    if (m->run_all([m](int i) {
            return cudaSetDevice(m->device[i]) == cudaSuccess ? 0 : ptmb_internal_fail("cudaSetDevice", "failed");
        })) {
        ptmb_multi_destroy(m);
        return 1;
    }
    *out = m;
    return 0;
}

int ptmb_multi_device_count(const ptmb_multi_t *m) { return m ? m->n : 0; }

ptmb_ctx_t *ptmb_multi_ctx(ptmb_multi_t *m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }

int ptmb_multi_slab(const ptmb_multi_t *m, size_t height, int i, size_t *row0, size_t *row1) {
    if (!m || i < 0 || i >= m->n || !row0 || !row1)
        return ptmb_internal_fail("ptmb_multi_slab", "bad argument");
    *row0 = (size_t) i * height / (size_t) m->n;
    *row1 = (size_t) (i + 1) * height / (size_t) m->n;
    return 0;
}

int ptmb_multi_host_slab(const ptmb_multi_t *m, size_t height, int i, size_t *row0, size_t *row1) {
    if (!m || i < 0 || i >= m->n || !row0 || !row1)
        return ptmb_internal_fail("ptmb_multi_host_slab", "bad argument");
    *row0 = m->cut(height, i);
    *row1 = m->cut(height, i + 1);
    return 0;
}

int ptmb_multi_set_host_weights(ptmb_multi_t *m, const double *weights) {
    if (!m)
        return ptmb_internal_fail("ptmb_multi_set_host_weights", "NULL argument");
    for (int i = 0; i < m->n; ++i) {
        const double w = weights ? weights[i] : 1.0;
        if (!(w > 0) || !(w < 1e300))
            return ptmb_internal_fail("ptmb_multi_set_host_weights", "weights must be positive and finite");
    }
    for (int i = 0; i < m->n; ++i)
        m->weight[i] = weights ? weights[i] : 1.0;
    m->calibrated = true;   // an explicit choice is not overwritten by the automatic probe
    m->explicit_weights = weights != nullptr;
    m->refinements = 0;
    return 0;
}

// Every GPU copies from its own pinned probe buffer for the same span of wall-clock time, all at once; the bytes
// each one moved in that span are its share of the host's PCIe bandwidth under the load pattern of
// ptmb_multi_encode_host (on the 8-GPU boxes of this pool four GPUs get 23 GB/s and four 35 GB/s).
int ptmb_multi_calibrate_h2d(ptmb_multi_t *m, double seconds, double *gbs_out) {
    if (!m)
        return ptmb_internal_fail("ptmb_multi_calibrate_h2d", "NULL argument");
    if (!(seconds > 0))
        seconds = 0.012;
    const size_t chunk = 8u << 20;
    const int n = m->n;
    std::vector<double> rate(n, 0.0);
    std::atomic<int> ready(0);
    const int rc = m->run_all([&, m](int i) -> int {
        void *host = nullptr, *dev = nullptr;
        cudaStream_t st = m->stream[i];
        cudaError_t e = cudaHostAlloc(&host, 2 * chunk, cudaHostAllocPortable);
        if (e == cudaSuccess)
            e = cudaMalloc(&dev, 2 * chunk);
        if (e == cudaSuccess) {
            std::memset(host, 0x5a, 2 * chunk);
            e = cudaMemcpyAsync(dev, host, 2 * chunk, cudaMemcpyHostToDevice, st);   // first touch / mapping
        }
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
        // all GPUs start together, also when one of them failed to set up
        ready.fetch_add(1);
        while (ready.load() < n)
            std::this_thread::yield();
        size_t moved = 0;
        const auto t0 = std::chrono::steady_clock::now();
        double dt = 0;
        if (e == cudaSuccess) {
            // two chunk copies in flight: the link never idles while the host thread turns around
            cudaEvent_t ev[2];
            cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
            int k = 0;
            cudaMemcpyAsync(dev, host, chunk, cudaMemcpyHostToDevice, st);
            cudaEventRecord(ev[0], st);
            for (;;) {
                k ^= 1;
                cudaMemcpyAsync((char *) dev + k * chunk, (char *) host + k * chunk, chunk, cudaMemcpyHostToDevice, st);
                cudaEventRecord(ev[k], st);
                e = cudaEventSynchronize(ev[k ^ 1]);
                moved += chunk;
                dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (e != cudaSuccess || dt >= seconds)
                    break;
            }
            cudaStreamSynchronize(st);
            cudaEventDestroy(ev[0]);
            cudaEventDestroy(ev[1]);
        }
        if (dev)
            cudaFree(dev);
        if (host)
            cudaFreeHost(host);
        if (e != cudaSuccess)
            return ptmb_internal_fail("ptmb_multi_calibrate_h2d", cudaGetErrorString(e));
        rate[i] = (double) moved / dt;
        return 0;
    });
    if (rc)
        return rc;
    for (int i = 0; i < n; ++i) {
        m->weight[i] = rate[i];
        if (gbs_out)
            gbs_out[i] = rate[i] / 1e9;
    }
    m->calibrated = true;
    m->explicit_weights = false;
    m->refinements = 0;
    m->host_encodes = 0;
    return 0;
}

int ptmb_multi_set_matrix(ptmb_multi_t *m, const float *M, unsigned n_lights) {
    if (!m)
        return ptmb_internal_fail("ptmb_multi_set_matrix", "NULL argument");
    return m->run_all([=](int i) { return ptmb_set_matrix(m->ctx[i], M, n_lights); });
}

int ptmb_multi_set_fit_mode(ptmb_multi_t *m, int flags) {
    if (!m)
        return ptmb_internal_fail("ptmb_multi_set_fit_mode", "NULL argument");
    for (int i = 0; i < m->n; ++i)
        if (ptmb_set_fit_mode(m->ctx[i], flags))
            return 1;
    return 0;
}

int ptmb_multi_synchronize(ptmb_multi_t *m) {
    if (!m)
        return ptmb_internal_fail("ptmb_multi_synchronize", "NULL argument");
    return m->run_all([=](int i) { return ptmb_synchronize(m->ctx[i]); });
}

// Whole encode from ONE host stack across the GPUs (same contract as ptmb_encode_host).  GPUs whose slab is
// empty (more GPUs than rows) still take part in the key exchange with the initial extrema.
int ptmb_multi_encode_host(ptmb_multi_t *m, const void *stack_host, int sample_type, size_t width, size_t height,
                           unsigned n_lights, const float *M, int format_planes, uint8_t *const *blocks_host,
                           float *coeffs_host, ptmb_scale_bias_t *sb_host) {
    if (!m || !stack_host || !M || !blocks_host || !sb_host)
        return ptmb_internal_fail("ptmb_multi_encode_host", "NULL argument");
    if (format_planes != 1 && format_planes != 3)
        return ptmb_internal_fail("ptmb_multi_encode_host", "format_planes must be 1 (LRGB) or 3 (RGB)");
    const size_t ss = sample_type == PTMB_U8 ? 1 : sample_type == PTMB_U16 ? 2 : 0;
    if (!ss || width * height == 0)
        return ptmb_internal_fail("ptmb_multi_encode_host", "uint8 or uint16 stacks of at least one pixel");
    if (m->n == 1)
        return ptmb_encode_host(m->ctx[0], stack_host, sample_type, width, height, n_lights, M, format_planes,
                                blocks_host, coeffs_host, sb_host);
    const size_t image_bytes = width * height * 3 * ss;
    const size_t plane_floats = width * height * PTMB_COEFFICIENTS;
    // the first large host encode measures every GPU's share of the host bandwidth once (PTM_MULTI_BALANCE=0:
    // equal slabs); results do not depend on the split (tests/test_gpu_multi.py)
    if (!m->calibrated && image_bytes * n_lights >= ((size_t) 256 << 20)) {
        const char *env = getenv("PTM_MULTI_BALANCE");
        if (!(env && env[0] == '0') && ptmb_multi_calibrate_h2d(m, 0.012, nullptr)) {
            // the probe is an optimisation: without it (no pinned memory to be had, ...) the slabs stay equal
            cudaGetLastError();
            m->weight.assign(m->n, 1.0);
        }
        m->calibrated = true;
    }
    std::vector<ptmb_scale_bias_t> sb(m->n);
    std::vector<double> copy_ms(m->n, -1.0);
    const int rc = m->run_all([&, m](int i) -> int {
        const size_t r0 = m->cut(height, i), r1 = m->cut(height, i + 1);
        const size_t rows = r1 - r0, px0 = r0 * width;
        ptmb_ctx_t *c = m->ctx[i];
        int32_t *keys = ptmb_internal_keys(c);
        uint8_t *blk[3] = {nullptr, nullptr, nullptr};
        for (int b = 0; b < format_planes; ++b)
            blk[b] = blocks_host[b] + px0 * PTMB_COEFFICIENTS;
        if (format_planes == 1)
            blk[1] = blocks_host[1] + px0 * 3;
        if (rows == 0) {
            if (ptmb_set_matrix(c, M, n_lights) || ptmb_keys_init(c, keys))
                return 1;
            return ptmb_xchg_allreduce_max(c, m->xchg[i], keys);
        }
        if (format_planes == 1 && ptmb_encode_host_early_rgb(c, blk[1]))
            return 1;
        if (ptmb_internal_encode_host_fit(c, (const uint8_t *) stack_host + px0 * 3 * ss, image_bytes, sample_type,
                                          width, rows, n_lights, M, format_planes, nullptr))
            return 1;
        if (ptmb_xchg_allreduce_max(c, m->xchg[i], keys))
            return 1;
        if (ptmb_internal_encode_host_finish(c, nullptr, blk, coeffs_host ? coeffs_host + px0 * PTMB_COEFFICIENTS : nullptr,
                                             plane_floats, &sb[i]))
            return 1;
        copy_ms[i] = ptmb_internal_last_copy_ms(c);   // on the GPU's own thread (its device is current)
        return 0;
    });
    if (rc)
        return rc;
    // The probe's small copies do not load the links exactly like slab-sized ones: while the split was chosen
    // automatically, the copy times of this encode (first to last input byte per GPU) correct it for the next
    // one when the GPUs finished more than 6 % apart (at most 3 corrections, never from the first encode).
    const bool large = image_bytes * n_lights >= ((size_t) 256 << 20);
    if (large)
        ++m->host_encodes;
    static const bool debug = getenv("PTM_MULTI_DEBUG") != nullptr;
    if (debug) {
        fprintf(stderr, "ptmb_multi_encode_host #%d: rows / copy ms per GPU:", m->host_encodes);
        for (int i = 0; i < m->n; ++i)
            fprintf(stderr, " %zu/%.2f", m->cut(height, i + 1) - m->cut(height, i), copy_ms[i]);
        fprintf(stderr, "  (refinements so far %d)\n", m->refinements);
    }
    if (m->calibrated && !m->explicit_weights && m->refinements < 3 && large && m->host_encodes >= 2) {
        const std::vector<double> &ms = copy_ms;
        double lo = 1e300, hi = 0;
        bool ok = true;
        for (int i = 0; i < m->n && ok; ++i) {
            const size_t rows = m->cut(height, i + 1) - m->cut(height, i);
            ok = rows > 0 && ms[i] > 0;
            lo = std::min(lo, ms[i]);
            hi = std::max(hi, ms[i]);
        }
        if (ok && hi > 1.06 * lo) {
            // damped (half the indicated correction, each share changed by at most 25 %): the copy times carry a few
            // per cent of noise, and the links are coupled -- what one GPU gives up the others do not simply gain
            double mean = 0;
            for (int i = 0; i < m->n; ++i)
                mean += ms[i] / m->n;
            std::vector<double> w(m->n);
            for (int i = 0; i < m->n; ++i) {
                const double corr = std::min(1.25, std::max(0.8, std::sqrt(mean / ms[i])));
                w[i] = (double) (m->cut(height, i + 1) - m->cut(height, i)) * corr;
            }
            m->weight = w;
            ++m->refinements;
        }
    }
    // every GPU derived scale / bias from the same merged keys; report those of the first non-empty slab
    for (int i = 0; i < m->n; ++i) {
        const size_t r0 = m->cut(height, i), r1 = m->cut(height, i + 1);
        if (r1 > r0) {
            *sb_host = sb[i];
            break;
        }
    }
    int status = 0;
    if (ptmb_xchg_status(m->ctx[0], m->xchg[0], &status))
        return 1;
    if (status)
        return ptmb_internal_fail("ptmb_multi_encode_host", "a GPU did not deliver its extrema in time");
    return 0;
}

// Device-resident LRGB encode of one slab per GPU, two launches each, keys handed over NVLink by K1's last
// block (ptmb_encode_lrgb_dev with the in-process mailboxes).  Arrays have one entry per GPU.  Asynchronous:
// returns when every GPU's work is enqueued on its stream (ptmb_multi_synchronize waits for the results).
int ptmb_multi_encode_lrgb_dev(ptmb_multi_t *m, const void *const *stack_dev, int sample_type,
                               const size_t *image_stride, const size_t *pixels, float *const *coeffs_dev,
                               uint8_t *const *rgb_dev, uint8_t *const *coef_bytes_dev,
                               ptmb_scale_bias_t *const *sb_dev) {
    if (!m || !stack_dev || !image_stride || !pixels || !coeffs_dev || !rgb_dev || !coef_bytes_dev)
        return ptmb_internal_fail("ptmb_multi_encode_lrgb_dev", "NULL argument");
    return m->run_all([=](int i) {
        return ptmb_encode_lrgb_dev(m->ctx[i], m->xchg[i], stack_dev[i], sample_type, image_stride[i], pixels[i],
                                    coeffs_dev[i], rgb_dev[i], coef_bytes_dev[i], sb_dev ? sb_dev[i] : nullptr);
    });
}

}  // extern "C"

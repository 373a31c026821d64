// Host-buffer forms of the individual ptmlib.h steps (include/ptm_b200.h, "host" section).
// These exist so that ptmlib.h callers that keep the reference's call sequence
// (fit, then average, then scale: rti-builder/ptm-encoder.c:316-325,406) still run every
// step on the GPU.  Each call streams its input through device slabs; they are copy bound
// by construction -- the fused ptmb_encode_host_* path is the fast one.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/ptm_b200.h"
#include "ptm_hostpipe.h"
#include "ptm_kernels.h"

namespace {

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

size_t ssize(int sample_type) {
    return sample_type == PTMB_U8 ? 1 : sample_type == PTMB_U16 ? 2 : sample_type == PTMB_U32 ? 4 : 0;
}

}  // namespace

// ptm_api.cu
extern "C" int ptmb_internal_fail(const char *what, const char *detail);
extern "C" int ptmb_internal_device(const ptmb_ctx_t *c);
extern "C" void *ptmb_internal_stream(const ptmb_ctx_t *c);
extern "C" const float *ptmb_internal_matrix(const ptmb_ctx_t *c);
extern "C" ptm::HostState **ptmb_internal_host(ptmb_ctx_t *c);
extern "C" int ptmb_internal_host_threads(const ptmb_ctx_t *c);

#define CKH(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess)                                             \
            return ptmb_internal_fail(#call, cudaGetErrorString(e_));      \
    } while (0)

#define REQH(cond, msg)                                                    \
    do {                                                                   \
        if (!(cond))                                                       \
            return ptmb_internal_fail("invalid argument", msg);            \
    } while (0)

extern "C" {

// ---------------------------------------------------------------------------------------------
// Per-context state of these entries: the staging pipe, scratch that survives between calls, and
// a DEVICE MIRROR of the caller's stack.  The reference's sequence touches the stack two to four
// times (ptm_fit_poly_jsample per channel, then ptm_cbcr_avg: rti-builder/ptm-encoder.c:272-325);
// crossing PCIe once is enough.  The mirror is byte for byte the host range the first fit read,
// so a later call whose pointer lies inside it (same stack, another channel offset, or the
// averaging pass) is served from the device.  It is NOT a cache of results and it is never
// trusted across encodes: a fit for a channel offset that has already been served from this
// mirror, or a second averaging pass, means "the caller starts over" and uploads again; so does
// any other pointer, size or light count; ptm_scale_coefficients ends the sequence.  What is
// assumed: the caller does not modify the stack BETWEEN the calls of one sequence (the reference's
// encoder does not).  PTM_DROPIN_MIRROR=0 turns the mirror off (every call uploads).
// ---------------------------------------------------------------------------------------------
}  // extern "C"

// Bitwise comparison of two float planes (16 bytes per thread and step); *flag becomes non-zero on any difference.
__global__ void planes_differ_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, size_t n16, int *flag) {
    bool differ = false;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x) {
        const uint4 x = a[i], y = b[i];
        differ |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    if (differ)
        *flag = 1;
}

ptm::HostState *ptm::host_state(ptmb_ctx_t *c) {
    ptm::HostState *&h = *ptmb_internal_host(c);
    if (!h)
        h = new ptm::HostState;
    if (!h->pipe_ok) {
        if (h->pipe.init(ptmb_internal_device(c), ptmb_internal_host_threads(c)) != cudaSuccess) {
            h->pipe.destroy();
            return nullptr;
        }
        h->pipe_ok = true;
        for (cudaEvent_t &e : h->ev)
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
    return h;
}

extern "C" {

extern "C" void ptmb_internal_host_release(ptmb_ctx_t *c) {
    ptm::HostState *&h = *ptmb_internal_host(c);
    if (h) {
        h->release();
        delete h;
        h = nullptr;
    }
}

static bool mirror_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *v = getenv("PTM_DROPIN_MIRROR");
        on = !(v && v[0] == '0');
    }
    return on != 0;
}

#define HS(c)                                                                           \
    ptm::HostState *h = ptm::host_state(c);                                             \
    if (!h)                                                                             \
        return ptmb_internal_fail("host staging", "cannot allocate the pinned staging ring")

// Row slabs of about 8 MiB per light: small enough that the kernels start early, large enough for PCIe.
static size_t slab_rows_for(size_t width, size_t bytes_per_pixel, size_t height) {
    size_t r = std::max<size_t>(1, (8u << 20) / std::max<size_t>(1, width * bytes_per_pixel));
    return std::min(r, height);
}

// ptm_fit_poly_jsample / ptm_fit_poly_uint with host pointers (rti-builder/ptmlib.c:692-760).
// samples_host: first sample of the channel; image l, pixel p sits at
// samples_host[(l * pixels + p) * pixel_stride] exactly as in the reference (:701-702,710).
int ptmb_fit_channel_host(ptmb_ctx_t *c, const void *samples_host, int sample_type, size_t pixel_stride,
                          size_t width, size_t height, unsigned n_lights, const float *M, float *coeffs_host) {
    REQH(c && samples_host && M && coeffs_host, "NULL argument");
    const size_t ss = ssize(sample_type);
    REQH(ss != 0 && pixel_stride >= 1, "bad sample type or stride");
    if (ptmb_set_matrix(c, M, n_lights))
        return 1;
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    HS(c);
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t pixels = width * height;
    if (pixels == 0)
        return 0;
    h->spec_valid = false;
    const size_t image_bytes = pixels * pixel_stride * ss;   // reference: height * width * pixel_stride samples
    // bytes this call needs: up to the wanted channel's last sample (the trailing samples of the last pixel
    // of the last image may lie outside the caller's buffer when it was offset by a channel)
    const size_t need = (size_t) (n_lights - 1) * image_bytes + ((pixels - 1) * pixel_stride + 1) * ss;
    const uint8_t *src = (const uint8_t *) samples_host;

    // ---- mirror look-up
    bool hit = false;
    size_t delta = 0;
    if (mirror_enabled() && h->valid && src >= h->base && (size_t) (src - h->base) < pixel_stride * ss &&
        (src - h->base) % ss == 0 && h->bytes + pixel_stride * ss > need + (size_t) (src - h->base)) {
        delta = (size_t) (src - h->base);
        const unsigned ch = (unsigned) (delta / ss);
        hit = ch < 32 && !(h->served_fits & (1u << ch)) && h->mirror_image_bytes == image_bytes &&
              h->mirror_lights == n_lights;
    }
    if (!hit) {
        h->valid = false;
        if (h->stack_cap < need + 64) {
            CKH(cudaDeviceSynchronize());
            cudaFree(h->stack_dev);
            h->stack_dev = nullptr;
            h->stack_cap = 0;
            if (cudaMalloc((void **) &h->stack_dev, need + 64) != cudaSuccess) {
                cudaGetLastError();
                return ptmb_internal_fail("ptmb_fit_channel_host", "not enough device memory for the stack");
            }
            h->stack_cap = need + 64;
        }
        delta = 0;
    } else if (delta + need > h->bytes) {
        // the same stack through a later channel offset: the few bytes past the mirrored range
        CKH(h->pipe.h2d(h->stack_dev + h->bytes, h->base + h->bytes, delta + need - h->bytes, st));
        h->bytes = delta + need;
    }

    // ---- slab pipeline: upload (copy stream) -> fit (context stream) -> download (second copy stream)
    const size_t slab_rows = slab_rows_for(width, pixel_stride * ss, height);
    const size_t slab_px = slab_rows * width;
    CKH(h->reserve(0, 2 * slab_px * PTMB_COEFFICIENTS * sizeof(float)));
    float *out[2] = {(float *) h->scratch[0], (float *) h->scratch[0] + slab_px * PTMB_COEFFICIENTS};
    cudaEvent_t up_ev = h->ev[0], fit_ev[2] = {h->ev[1], h->ev[2]}, down_ev[2] = {h->ev[3], h->ev[4]};
    if (!hit)
        CKH(cudaStreamWaitEvent(h->pipe.up, h->ev[7], 0));   // earlier kernels that read the mirror (see the end)
    int buf = 0;
    for (size_t row = 0; row < height; row += slab_rows, buf ^= 1) {
        const size_t rows = std::min(slab_rows, height - row);
        const size_t npx = rows * width, off = row * width * pixel_stride * ss;
        if (!hit) {
            // n_lights pieces of this slab; the very last one stops at the last needed sample
            const size_t piece = npx * pixel_stride * ss;
            const bool last_slab = row + rows == height;
            const size_t full_rows = last_slab ? n_lights - 1 : n_lights;
            CKH(h->pipe.h2d_2d(h->stack_dev + off, image_bytes, src + off, image_bytes, piece, full_rows, h->pipe.up));
            if (last_slab)
                CKH(h->pipe.h2d(h->stack_dev + off + (size_t) (n_lights - 1) * image_bytes,
                                src + off + (size_t) (n_lights - 1) * image_bytes,
                                ((npx - 1) * pixel_stride + 1) * ss, h->pipe.up));
            CKH(cudaEventRecord(up_ev, h->pipe.up));
            CKH(cudaStreamWaitEvent(st, up_ev, 0));
        }
        CKH(cudaStreamWaitEvent(st, down_ev[buf], 0));      // the download that last read this output buffer
        if (ptmb_fit_channel(c, h->stack_dev + delta + off, sample_type, pixel_stride, pixels * pixel_stride, npx, out[buf]))
            return 1;
        CKH(cudaEventRecord(fit_ev[buf], st));
        CKH(cudaStreamWaitEvent(h->pipe.down, fit_ev[buf], 0));
        CKH(h->pipe.d2h(coeffs_host + row * width * PTMB_COEFFICIENTS, out[buf], npx * PTMB_COEFFICIENTS * sizeof(float),
                        h->pipe.down));
        CKH(cudaEventRecord(down_ev[buf], h->pipe.down));
    }
    CKH(cudaEventRecord(h->ev[7], st));
    CKH(h->pipe.finish());
    CKH(cudaStreamSynchronize(h->pipe.down));
    CKH(cudaStreamSynchronize(st));
    if (!hit) {
        h->base = src;
        h->bytes = need;
        h->mirror_image_bytes = image_bytes;
        h->mirror_lights = n_lights;
        h->served_fits = 0;
        h->served_avg = false;
        h->valid = mirror_enabled();
    }
    h->served_fits |= 1u << (unsigned) (delta / ss);
    return 0;
}

// ptm_cbcr_avg with host pointers (rti-builder/ptmlib.c:762-802).
int ptmb_chroma_avg_host(ptmb_ctx_t *c, const uint8_t *stack_host, size_t pixels, unsigned n_lights,
                         uint8_t *avg_host) {
    REQH(c && stack_host && avg_host && n_lights >= 1, "NULL argument");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    HS(c);
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    if (pixels == 0)
        return 0;
    const size_t image_bytes = pixels * 3, need = image_bytes * n_lights;
    const bool hit = mirror_enabled() && h->valid && stack_host == h->base && !h->served_avg &&
                     h->mirror_image_bytes == image_bytes && h->mirror_lights == n_lights && h->bytes + 3 > need;
    CKH(h->reserve(1, pixels * 3));
    uint8_t *avg_dev = (uint8_t *) h->scratch[1];
    if (hit) {
        if (h->bytes < need) {   // the last pixel's chroma bytes that a channel-0 fit did not need
            CKH(h->pipe.h2d(h->stack_dev + h->bytes, h->base + h->bytes, need - h->bytes, st));
            h->bytes = need;
        }
        if (ptmb_chroma_avg(c, h->stack_dev, image_bytes, pixels, n_lights, avg_dev))
            return 1;
        CKH(cudaEventRecord(h->ev[7], st));
        CKH(h->pipe.d2h(avg_host, avg_dev, pixels * 3, st));
        CKH(h->pipe.finish());
        CKH(cudaStreamSynchronize(st));
        h->served_avg = true;
        return 0;
    }
    // no mirror: stream the stack through two slabs
    h->valid = false;
    const size_t slab = std::min<size_t>(pixels, (8u << 20) / 3);
    CKH(h->reserve(2, 2 * slab * 3 * (size_t) n_lights));
    uint8_t *stage[2] = {(uint8_t *) h->scratch[2], (uint8_t *) h->scratch[2] + slab * 3 * n_lights};
    cudaEvent_t up_ev = h->ev[0], used[2] = {h->ev[1], h->ev[2]};
    int buf = 0;
    for (size_t p0 = 0; p0 < pixels; p0 += slab, buf ^= 1) {
        const size_t npx = std::min(slab, pixels - p0);
        CKH(cudaStreamWaitEvent(h->pipe.up, used[buf], 0));
        CKH(h->pipe.h2d_2d(stage[buf], slab * 3, stack_host + p0 * 3, image_bytes, npx * 3, n_lights, h->pipe.up));
        CKH(cudaEventRecord(up_ev, h->pipe.up));
        CKH(cudaStreamWaitEvent(st, up_ev, 0));
        if (ptmb_chroma_avg(c, stage[buf], slab * 3, npx, n_lights, avg_dev + p0 * 3))
            return 1;
        CKH(cudaEventRecord(used[buf], st));
    }
    CKH(h->pipe.d2h(avg_host, avg_dev, pixels * 3, st));
    CKH(h->pipe.finish());
    CKH(cudaStreamSynchronize(st));
    return 0;
}

// The encoder's LRGB colour + normalise loop with host pointers (rti-builder/ptm-encoder.c:327-353).
// Both operands are read AND written by the caller's contract, so they travel both ways; uploads and
// downloads of neighbouring slabs overlap (PCIe is full duplex).
int ptmb_lrgb_colour_normalise_host(ptmb_ctx_t *c, uint8_t *ycc_to_rgb_host, float *coeffs_host, size_t pixels) {
    REQH(c && ycc_to_rgb_host && coeffs_host, "NULL argument");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    HS(c);
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    // 2 M pixels (54 MB) per slab.  Shorter slabs do not help: 512 K-pixel slabs took 17.1 instead of 15.7 ms for 24 MP
    // from registered memory (29.9 instead of 21.7 ms from pageable memory) -- the call is bound by PCIe running
    // both ways at once (648 MB each way), not by the fill and drain of the pipeline
    const size_t slab = std::min<size_t>(pixels, 2u << 20);
    CKH(h->reserve(0, 2 * slab * PTMB_COEFFICIENTS * sizeof(float)));
    CKH(h->reserve(1, 2 * slab * 3));
    float *cf[2] = {(float *) h->scratch[0], (float *) h->scratch[0] + slab * PTMB_COEFFICIENTS};
    uint8_t *ycc[2] = {(uint8_t *) h->scratch[1], (uint8_t *) h->scratch[1] + slab * 3};
    cudaEvent_t up_ev = h->ev[0], done[2] = {h->ev[1], h->ev[2]}, down_ev[2] = {h->ev[3], h->ev[4]};
    // The reference's encoder hands these very floats to ptm_scale_coefficients next (rti-builder/ptm-encoder.c:406).
    // A copy of the results stays on the device and is quantised right away, so that call only has to VERIFY that the
    // host buffer still holds them while the bytes already travel back (ptmb_scale_host).  Best effort: without the
    // device memory for it the sequence simply runs as before.
    h->spec_valid = false;
    const size_t plane_bytes = pixels * PTMB_COEFFICIENTS * sizeof(float);
    const bool keep = pixels > 0 && h->reserve(4, plane_bytes) == cudaSuccess &&
                      h->reserve(5, pixels * PTMB_COEFFICIENTS + 256) == cudaSuccess;
    if (!keep)
        cudaGetLastError();
    int buf = 0;
    for (size_t p0 = 0; p0 < pixels; p0 += slab, buf ^= 1) {
        const size_t npx = std::min(slab, pixels - p0);
        // with the kept plane the floats are converted where they will stay (no second copy, and no wait for a
        // staging buffer to come free); otherwise through the two staging slabs
        float *cfp = keep ? (float *) h->scratch[4] + p0 * PTMB_COEFFICIENTS : cf[buf];
        CKH(cudaStreamWaitEvent(h->pipe.up, down_ev[buf], 0));   // the slab that used these buffers has left
        CKH(h->pipe.h2d(ycc[buf], ycc_to_rgb_host + p0 * 3, npx * 3, h->pipe.up));
        CKH(h->pipe.h2d(cfp, coeffs_host + p0 * PTMB_COEFFICIENTS, npx * PTMB_COEFFICIENTS * sizeof(float), h->pipe.up));
        CKH(cudaEventRecord(up_ev, h->pipe.up));
        CKH(cudaStreamWaitEvent(st, up_ev, 0));
        if (ptmb_lrgb_colour_normalise(c, ycc[buf], cfp, npx))
            return 1;
        CKH(cudaEventRecord(done[buf], st));
        CKH(cudaStreamWaitEvent(h->pipe.down, done[buf], 0));
        CKH(h->pipe.d2h(ycc_to_rgb_host + p0 * 3, ycc[buf], npx * 3, h->pipe.down));
        CKH(h->pipe.d2h(coeffs_host + p0 * PTMB_COEFFICIENTS, cfp, npx * PTMB_COEFFICIENTS * sizeof(float), h->pipe.down));
        CKH(cudaEventRecord(down_ev[buf], h->pipe.down));
    }
    CKH(h->pipe.finish());
    CKH(cudaStreamSynchronize(h->pipe.down));
    CKH(cudaStreamSynchronize(st));
    if (keep) {
        // extrema, scale / bias and bytes of the kept plane: enqueued, not waited for (0.2 ms of GPU time at 24 MP)
        uint8_t *by = (uint8_t *) h->scratch[5];
        int32_t *keys = (int32_t *) (by + ((pixels * PTMB_COEFFICIENTS + 15) & ~(size_t) 15));
        ptmb_scale_bias_t *sb = (ptmb_scale_bias_t *) (keys + PTMB_KEYS);
        if (ptmb_keys_init(c, keys) || ptmb_minmax(c, (const float *) h->scratch[4], pixels, keys) ||
            ptmb_quantise(c, (const float *) h->scratch[4], pixels, keys, by, sb))
            return 1;
        CKH(cudaEventRecord(h->ev[6], st));
        h->spec_host = coeffs_host;
        h->spec_pixels = pixels;
        h->spec_valid = true;
    }
    return 0;
}

// ptm_scale_coefficients with host pointers (rti-builder/ptmlib.c:816-881): extrema over the
// first plane, then every plane quantised with the shared scale / bias.  The floats cross PCIe ONCE:
// they stay on the device between the extrema pass (which runs slab by slab as they arrive) and the
// quantisation.
int ptmb_scale_host(ptmb_ctx_t *c, const float *coeffs_host, size_t pixels, int planes, uint8_t *const *blocks_host,
                    ptmb_scale_bias_t *sb_host) {
    REQH(c && coeffs_host && blocks_host && sb_host && planes >= 1 && planes <= 3, "bad argument");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    HS(c);
    h->valid = false;   // the encode sequence ends here (rti-builder/ptm-encoder.c:406)
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t plane_floats = pixels * PTMB_COEFFICIENTS;
    const bool spec = h->spec_valid && h->spec_host == coeffs_host && h->spec_pixels == pixels && planes == 1 &&
                      pixels > 0 && ptm::HostPipe::is_pinned(coeffs_host) && ptm::HostPipe::is_pinned(blocks_host[0]);
    h->spec_valid = false;
    if (spec) {
        // The floats are (almost certainly) the ones ptm_lrgb_finish just delivered: their bytes exist already and go
        // home on the download stream WHILE the floats come up and are compared, bit for bit, with the kept plane.
        // Any difference (the caller changed something) falls through to the ordinary path below.
        CKH(h->reserve(0, std::max<size_t>(1, plane_floats * sizeof(float))));
        CKH(h->reserve(3, PTMB_KEYS * sizeof(int32_t) + sizeof(ptmb_scale_bias_t) + 16));
        float *cf = (float *) h->scratch[0];
        const float *kept = (const float *) h->scratch[4];
        const uint8_t *by = (const uint8_t *) h->scratch[5];
        const int32_t *skeys = (const int32_t *) (by + ((plane_floats + 15) & ~(size_t) 15));
        const ptmb_scale_bias_t *ssb = (const ptmb_scale_bias_t *) (skeys + PTMB_KEYS);
        int *flag = (int *) ((uint8_t *) h->scratch[3] + PTMB_KEYS * sizeof(int32_t) + sizeof(ptmb_scale_bias_t));
        CKH(cudaMemsetAsync(flag, 0, sizeof(int), st));
        CKH(cudaStreamWaitEvent(h->pipe.down, h->ev[6], 0));   // the speculative quantisation
        CKH(cudaMemcpyAsync(blocks_host[0], by, plane_floats, cudaMemcpyDeviceToHost, h->pipe.down));
        const size_t slab_px = std::min<size_t>(pixels, 2u << 20);
        for (size_t p0 = 0; p0 < pixels; p0 += slab_px) {
            const size_t npx = std::min(slab_px, pixels - p0), o = p0 * PTMB_COEFFICIENTS;
            CKH(cudaMemcpyAsync(cf + o, coeffs_host + o, npx * PTMB_COEFFICIENTS * sizeof(float), cudaMemcpyHostToDevice,
                                h->pipe.up));
            CKH(cudaEventRecord(h->ev[0], h->pipe.up));
            CKH(cudaStreamWaitEvent(st, h->ev[0], 0));
            // 24-byte pixels: a slab of npx pixels is 1.5 npx 16-byte pieces (npx even for every slab but a last odd one)
            const size_t n16 = npx * PTMB_COEFFICIENTS * sizeof(float) / 16;
            planes_differ_kernel<<<296, 256, 0, st>>>((const uint4 *) (cf + o), (const uint4 *) (kept + o), n16, flag);
            CKH(cudaGetLastError());
        }
        int differ = 1;
        // an odd pixel count leaves 8 bytes uncompared: treat as different (ordinary path)
        const bool whole = (plane_floats * sizeof(float)) % 16 == 0 && slab_px % 2 == 0;
        // (sb_host and &differ are pageable: these two copies block the host, so they come after everything is queued)
        CKH(cudaMemcpyAsync(&differ, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CKH(cudaMemcpyAsync(sb_host, ssb, sizeof(ptmb_scale_bias_t), cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
        CKH(cudaStreamSynchronize(h->pipe.down));
        if (whole && differ == 0)
            return 0;
    }
    CKH(h->reserve(0, std::max<size_t>(1, plane_floats * planes * sizeof(float))));
    CKH(h->reserve(1, std::max<size_t>(1, plane_floats * planes)));
    CKH(h->reserve(3, PTMB_KEYS * sizeof(int32_t) + sizeof(ptmb_scale_bias_t)));
    float *cf = (float *) h->scratch[0];
    uint8_t *by = (uint8_t *) h->scratch[1];
    int32_t *keys = (int32_t *) h->scratch[3];
    ptmb_scale_bias_t *sb = (ptmb_scale_bias_t *) (keys + PTMB_KEYS);
    if (ptmb_keys_init(c, keys))
        return 1;
    const size_t slab = std::min<size_t>(std::max<size_t>(pixels, 1), 2u << 20);
    cudaEvent_t up_ev = h->ev[0];
    for (int b = 0; b < planes; ++b)
        for (size_t p0 = 0; p0 < pixels; p0 += slab) {
            const size_t npx = std::min(slab, pixels - p0), o = (size_t) b * plane_floats + p0 * PTMB_COEFFICIENTS;
            CKH(h->pipe.h2d(cf + o, coeffs_host + o, npx * PTMB_COEFFICIENTS * sizeof(float), h->pipe.up));
            if (b == 0) {   // only the first plane feeds the extrema (ref :831)
                CKH(cudaEventRecord(up_ev, h->pipe.up));
                CKH(cudaStreamWaitEvent(st, up_ev, 0));
                if (ptmb_minmax(c, cf + o, npx, keys))
                    return 1;
            }
        }
    CKH(cudaEventRecord(up_ev, h->pipe.up));
    CKH(cudaStreamWaitEvent(st, up_ev, 0));
    cudaEvent_t q_ev = h->ev[1];
    for (int b = 0; b < planes; ++b)
        for (size_t p0 = 0; p0 < pixels; p0 += slab) {
            const size_t npx = std::min(slab, pixels - p0), o = (size_t) b * plane_floats + p0 * PTMB_COEFFICIENTS;
            if (ptmb_quantise(c, cf + o, npx, keys, by + o, (b == 0 && p0 == 0) ? sb : nullptr))
                return 1;
            CKH(cudaEventRecord(q_ev, st));
            CKH(cudaStreamWaitEvent(h->pipe.down, q_ev, 0));
            CKH(h->pipe.d2h(blocks_host[b] + p0 * PTMB_COEFFICIENTS, by + o, npx * PTMB_COEFFICIENTS, h->pipe.down));
        }
    CKH(cudaMemcpyAsync(sb_host, sb, sizeof(ptmb_scale_bias_t), cudaMemcpyDeviceToHost, st));
    CKH(h->pipe.finish());
    CKH(cudaStreamSynchronize(h->pipe.down));
    CKH(cudaStreamSynchronize(st));
    return 0;
}

// The pixel loop of ptm_write_jpeg with host pointers (rti-builder/ptmlib.c:587-621): blocks in,
// n_dirs RGB rasters out.  planes = 1: blocks[0] coefficients + blocks[1] RGB; planes = 3: R, G, B.
int ptmb_relight_host(ptmb_ctx_t *c, const uint8_t *const *blocks_host, int planes, size_t width, size_t height,
                      const float *scale, const int32_t *bias, const float *uv, unsigned n_dirs,
                      uint8_t *out_host) {
    REQH(c && blocks_host && scale && bias && (planes == 1 || planes == 2 || planes == 3),
         "bad argument (planes: 1 = LRGB, 2 = LUM, 3 = RGB)");
    REQH((uv && out_host) || n_dirs == 0, "NULL argument");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t pixels = width * height;
    if (pixels == 0 || n_dirs == 0)
        return 0;
    DevBuf b0, b1, b2, out;
    CKH(b0.alloc(pixels * 6));
    const size_t b1_bytes = pixels * (planes == 1 ? 3 : planes == 2 ? 2 : 6);   // LUM: 2-byte (cr, cb) records
    CKH(b1.alloc(b1_bytes));
    CKH(b2.alloc(planes == 3 ? pixels * 6 : 1));
    CKH(cudaMemcpyAsync(b0.p, blocks_host[0], pixels * 6, cudaMemcpyHostToDevice, st));
    CKH(cudaMemcpyAsync(b1.p, blocks_host[1], b1_bytes, cudaMemcpyHostToDevice, st));
    if (planes == 3)
        CKH(cudaMemcpyAsync(b2.p, blocks_host[2], pixels * 6, cudaMemcpyHostToDevice, st));
    // directions in batches that keep the output staging below ~1 GiB
    const unsigned batch = (unsigned) std::max<size_t>(1, std::min<size_t>(n_dirs, (1u << 30) / (pixels * 3)));
    CKH(out.alloc((size_t) batch * pixels * 3));
    for (unsigned d0 = 0; d0 < n_dirs; d0 += batch) {
        const unsigned nd = std::min(batch, n_dirs - d0);
        int rc = planes == 1
                     ? ptmb_relight_lrgb(c, (const uint8_t *) b0.p, (const uint8_t *) b1.p, width, height, scale, bias,
                                         uv + 2 * (size_t) d0, nd, (uint8_t *) out.p)
                 : planes == 2
                     ? ptmb_relight_lum(c, (const uint8_t *) b0.p, (const uint8_t *) b1.p, width, height, scale, bias,
                                        uv + 2 * (size_t) d0, nd, (uint8_t *) out.p)
                     : ptmb_relight_rgb(c, (const uint8_t *) b0.p, (const uint8_t *) b1.p, (const uint8_t *) b2.p, width,
                                        height, scale, bias, uv + 2 * (size_t) d0, nd, (uint8_t *) out.p);
        if (rc)
            return 1;
        CKH(cudaMemcpyAsync(out_host + (size_t) d0 * pixels * 3, out.p, (size_t) nd * pixels * 3,
                            cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    return 0;
}

// ptm_normal over a whole PTM with host pointers: exactly one of coeffs_host (float [pixels][6]) and
// coef_bytes_host (uint8 [pixels][6], unscaled with scale / bias) is given; normals_host float [pixels][3].
int ptmb_normal_map_host(ptmb_ctx_t *c, const float *coeffs_host, const uint8_t *coef_bytes_host, const float *scale,
                         const int32_t *bias, size_t pixels, float *normals_host) {
    REQH(c && normals_host && ((coeffs_host != nullptr) != (coef_bytes_host != nullptr)), "bad argument");
    REQH(coeffs_host || (scale && bias), "scale / bias needed for a quantised block");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t slab = std::min<size_t>(pixels, 4u << 20);
    const size_t in_px = coeffs_host ? PTMB_COEFFICIENTS * sizeof(float) : PTMB_COEFFICIENTS;
    DevBuf in, out;
    CKH(in.alloc(slab * in_px));
    CKH(out.alloc(slab * 3 * sizeof(float)));
    for (size_t p0 = 0; p0 < pixels; p0 += slab) {
        const size_t npx = std::min(slab, pixels - p0);
        const void *src = coeffs_host ? (const void *) (coeffs_host + p0 * PTMB_COEFFICIENTS)
                                      : (const void *) (coef_bytes_host + p0 * PTMB_COEFFICIENTS);
        CKH(cudaMemcpyAsync(in.p, src, npx * in_px, cudaMemcpyHostToDevice, st));
        if (coeffs_host ? ptmb_normal_map(c, (const float *) in.p, npx, (float *) out.p)
                        : ptmb_normal_map_blocks(c, (const uint8_t *) in.p, scale, bias, npx, (float *) out.p))
            return 1;
        CKH(cudaMemcpyAsync(normals_host + p0 * 3, out.p, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    return 0;
}

}  // extern "C"

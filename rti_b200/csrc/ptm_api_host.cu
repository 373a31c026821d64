// Host-buffer forms of the individual ptmlib.h steps (include/ptm_b200.h, "host" section).
// These exist so that ptmlib.h callers that keep the reference's call sequence
// (fit, then average, then scale: rti-builder/ptm-encoder.c:316-325,406) still run every
// step on the GPU.  Each call streams its input through device slabs; they are copy bound
// by construction -- the fused ptmb_encode_host_* path is the fast one.
#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/ptm_b200.h"
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
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t pixels = width * height;
    const size_t image_samples = pixels * pixel_stride;  // reference: height * width * pixel_stride
    size_t slab_rows = std::max<size_t>(1, (8u << 20) / std::max<size_t>(1, width * pixel_stride * ss));
    slab_rows = std::min(slab_rows, height);
    const size_t slab_samples = slab_rows * width * pixel_stride;
    DevBuf stage, out;
    CKH(stage.alloc(slab_samples * ss * n_lights));
    CKH(out.alloc(slab_rows * width * PTMB_COEFFICIENTS * sizeof(float)));
    for (size_t row = 0; row < height; row += slab_rows) {
        const size_t rows = std::min(slab_rows, height - row);
        const size_t npx = rows * width, nsamp = npx * pixel_stride;
        // the last pixel's trailing (pixel_stride - 1) samples of the last image may lie outside the
        // caller's buffer when it was offset by a channel: copy only up to the last needed sample
        const size_t need = (npx - 1) * pixel_stride + 1;
        for (unsigned l = 0; l < n_lights; ++l)
            CKH(cudaMemcpyAsync((uint8_t *) stage.p + (size_t) l * slab_samples * ss,
                                (const uint8_t *) samples_host + ((size_t) l * image_samples + row * width * pixel_stride) * ss,
                                (l + 1 == n_lights ? need : nsamp) * ss, cudaMemcpyHostToDevice, st));
        if (ptmb_fit_channel(c, stage.p, sample_type, pixel_stride, slab_samples, npx, (float *) out.p))
            return 1;
        CKH(cudaMemcpyAsync(coeffs_host + row * width * PTMB_COEFFICIENTS, out.p,
                            npx * PTMB_COEFFICIENTS * sizeof(float), cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    return 0;
}

// ptm_cbcr_avg with host pointers (rti-builder/ptmlib.c:762-802).
int ptmb_chroma_avg_host(ptmb_ctx_t *c, const uint8_t *stack_host, size_t pixels, unsigned n_lights,
                         uint8_t *avg_host) {
    REQH(c && stack_host && avg_host && n_lights >= 1, "NULL argument");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t slab = std::min<size_t>(pixels, (8u << 20) / 3);
    DevBuf stage, out;
    CKH(stage.alloc(slab * 3 * n_lights));
    CKH(out.alloc(slab * 3));
    for (size_t p0 = 0; p0 < pixels; p0 += slab) {
        const size_t npx = std::min(slab, pixels - p0);
        for (unsigned l = 0; l < n_lights; ++l)
            CKH(cudaMemcpyAsync((uint8_t *) stage.p + (size_t) l * slab * 3, stack_host + ((size_t) l * pixels + p0) * 3,
                                npx * 3, cudaMemcpyHostToDevice, st));
        if (ptmb_chroma_avg(c, (const uint8_t *) stage.p, slab * 3, npx, n_lights, (uint8_t *) out.p))
            return 1;
        CKH(cudaMemcpyAsync(avg_host + p0 * 3, out.p, npx * 3, cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    return 0;
}

// The encoder's LRGB colour + normalise loop with host pointers (rti-builder/ptm-encoder.c:327-353).
int ptmb_lrgb_colour_normalise_host(ptmb_ctx_t *c, uint8_t *ycc_to_rgb_host, float *coeffs_host, size_t pixels) {
    REQH(c && ycc_to_rgb_host && coeffs_host, "NULL argument");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t slab = std::min<size_t>(pixels, 4u << 20);
    DevBuf ycc, cf;
    CKH(ycc.alloc(slab * 3));
    CKH(cf.alloc(slab * PTMB_COEFFICIENTS * sizeof(float)));
    for (size_t p0 = 0; p0 < pixels; p0 += slab) {
        const size_t npx = std::min(slab, pixels - p0);
        CKH(cudaMemcpyAsync(ycc.p, ycc_to_rgb_host + p0 * 3, npx * 3, cudaMemcpyHostToDevice, st));
        CKH(cudaMemcpyAsync(cf.p, coeffs_host + p0 * PTMB_COEFFICIENTS, npx * PTMB_COEFFICIENTS * sizeof(float),
                            cudaMemcpyHostToDevice, st));
        if (ptmb_lrgb_colour_normalise(c, (uint8_t *) ycc.p, (float *) cf.p, npx))
            return 1;
        CKH(cudaMemcpyAsync(ycc_to_rgb_host + p0 * 3, ycc.p, npx * 3, cudaMemcpyDeviceToHost, st));
        CKH(cudaMemcpyAsync(coeffs_host + p0 * PTMB_COEFFICIENTS, cf.p, npx * PTMB_COEFFICIENTS * sizeof(float),
                            cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    return 0;
}

// ptm_scale_coefficients with host pointers (rti-builder/ptmlib.c:816-881): extrema over the
// first plane, then every plane quantised with the shared scale / bias.
int ptmb_scale_host(ptmb_ctx_t *c, const float *coeffs_host, size_t pixels, int planes, uint8_t *const *blocks_host,
                    ptmb_scale_bias_t *sb_host) {
    REQH(c && coeffs_host && blocks_host && sb_host && planes >= 1 && planes <= 3, "bad argument");
    CKH(cudaSetDevice(ptmb_internal_device(c)));
    cudaStream_t st = (cudaStream_t) ptmb_internal_stream(c);
    const size_t slab = std::min<size_t>(pixels, 4u << 20);
    DevBuf cf, by, keys, sb;
    CKH(cf.alloc(slab * PTMB_COEFFICIENTS * sizeof(float)));
    CKH(by.alloc(slab * PTMB_COEFFICIENTS));
    CKH(keys.alloc(PTMB_KEYS * sizeof(int32_t)));
    CKH(sb.alloc(sizeof(ptmb_scale_bias_t)));
    if (ptmb_keys_init(c, (int32_t *) keys.p))
        return 1;
    for (size_t p0 = 0; p0 < pixels; p0 += slab) {
        const size_t npx = std::min(slab, pixels - p0);
        CKH(cudaMemcpyAsync(cf.p, coeffs_host + p0 * PTMB_COEFFICIENTS, npx * PTMB_COEFFICIENTS * sizeof(float),
                            cudaMemcpyHostToDevice, st));
        if (ptmb_minmax(c, (const float *) cf.p, npx, (int32_t *) keys.p))
            return 1;
        CKH(cudaStreamSynchronize(st));
    }
    for (int b = 0; b < planes; ++b) {
        const float *src = coeffs_host + (size_t) b * pixels * PTMB_COEFFICIENTS;
        for (size_t p0 = 0; p0 < pixels; p0 += slab) {
            const size_t npx = std::min(slab, pixels - p0);
            CKH(cudaMemcpyAsync(cf.p, src + p0 * PTMB_COEFFICIENTS, npx * PTMB_COEFFICIENTS * sizeof(float),
                                cudaMemcpyHostToDevice, st));
            if (ptmb_quantise(c, (const float *) cf.p, npx, (const int32_t *) keys.p, (uint8_t *) by.p,
                              (ptmb_scale_bias_t *) sb.p))
                return 1;
            CKH(cudaMemcpyAsync(blocks_host[b] + p0 * PTMB_COEFFICIENTS, by.p, npx * PTMB_COEFFICIENTS,
                                cudaMemcpyDeviceToHost, st));
            CKH(cudaStreamSynchronize(st));
        }
    }
    CKH(cudaMemcpyAsync(sb_host, sb.p, sizeof(ptmb_scale_bias_t), cudaMemcpyDeviceToHost, st));
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

// K1 -- streaming PTM fit kernels for sm_100a.
//
// Input is the reference encoder's decoded stack (rti-builder/ptm-encoder.c:245-264):
// samples [light][pixel][3], so for a run of consecutive pixels every light
// contributes ONE contiguous byte run.  The reference walks it pixel by pixel
// with an n-way gather at image stride (rti-builder/ptmlib.c:712-715); here each
// thread owns a few consecutive pixels, streams over the lights and keeps the
// six accumulators per pixel in registers.
#include "ptm_device.cuh"
#include "ptm_kernels.h"

namespace ptm {

// ---------------------------------------------------------------------------
// LRGB, uint8 triplets, 4 pixels (12 bytes = 3 aligned words) per thread.
//
//   word0 = Y0 Cb0 Cr0 Y1 | word1 = Cb1 Cr1 Y2 Cb2 | word2 = Cr2 Y3 Cb3 Cr3
//
// Luma bytes go PRMT -> FADD -> 6 FFMA against the pseudo-inverse column held
// in shared memory.  All twelve bytes are also summed for ptm_cbcr_avg with
// 16-bit SIMD-in-a-register lanes: (w & 0x00FF00FF) and ((w >> 8) & 0x00FF00FF)
// added into two accumulators per word (sums <= 255 * n < 65536 for n <= 257).
// ---------------------------------------------------------------------------
template <bool kWriteCoeffs>
__global__ void __launch_bounds__(kFitThreads, 3)
fit_lrgb_u8_kernel(FitArgs a) {
    extern __shared__ float4 m_s[];  // [n][2]: (m0 m1 m2 m3) (m4 m5 0 0)
    __shared__ int keys_s[2 * PTM_NCOEF];

    for (unsigned i = threadIdx.x; i < a.n_lights; i += blockDim.x) {
        const float *M = a.M;
        const unsigned n = a.n_lights;
        m_s[2 * i] = make_float4(M[i], M[n + i], M[2 * n + i], M[3 * n + i]);
        m_s[2 * i + 1] = make_float4(M[4 * n + i], M[5 * n + i], 0.f, 0.f);
    }
    __syncthreads();

    Extrema ext;
    ext.init();

    const size_t groups = a.pixels / 4;
    const size_t stride_w = a.image_stride / 4;  // words between lights
    const uint32_t *base = reinterpret_cast<const uint32_t *>(a.stack);

    for (size_t g = (size_t) blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (size_t) gridDim.x * blockDim.x) {
        const uint32_t *src = base + g * 3;
        float acc[4][PTM_NCOEF];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                acc[p][k] = 0.f;
        unsigned ev[3] = {0, 0, 0}, od[3] = {0, 0, 0};

#pragma unroll 4
        for (unsigned n = 0; n < a.n_lights; ++n) {
            const uint32_t *s = src + (size_t) n * stride_w;
            unsigned w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2);
            if (a.mode & kFitRgbIn)
                rgb_words_to_ycc(w0, w1, w2);
            const float4 ma = m_s[2 * n], mb = m_s[2 * n + 1];
            const float y[4] = {byte_to_float(w0, PTM_SEL_B0), byte_to_float(w0, PTM_SEL_B3),
                                byte_to_float(w1, PTM_SEL_B2), byte_to_float(w2, PTM_SEL_B1)};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                acc[p][0] = fmaf(ma.x, y[p], acc[p][0]);
                acc[p][1] = fmaf(ma.y, y[p], acc[p][1]);
                acc[p][2] = fmaf(ma.z, y[p], acc[p][2]);
                acc[p][3] = fmaf(ma.w, y[p], acc[p][3]);
                acc[p][4] = fmaf(mb.x, y[p], acc[p][4]);
                acc[p][5] = fmaf(mb.y, y[p], acc[p][5]);
            }
            ev[0] += w0 & 0x00FF00FFu;
            od[0] += (w0 >> 8) & 0x00FF00FFu;
            ev[1] += w1 & 0x00FF00FFu;
            od[1] += (w1 >> 8) & 0x00FF00FFu;
            ev[2] += w2 & 0x00FF00FFu;
            od[2] += (w2 >> 8) & 0x00FF00FFu;
        }

        // byte b of word j summed: b even -> ev[j] lane b/2, b odd -> od[j] lane b/2
        const unsigned sum[12] = {ev[0] & 0xFFFFu, od[0] & 0xFFFFu, ev[0] >> 16, od[0] >> 16,
                                  ev[1] & 0xFFFFu, od[1] & 0xFFFFu, ev[1] >> 16, od[1] >> 16,
                                  ev[2] & 0xFFFFu, od[2] & 0xFFFFu, ev[2] >> 16, od[2] >> 16};
        unsigned out_b[12];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const unsigned ya = div_magic(sum[3 * p], a.magic);
            const unsigned cba = div_magic(sum[3 * p + 1], a.magic);
            const unsigned cra = div_magic(sum[3 * p + 2], a.magic);
            lrgb_or_lum_finish(a.mode & kFitLum, ya, cba, cra, acc[p], out_b[3 * p], out_b[3 * p + 1], out_b[3 * p + 2]);
            ext.add(acc[p]);
        }
        if (kWriteCoeffs) {
            float4 *dst = reinterpret_cast<float4 *>(a.coeffs + g * 24);
            const float *f = &acc[0][0];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        uint32_t *rgb = reinterpret_cast<uint32_t *>(a.rgb + g * 12);
        rgb[0] = out_b[0] | (out_b[1] << 8) | (out_b[2] << 16) | (out_b[3] << 24);
        rgb[1] = out_b[4] | (out_b[5] << 8) | (out_b[6] << 16) | (out_b[7] << 24);
        rgb[2] = out_b[8] | (out_b[9] << 8) | (out_b[10] << 16) | (out_b[11] << 24);
    }
    ext.flush(a.keys, keys_s);
}

// ---------------------------------------------------------------------------
// Generic LRGB path: one pixel per thread, byte / halfword loads.  Used for the
// (< 4 pixel) tail of the vector kernel and for stacks whose base or image
// stride is not word aligned, or with more than 257 lights.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFitThreads)
fit_lrgb_generic_kernel(FitArgs a, size_t first_pixel) {
    // 16-bit stacks (no counterpart in the reference, whose ptm_cbcr_avg is 8 bit only): the fit
    // runs on the 16-bit luma exactly like ptm_fit_poly_uint, the means are taken at 16 bits, the
    // colour conversion uses their top 8 bits, and the normalisation is 256 / (16-bit mean luma),
    // which yields the same dimensionless coefficients as the 8-bit path.
    constexpr int kShift = sizeof(T) == 2 ? 8 : 0;
    extern __shared__ float m_gen[];  // [n][6]
    __shared__ int keys_s[2 * PTM_NCOEF];
    for (unsigned i = threadIdx.x; i < a.n_lights * PTM_NCOEF; i += blockDim.x)
        m_gen[i] = a.M[(size_t) (i % PTM_NCOEF) * a.n_lights + i / PTM_NCOEF];
    __syncthreads();
    Extrema ext;
    ext.init();
    const size_t count = a.pixels - first_pixel;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (size_t) gridDim.x * blockDim.x) {
        const size_t p = first_pixel + i;
        float acc[PTM_NCOEF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll 4
        for (unsigned n = 0; n < a.n_lights; ++n) {
            const T *t = reinterpret_cast<const T *>(a.stack + (size_t) n * a.image_stride) + p * 3;
            unsigned v0 = __ldg(t), v1 = __ldg(t + 1), v2 = __ldg(t + 2);
            if (kShift == 0 && (a.mode & kFitRgbIn))
                rgb_to_ycc(v0, v1, v2, v0, v1, v2);
            const float y = (float) v0;
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                acc[k] = fmaf(m_gen[n * PTM_NCOEF + k], y, acc[k]);
            s0 += v0;
            s1 += v1;
            s2 += v2;
        }
        const unsigned ya = s0 / a.n_lights, cba = s1 / a.n_lights, cra = s2 / a.n_lights;
        unsigned r, g, b;
        if (kShift == 0) {
            lrgb_or_lum_finish(a.mode & kFitLum, ya, cba, cra, acc, r, g, b);
        } else {
            float scratch[PTM_NCOEF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            lrgb_finish(ya >> kShift, cba >> kShift, cra >> kShift, scratch, r, g, b);   // colour only
            const float norm = __fdiv_rn(256.0f, (float) ya);
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                acc[k] = __fmul_rn(acc[k], norm);
        }
        ext.add(acc);
        if (a.coeffs) {
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                a.coeffs[p * PTM_NCOEF + k] = acc[k];
        }
        a.rgb[p * 3 + 0] = (uint8_t) r;
        a.rgb[p * 3 + 1] = (uint8_t) g;
        a.rgb[p * 3 + 2] = (uint8_t) b;
    }
    ext.flush(a.keys, keys_s);
}

// ---------------------------------------------------------------------------
// RGB formats, uint8 triplets: three independent fits per pixel
// (rti-builder/ptm-encoder.c:272-284) in one pass over the stack, 4 pixels per
// thread (72 accumulators).  Output planes R, G, B of [pixels][6] floats.  Only
// plane 0 contributes to the extrema (rti-builder/ptmlib.c:831).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFitThreads)
fit_rgb_u8_kernel(FitArgs a) {
    extern __shared__ float4 m_s[];
    __shared__ int keys_s[2 * PTM_NCOEF];
    for (unsigned i = threadIdx.x; i < a.n_lights; i += blockDim.x) {
        const float *M = a.M;
        const unsigned n = a.n_lights;
        m_s[2 * i] = make_float4(M[i], M[n + i], M[2 * n + i], M[3 * n + i]);
        m_s[2 * i + 1] = make_float4(M[4 * n + i], M[5 * n + i], 0.f, 0.f);
    }
    __syncthreads();

    Extrema ext;
    ext.init();
    const size_t groups = a.pixels / 4;
    const size_t stride_w = a.image_stride / 4;
    const uint32_t *base = reinterpret_cast<const uint32_t *>(a.stack);

    for (size_t g = (size_t) blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (size_t) gridDim.x * blockDim.x) {
        const uint32_t *src = base + g * 3;
        float acc[12][PTM_NCOEF];  // [pixel * 3 + channel]
#pragma unroll
        for (int p = 0; p < 12; ++p)
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                acc[p][k] = 0.f;
#pragma unroll 2
        for (unsigned n = 0; n < a.n_lights; ++n) {
            const uint32_t *s = src + (size_t) n * stride_w;
            const unsigned w[3] = {__ldg(s), __ldg(s + 1), __ldg(s + 2)};
            const float4 ma = m_s[2 * n], mb = m_s[2 * n + 1];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float v[4] = {byte_to_float(w[j], PTM_SEL_B0), byte_to_float(w[j], PTM_SEL_B1),
                                    byte_to_float(w[j], PTM_SEL_B2), byte_to_float(w[j], PTM_SEL_B3)};
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float *c = acc[4 * j + b];
                    c[0] = fmaf(ma.x, v[b], c[0]);
                    c[1] = fmaf(ma.y, v[b], c[1]);
                    c[2] = fmaf(ma.z, v[b], c[2]);
                    c[3] = fmaf(ma.w, v[b], c[3]);
                    c[4] = fmaf(mb.x, v[b], c[4]);
                    c[5] = fmaf(mb.y, v[b], c[5]);
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float4 *dst = reinterpret_cast<float4 *>(a.coeffs + (size_t) ch * a.plane_stride + g * 24);
            float f[24];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int k = 0; k < PTM_NCOEF; ++k)
                    f[p * PTM_NCOEF + k] = acc[p * 3 + ch][k];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
            ext.add(acc[p * 3]);
    }
    ext.flush(a.keys, keys_s);
}

template <typename T>
__global__ void __launch_bounds__(kFitThreads)
fit_rgb_generic_kernel(FitArgs a, size_t first_pixel) {
    extern __shared__ float m_gen[];  // [n][6]
    __shared__ int keys_s[2 * PTM_NCOEF];
    for (unsigned i = threadIdx.x; i < a.n_lights * PTM_NCOEF; i += blockDim.x)
        m_gen[i] = a.M[(size_t) (i % PTM_NCOEF) * a.n_lights + i / PTM_NCOEF];
    __syncthreads();
    Extrema ext;
    ext.init();
    const size_t count = (a.pixels - first_pixel) * 3;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (size_t) gridDim.x * blockDim.x) {
        const size_t p = first_pixel + i / 3;
        const unsigned ch = (unsigned) (i % 3);
        float acc[PTM_NCOEF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (unsigned n = 0; n < a.n_lights; ++n) {
            const T *t = reinterpret_cast<const T *>(a.stack + (size_t) n * a.image_stride) + p * 3 + ch;
            const float v = (float) (unsigned) __ldg(t);
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                acc[k] = fmaf(m_gen[n * PTM_NCOEF + k], v, acc[k]);
        }
        if (ch == 0)
            ext.add(acc);
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            a.coeffs[(size_t) ch * a.plane_stride + p * PTM_NCOEF + k] = acc[k];
    }
    ext.flush(a.keys, keys_s);
}

// ---------------------------------------------------------------------------
// ptm_fit_poly_jsample / ptm_fit_poly_uint in their general form
// (rti-builder/ptmlib.c:692-760): one channel, any pixel stride.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFitThreads)
fit_channel_kernel(const T *samples, size_t pixel_stride, size_t image_stride, size_t pixels,
                   unsigned n_lights, const float *M, float *coeffs) {
    for (size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x; p < pixels;
         p += (size_t) gridDim.x * blockDim.x) {
        float acc[PTM_NCOEF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const T *s = samples + p * pixel_stride;
        for (unsigned n = 0; n < n_lights; ++n) {
            const float v = (float) (unsigned) s[(size_t) n * image_stride];
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                acc[k] = fmaf(M[(size_t) k * n_lights + n], v, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            coeffs[p * PTM_NCOEF + k] = acc[k];
    }
}

// ptm_cbcr_avg alone (rti-builder/ptmlib.c:762-802); one byte position per thread.
__global__ void __launch_bounds__(kFitThreads)
chroma_avg_kernel(const uint8_t *stack, size_t image_stride, size_t bytes, unsigned n_lights, uint8_t *avg) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < bytes;
         i += (size_t) gridDim.x * blockDim.x) {
        unsigned s = 0;
        for (unsigned n = 0; n < n_lights; ++n)
            s += stack[(size_t) n * image_stride + i];
        avg[i] = (uint8_t) (s / n_lights);
    }
}

// The LRGB colour + normalise loop alone (rti-builder/ptm-encoder.c:327-353).
__global__ void __launch_bounds__(kFitThreads)
lrgb_colour_normalise_kernel(uint8_t *ycc, float *coeffs, size_t pixels) {
    for (size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x; p < pixels;
         p += (size_t) gridDim.x * blockDim.x) {
        float c[PTM_NCOEF];
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            c[k] = coeffs[p * PTM_NCOEF + k];
        unsigned r, g, b;
        lrgb_finish(ycc[p * 3], ycc[p * 3 + 1], ycc[p * 3 + 2], c, r, g, b);
        ycc[p * 3] = (uint8_t) r;
        ycc[p * 3 + 1] = (uint8_t) g;
        ycc[p * 3 + 2] = (uint8_t) b;
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            coeffs[p * PTM_NCOEF + k] = c[k];
    }
}

// Phase A of ptm_scale_coefficients alone (rti-builder/ptmlib.c:826-847).
__global__ void __launch_bounds__(kFitThreads)
minmax_kernel(const float *coeffs, size_t pixels, int *keys) {
    __shared__ int keys_s[2 * PTM_NCOEF];
    Extrema ext;
    ext.init();
    for (size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x; p < pixels;
         p += (size_t) gridDim.x * blockDim.x) {
        const float2 *s = reinterpret_cast<const float2 *>(coeffs + p * PTM_NCOEF);
        const float2 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2);
        const float v[PTM_NCOEF] = {a.x, a.y, b.x, b.y, c.x, c.y};
        ext.add(v);
    }
    ext.flush(keys, keys_s);
}

__global__ void keys_init_kernel(int *keys) {
    if (threadIdx.x < 2 * PTM_NCOEF)
        keys[threadIdx.x] = float_to_key(-FLT_MAX);
}

// ------------------------------------------------------------------ launchers

static inline unsigned grid_for(size_t items, unsigned threads, int sm_count, int blocks_per_sm) {
    size_t want = (items + threads - 1) / threads;
    size_t cap = (size_t) sm_count * blocks_per_sm;
    if (want < 1)
        want = 1;
    return (unsigned) (want < cap ? want : cap);
}

// Which engine serves a fit: the caller's explicit choice, else PTM_FIT_ENGINE (ffma | mma),
// else the per-format default.  1 = FFMA kernels, 2 = tensor-core kernel.
int resolve_engine(int requested, int format_default) {
    if (requested == 1 || requested == 2)
        return requested;
    static int env = -1;
    if (env < 0) {
        const char *v = getenv("PTM_FIT_ENGINE");
        env = !v ? 0 : !strcmp(v, "mma") ? 2 : !strcmp(v, "ffma") ? 1 : 0;
    }
    return env ? env : format_default;
}

cudaError_t launch_keys_init(int *keys, cudaStream_t st) {
    keys_init_kernel<<<1, 32, 0, st>>>(keys);
    return cudaGetLastError();
}

// The pinv table of the non-streaming kernels lives in dynamic shared memory: 24 (generic) or 32 (register
// kernels) bytes per light, i.e. up to 64 KB at the 2048 lights ptmb_set_matrix accepts -- above the 48 KB a
// launch may use without opting in.  Attributes are per device, hence the per-device flag.
static cudaError_t allow_large_tables() {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= 64 || done[dev])
        return e;
    const int bytes = 2048 * 2 * (int) sizeof(float4);
#define PTM_ALLOW(k)                                                                         \
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) \
        return e;
    PTM_ALLOW(fit_lrgb_generic_kernel<uint8_t>)
    PTM_ALLOW(fit_lrgb_generic_kernel<uint16_t>)
    PTM_ALLOW(fit_rgb_generic_kernel<uint8_t>)
    PTM_ALLOW(fit_rgb_generic_kernel<uint16_t>)
    PTM_ALLOW(fit_lrgb_u8_kernel<true>)
    PTM_ALLOW(fit_lrgb_u8_kernel<false>)
    PTM_ALLOW(fit_rgb_u8_kernel)
#undef PTM_ALLOW
    done[dev] = true;
    return cudaSuccess;
}

cudaError_t launch_fit_lrgb(const FitArgs &a, int sample_type, int sm_count, cudaStream_t st) {
    if (a.pixels == 0)
        return cudaSuccess;
    if (cudaError_t ea = allow_large_tables())
        return ea;
    if (sample_type != 0 && sample_type != 1)
        return cudaErrorInvalidValue;
    const size_t gen_smem = (size_t) a.n_lights * PTM_NCOEF * sizeof(float);
    if (a.mode != 0 && sample_type != 0)
        return cudaErrorNotSupported;   // LUM / RGB-input are 8-bit paths, like the reference's
    if (sample_type == 1) {   // 16-bit stacks: streaming kernel over the tiles, generic kernel for the rest
        size_t done16 = 0;
        cudaError_t e16 = cudaSuccess;
        if (resolve_engine(a.engine, kDefaultEngineLrgbU16) == 2)
            e16 = launch_fit_mma(a, 3, sm_count, st, &done16);
        if (e16 == cudaSuccess && done16 == 0)
            e16 = launch_fit_u16_tma(a, false, sm_count, st, &done16);
        if (e16 != cudaSuccess)
            return e16;
        if (done16 < a.pixels) {
            const unsigned grid = grid_for(a.pixels - done16, kFitThreads, sm_count, 8);
            fit_lrgb_generic_kernel<uint16_t><<<grid, kFitThreads, gen_smem, st>>>(a, done16);
        }
        return cudaGetLastError();
    }
    // 1) whole 1024-pixel tiles through the streaming (bulk copy) kernel when the layout allows it,
    //    or whole 128-pixel units through the tensor-core kernel
    size_t done = 0;
    cudaError_t e = cudaSuccess;
    if (a.mode == 0 && resolve_engine(a.engine, kDefaultEngineLrgbU8) == 2)
        e = launch_fit_mma(a, 1, sm_count, st, &done);
    if (e == cudaSuccess && done == 0)
        e = launch_fit_lrgb_tma(a, sm_count, st, &done);
    if (e != cudaSuccess)
        return e;
    // 2) otherwise 4-pixel groups through the register kernel (word-aligned layouts)
    const bool vec_ok = a.n_lights <= 257 && ((uintptr_t) a.stack % 4 == 0) && (a.image_stride % 4 == 0) &&
                        ((uintptr_t) a.rgb % 4 == 0) && (a.coeffs == nullptr || (uintptr_t) a.coeffs % 16 == 0);
    if (done == 0 && vec_ok && a.pixels >= 4) {
        const size_t groups = a.pixels / 4;
        const unsigned grid = grid_for(groups, kFitThreads, sm_count, 3);
        const size_t smem = (size_t) a.n_lights * 2 * sizeof(float4);
        if (a.coeffs)
            fit_lrgb_u8_kernel<true><<<grid, kFitThreads, smem, st>>>(a);
        else
            fit_lrgb_u8_kernel<false><<<grid, kFitThreads, smem, st>>>(a);
        done = groups * 4;
    }
    // 3) whatever is left (tile / group remainders, unaligned stacks, > 257 lights): one pixel per thread
    if (done < a.pixels) {
        const unsigned grid = grid_for(a.pixels - done, kFitThreads, sm_count, 8);
        fit_lrgb_generic_kernel<uint8_t><<<grid, kFitThreads, gen_smem, st>>>(a, done);
    }
    return cudaGetLastError();
}

cudaError_t launch_fit_rgb(const FitArgs &a, int sample_type, int sm_count, cudaStream_t st) {
    if (a.pixels == 0)
        return cudaSuccess;
    if (cudaError_t ea = allow_large_tables())
        return ea;
    const size_t gen_smem = (size_t) a.n_lights * PTM_NCOEF * sizeof(float);
    if (sample_type == 1) {   // 16-bit stacks: streaming kernel over the tiles, generic kernel for the rest
        size_t done16 = 0;
        cudaError_t e16 = cudaSuccess;
        if (resolve_engine(a.engine, kDefaultEngineRgbU16) == 2)
            e16 = launch_fit_mma(a, 2, sm_count, st, &done16);
        if (e16 == cudaSuccess && done16 == 0)
            e16 = launch_fit_u16_tma(a, true, sm_count, st, &done16);
        if (e16 != cudaSuccess)
            return e16;
        if (done16 < a.pixels) {
            const unsigned grid = grid_for((a.pixels - done16) * 3, kFitThreads, sm_count, 8);
            fit_rgb_generic_kernel<uint16_t><<<grid, kFitThreads, gen_smem, st>>>(a, done16);
        }
        return cudaGetLastError();
    }
    if (sample_type != 0)
        return cudaErrorInvalidValue;
    // 1) streaming kernel over whole tiles (+ a 16-pixel-granular partial tile)
    size_t done = 0;
    cudaError_t e = cudaSuccess;
    if (resolve_engine(a.engine, kDefaultEngineRgbU8) == 2)
        e = launch_fit_mma(a, 0, sm_count, st, &done);
    if (e == cudaSuccess && done == 0)
        e = launch_fit_rgb_tma(a, sm_count, st, &done);
    if (e != cudaSuccess)
        return e;
    // 2) 4-pixel groups through the register kernel for word-aligned layouts
    const bool vec_ok = ((uintptr_t) a.stack % 4 == 0) && (a.image_stride % 4 == 0) &&
                        ((uintptr_t) a.coeffs % 16 == 0) && (a.plane_stride % 4 == 0);
    if (done == 0 && vec_ok && a.pixels >= 4) {
        const size_t groups = a.pixels / 4;
        const unsigned grid = grid_for(groups, kFitThreads, sm_count, 2);
        const size_t smem = (size_t) a.n_lights * 2 * sizeof(float4);
        fit_rgb_u8_kernel<<<grid, kFitThreads, smem, st>>>(a);
        done = groups * 4;
    }
    // 3) the rest, one (pixel, channel) per thread
    if (done < a.pixels) {
        const unsigned grid = grid_for((a.pixels - done) * 3, kFitThreads, sm_count, 8);
        fit_rgb_generic_kernel<uint8_t><<<grid, kFitThreads, gen_smem, st>>>(a, done);
    }
    return cudaGetLastError();
}

cudaError_t launch_fit_channel(const void *samples, int sample_type, size_t pixel_stride, size_t image_stride,
                               size_t pixels, unsigned n_lights, const float *M, float *coeffs, int sm_count,
                               cudaStream_t st) {
    if (pixels == 0)
        return cudaSuccess;
    const unsigned grid = grid_for(pixels, kFitThreads, sm_count, 8);
    if (sample_type == 0)
        fit_channel_kernel<uint8_t><<<grid, kFitThreads, 0, st>>>((const uint8_t *) samples, pixel_stride,
                                                                 image_stride, pixels, n_lights, M, coeffs);
    else if (sample_type == 1)
        fit_channel_kernel<uint16_t><<<grid, kFitThreads, 0, st>>>((const uint16_t *) samples, pixel_stride,
                                                                  image_stride, pixels, n_lights, M, coeffs);
    else if (sample_type == 2)
        fit_channel_kernel<uint32_t><<<grid, kFitThreads, 0, st>>>((const uint32_t *) samples, pixel_stride,
                                                                  image_stride, pixels, n_lights, M, coeffs);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_chroma_avg(const uint8_t *stack, size_t image_stride, size_t pixels, unsigned n_lights,
                              uint8_t *avg, int sm_count, cudaStream_t st) {
    if (pixels == 0)
        return cudaSuccess;
    chroma_avg_kernel<<<grid_for(pixels * 3, kFitThreads, sm_count, 8), kFitThreads, 0, st>>>(
        stack, image_stride, pixels * 3, n_lights, avg);
    return cudaGetLastError();
}

cudaError_t launch_lrgb_colour_normalise(uint8_t *ycc, float *coeffs, size_t pixels, int sm_count,
                                         cudaStream_t st) {
    if (pixels == 0)
        return cudaSuccess;
    lrgb_colour_normalise_kernel<<<grid_for(pixels, kFitThreads, sm_count, 8), kFitThreads, 0, st>>>(ycc, coeffs,
                                                                                                      pixels);
    return cudaGetLastError();
}

cudaError_t launch_minmax(const float *coeffs, size_t pixels, int *keys, int sm_count, cudaStream_t st) {
    if (pixels == 0)
        return cudaSuccess;
    minmax_kernel<<<grid_for(pixels, kFitThreads, sm_count, 8), kFitThreads, 0, st>>>(coeffs, pixels, keys);
    return cudaGetLastError();
}

}  // namespace ptm

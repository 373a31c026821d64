// K3 -- relighting: the pixel loop of ptm_write_jpeg (rti-builder/ptmlib.c:553-562,
// 587-621) evaluated for a batch of light directions.
//
// POLY / UNSCALE (rti-builder/ptmlib.c:54-63): (scale_k * (float)(byte_k - bias_k)) is
// direction independent, so it is computed once per pixel and kept in registers; per
// direction the five products are added left to right with separate multiplies and adds,
// exactly the reference's association.  LRGB: L = POLY / 255.0f (:595), then CLIP(L * colour).
// Output row y reads stored row height-1-y (:587-589).
//
// Everything that produces a byte is kept bit-exact with the reference's C:
//   x / 255.0f   q0 = RN(x * r), e = fma(-q0, 255, x), q = fma(e, r, q0) with r = RN(1/255): three
//                FMA-pipe instructions instead of the IEEE division sequence.  Checked against
//                __fdiv_rn for EVERY finite float by div255_selftest_kernel (tests); only x = +-inf
//                differs (NaN instead of inf), which needs an overflowing polynomial.
//   CLIP(f)      f > 255 -> 255, f < 0 -> 0, else truncate: min(max(f, 0), 255) then
//                __fadd_rz(., 2^23), whose low mantissa byte is the truncated value -- no trip
//                through the (quarter-rate) conversion unit.  For LRGB the lower clamp is applied
//                once to L (the colour bytes are non-negative).
#include <cstdlib>

#include "ptm_device.cuh"
#include "ptm_kernels.h"

namespace ptm {

constexpr int kRelightThreads = 256;

__device__ __forceinline__ float div255(float x) {
    const float r = 0.003921568859368562698364257812f;  // RN(1 / 255)
    const float q0 = __fmul_rn(x, r);
    const float e = fmaf(-q0, 255.0f, x);
    return fmaf(e, r, q0);
}

__device__ __forceinline__ float poly_eval(const float *a, const float *l) {
    float s = __fmul_rn(a[0], l[0]);
    s = __fadd_rn(s, __fmul_rn(a[1], l[1]));
    s = __fadd_rn(s, __fmul_rn(a[2], l[2]));
    s = __fadd_rn(s, __fmul_rn(a[3], l[3]));
    s = __fadd_rn(s, __fmul_rn(a[4], l[4]));
    return __fadd_rn(s, a[5]);
}

// float in [0, 255] -> word whose low byte is trunc(f)
__device__ __forceinline__ unsigned trunc_bits(float f) { return __float_as_uint(__fadd_rz(f, 8388608.0f)); }

// CLIP for a value already known to be >= 0 (or -0)
__device__ __forceinline__ unsigned clip_hi_bits(float f) { return trunc_bits(fminf(f, 255.0f)); }
__device__ __forceinline__ unsigned clip_bits(float f) { return trunc_bits(fminf(fmaxf(f, 0.0f), 255.0f)); }

// low bytes of four words -> one word
__device__ __forceinline__ unsigned pack4(unsigned a, unsigned b, unsigned c, unsigned d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

__device__ __forceinline__ void unscale6(const uint8_t *c, const RelightArgs &a, float *out) {
#pragma unroll
    for (int k = 0; k < PTM_NCOEF; ++k)
        out[k] = __fmul_rn(a.scale[k], (float) ((int) c[k] - a.bias[k]));
}

// ---------------------------------------------------------------------------
// LRGB, 4 consecutive pixels of an output row per thread (width % 4 == 0).
// Inputs: 24 B of coefficients (3 x 8-byte loads) + 12 B of colour (3 words) per thread.
// Per direction each thread emits 12 bytes as 3 words.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kRelightThreads)
relight_lrgb_x4_kernel(RelightArgs a) {
    extern __shared__ float light_s[];  // [n_dirs][8]: u^2 v^2 uv u v (1) pad pad
    for (unsigned i = threadIdx.x; i < a.n_dirs * 8; i += blockDim.x)
        light_s[i] = (i & 7) < PTM_NCOEF ? a.light[(i >> 3) * PTM_NCOEF + (i & 7)] : 0.f;
    __syncthreads();

    const size_t pixels = a.width * a.height;
    const size_t groups = pixels / 4;
    const size_t gw = a.width / 4;  // groups per row
    for (size_t g = (size_t) blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (size_t) gridDim.x * blockDim.x) {
        const size_t y = g / gw, xg = g - y * gw;
        const size_t p = (a.height - 1 - y) * a.width + xg * 4;  // first stored pixel of the group
        const uint2 *c2 = reinterpret_cast<const uint2 *>(a.plane[0] + p * PTM_NCOEF);
        const uint32_t *r32 = reinterpret_cast<const uint32_t *>(a.plane[1] + p * 3);
        const uint2 cw[3] = {__ldg(c2), __ldg(c2 + 1), __ldg(c2 + 2)};
        const uint32_t rw[3] = {__ldg(r32), __ldg(r32 + 1), __ldg(r32 + 2)};
        const uint32_t cb[6] = {cw[0].x, cw[0].y, cw[1].x, cw[1].y, cw[2].x, cw[2].y};
        // (float)(byte - bias) without the conversion unit: both are integers below 2^23 (checked
        // by the launcher), so (2^23 + byte) - (2^23 + bias) is exact and equals the int subtraction
        float u[4][PTM_NCOEF], col[4][3];
        float kb[PTM_NCOEF];
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            kb[k] = 8388608.0f + (float) a.bias[k];
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            const float m = __uint_as_float(__byte_perm(cb[j >> 2], 0x4B000000u, 0x7650u + (j & 3)));
            u[j / 6][j % 6] = __fmul_rn(a.scale[j % 6], __fsub_rn(m, kb[j % 6]));
        }
#pragma unroll
        for (int j = 0; j < 12; ++j)
            col[j / 3][j % 3] = byte_to_float(rw[j >> 2], 0x7650u + (j & 3));

        uint32_t *dst = reinterpret_cast<uint32_t *>(a.out + g * 12);
        const size_t dir_words = pixels * 3 / 4;
        for (unsigned d = 0; d < a.n_dirs; ++d) {
            const float4 la = *reinterpret_cast<const float4 *>(light_s + d * 8);
            const float l4 = light_s[d * 8 + 4];
            const float l[5] = {la.x, la.y, la.z, la.w, l4};
            unsigned b[12];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const float L = fmaxf(div255(poly_eval(u[px], l)), 0.0f);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    b[px * 3 + ch] = clip_hi_bits(__fmul_rn(L, col[px][ch]));
            }
            __stcs(dst + 0, pack4(b[0], b[1], b[2], b[3]));
            __stcs(dst + 1, pack4(b[4], b[5], b[6], b[7]));
            __stcs(dst + 2, pack4(b[8], b[9], b[10], b[11]));
            dst += dir_words;
        }
    }
}

// ---------------------------------------------------------------------------
// RGB formats, 4 consecutive pixels of an output row per thread (width % 4 == 0): three planes of
// 24 B each, 72 unscaled coefficients in registers, 12 bytes = 3 words out per direction
// (rti-builder/ptmlib.c:610-619: CLIP(POLY) per channel, both clamps needed).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kRelightThreads, 2)
relight_rgb_x4_kernel(RelightArgs a) {
    extern __shared__ float light_s[];  // [n_dirs][8]: u^2 v^2 uv u v (1) pad pad
    for (unsigned i = threadIdx.x; i < a.n_dirs * 8; i += blockDim.x)
        light_s[i] = (i & 7) < PTM_NCOEF ? a.light[(i >> 3) * PTM_NCOEF + (i & 7)] : 0.f;
    __syncthreads();

    const size_t pixels = a.width * a.height;
    const size_t groups = pixels / 4;
    const size_t gw = a.width / 4;
    float kb[PTM_NCOEF];
#pragma unroll
    for (int k = 0; k < PTM_NCOEF; ++k)
        kb[k] = 8388608.0f + (float) a.bias[k];
    for (size_t g = (size_t) blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (size_t) gridDim.x * blockDim.x) {
        const size_t y = g / gw, xg = g - y * gw;
        const size_t p = (a.height - 1 - y) * a.width + xg * 4;
        float u[3][4][PTM_NCOEF];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const uint2 *c2 = reinterpret_cast<const uint2 *>(a.plane[ch] + p * PTM_NCOEF);
            const uint2 cw[3] = {__ldg(c2), __ldg(c2 + 1), __ldg(c2 + 2)};
            const uint32_t cb[6] = {cw[0].x, cw[0].y, cw[1].x, cw[1].y, cw[2].x, cw[2].y};
#pragma unroll
            for (int j = 0; j < 24; ++j) {
                const float m = __uint_as_float(__byte_perm(cb[j >> 2], 0x4B000000u, 0x7650u + (j & 3)));
                u[ch][j / 6][j % 6] = __fmul_rn(a.scale[j % 6], __fsub_rn(m, kb[j % 6]));
            }
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(a.out + g * 12);
        const size_t dir_words = pixels * 3 / 4;
        for (unsigned d = 0; d < a.n_dirs; ++d) {
            const float4 la = *reinterpret_cast<const float4 *>(light_s + d * 8);
            const float l4 = light_s[d * 8 + 4];
            const float l[5] = {la.x, la.y, la.z, la.w, l4};
            unsigned b[12];
#pragma unroll
            for (int px = 0; px < 4; ++px)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    b[px * 3 + ch] = clip_bits(poly_eval(u[ch][px], l));
            __stcs(dst + 0, pack4(b[0], b[1], b[2], b[3]));
            __stcs(dst + 1, pack4(b[4], b[5], b[6], b[7]));
            __stcs(dst + 2, pack4(b[8], b[9], b[10], b[11]));
            dst += dir_words;
        }
    }
}

// ---------------------------------------------------------------------------
// Generic path: one pixel per thread, any width; LRGB or RGB planes.
// ---------------------------------------------------------------------------
template <bool kLrgb>
__global__ void __launch_bounds__(kRelightThreads)
relight_kernel(RelightArgs a) {
    extern __shared__ float light_s[];  // [n_dirs][6]
    for (unsigned i = threadIdx.x; i < a.n_dirs * PTM_NCOEF; i += blockDim.x)
        light_s[i] = a.light[i];
    __syncthreads();

    const size_t pixels = a.width * a.height;
    for (size_t o = (size_t) blockIdx.x * blockDim.x + threadIdx.x; o < pixels;
         o += (size_t) gridDim.x * blockDim.x) {
        const size_t y = o / a.width, x = o - y * a.width;
        const size_t p = (a.height - 1 - y) * a.width + x;  // stored pixel
        if (kLrgb) {
            float u[PTM_NCOEF];
            unscale6(a.plane[0] + p * PTM_NCOEF, a, u);
            const float r = (float) a.plane[1][p * 3], g = (float) a.plane[1][p * 3 + 1],
                        b = (float) a.plane[1][p * 3 + 2];
            for (unsigned d = 0; d < a.n_dirs; ++d) {
                const float L = fmaxf(div255(poly_eval(u, light_s + d * PTM_NCOEF)), 0.0f);
                uint8_t *dst = a.out + ((size_t) d * pixels + o) * 3;
                dst[0] = (uint8_t) clip_hi_bits(__fmul_rn(L, r));
                dst[1] = (uint8_t) clip_hi_bits(__fmul_rn(L, g));
                dst[2] = (uint8_t) clip_hi_bits(__fmul_rn(L, b));
            }
        } else {
            float u[3][PTM_NCOEF];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                unscale6(a.plane[ch] + p * PTM_NCOEF, a, u[ch]);
            for (unsigned d = 0; d < a.n_dirs; ++d) {
                uint8_t *dst = a.out + ((size_t) d * pixels + o) * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    dst[ch] = (uint8_t) clip_bits(poly_eval(u[ch], light_s + d * PTM_NCOEF));
            }
        }
    }
}

// Exhaustive check of div255() against the IEEE division, and of the CLIP replacement against
// the reference's definition, over float bit patterns [first, first + count).  Counts mismatches
// (NaN results compare equal to NaN).
__global__ void div255_selftest_kernel(unsigned long long first, unsigned long long count,
                                       unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long) gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned) (first + i));
        const float want = __fdiv_rn(x, 255.0f), got = div255(x);
        const bool same = (want == got) || (want != want && got != got);   // -0 == +0
        if (!same)
            ++bad;
        if (x == x) {  // CLIP is undefined for NaN in the reference
            const unsigned ref = x > 255.0f ? 255u : (x < 0.0f ? 0u : (unsigned) (unsigned char) x);
            if ((clip_bits(x) & 0xFFu) != ref)
                ++bad;
        }
    }
    if (bad)
        atomicAdd(mismatches, bad);
}

cudaError_t launch_div255_selftest(unsigned long long first, unsigned long long count,
                                   unsigned long long *mismatches_dev, int sm_count, cudaStream_t st) {
    div255_selftest_kernel<<<sm_count * 8, 256, 0, st>>>(first, count, mismatches_dev);
    return cudaGetLastError();
}

cudaError_t launch_relight(const RelightArgs &a, bool lrgb, int sm_count, cudaStream_t st) {
    const size_t pixels = a.width * a.height;
    if (pixels == 0 || a.n_dirs == 0)
        return cudaSuccess;
    bool small_bias = true;
    for (int k = 0; k < PTM_NCOEF; ++k)
        small_bias = small_bias && a.bias[k] > -(1 << 22) && a.bias[k] < (1 << 22);
    const bool x4 = lrgb && small_bias && a.width % 4 == 0 && ((uintptr_t) a.plane[0] % 8 == 0) &&
                    ((uintptr_t) a.plane[1] % 4 == 0) && ((uintptr_t) a.out % 4 == 0);
    if (x4) {
        size_t want = (pixels / 4 + kRelightThreads - 1) / kRelightThreads;
        size_t cap = (size_t) sm_count * 8;
        unsigned grid = (unsigned) (want < cap ? want : cap);
        relight_lrgb_x4_kernel<<<grid, kRelightThreads, (size_t) a.n_dirs * 8 * sizeof(float), st>>>(a);
        return cudaGetLastError();
    }
    static int x4_rgb = -1;
    if (x4_rgb < 0) {
        const char *v = getenv("PTM_RELIGHT_RGB_X4");   // tuning aid: 0 = one pixel per thread
        x4_rgb = v ? atoi(v) : 1;
    }
    const bool rgb4 = x4_rgb && !lrgb && small_bias && a.width % 4 == 0 && ((uintptr_t) a.plane[0] % 8 == 0) &&
                      ((uintptr_t) a.plane[1] % 8 == 0) && ((uintptr_t) a.plane[2] % 8 == 0) &&
                      ((uintptr_t) a.out % 4 == 0);
    if (rgb4) {
        size_t want4 = (pixels / 4 + kRelightThreads - 1) / kRelightThreads;
        size_t cap4 = (size_t) sm_count * 2;
        unsigned grid4 = (unsigned) (want4 < cap4 ? want4 : cap4);
        relight_rgb_x4_kernel<<<grid4, kRelightThreads, (size_t) a.n_dirs * 8 * sizeof(float), st>>>(a);
        return cudaGetLastError();
    }
    size_t want = (pixels + kRelightThreads - 1) / kRelightThreads;
    size_t cap = (size_t) sm_count * 8;
    unsigned grid = (unsigned) (want < cap ? want : cap);
    const size_t smem = (size_t) a.n_dirs * PTM_NCOEF * sizeof(float);
    if (lrgb)
        relight_kernel<true><<<grid, kRelightThreads, smem, st>>>(a);
    else
        relight_kernel<false><<<grid, kRelightThreads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ptm

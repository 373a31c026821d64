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
// Packed form of the LRGB kernel: the FP32-pipe instructions work on PAIRS of pixels
// (mul / fma / add.rz .f32x2, SASS FMUL2 / FFMA2 / FADD2.RZ -- IEEE results identical to the scalar
// instructions; the adds of the polynomial stay scalar, see add2).  A packed instruction occupies the FMA pipe for two cycles but only one issue
// slot, so the min/max, byte-permute and store instructions of the ALU / LSU pipes issue in its
// shadow; the scalar kernel is issue bound (tools/ubench_fma.cu measures the overlap).
// ---------------------------------------------------------------------------
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpk(f32x2 p, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// a + b per half with SCALAR adds: ptxas (12.9) contracts mul.rn.f32x2 followed by add.rn.f32x2 into
// FFMA2 even under --fmad false (it also sees through fma(t, 1, c) and fma(a, b, -0)), which would change
// the reference's unfused results; a packed multiply feeding scalar adds is left alone (checked in SASS).
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    float a0, a1, b0, b1;
    unpk(a, a0, a1);
    unpk(b, b0, b1);
    return pk(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
}
__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void __launch_bounds__(kRelightThreads)
relight_lrgb_x4p_kernel(RelightArgs a) {
    extern __shared__ __align__(16) float light_s[];  // [n_dirs][12]: (u^2,u^2) (v^2,v^2) (uv,uv) (u,u) (v,v) pad
    for (unsigned i = threadIdx.x; i < a.n_dirs * 12; i += blockDim.x) {
        const unsigned d = i / 12, j = (i % 12) >> 1;
        light_s[i] = j < 5 ? a.light[d * PTM_NCOEF + j] : 0.f;
    }
    __syncthreads();

    const size_t pixels = a.width * a.height;
    const size_t groups = pixels / 4;
    const size_t gw = a.width / 4;
    const f32x2 kR = pk(0.003921568859368562698364257812f, 0.003921568859368562698364257812f);  // RN(1 / 255)
    const f32x2 kM255 = pk(-255.0f, -255.0f);
    const f32x2 kMagic = pk(8388608.0f, 8388608.0f);
    for (size_t g = (size_t) blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (size_t) gridDim.x * blockDim.x) {
        const size_t y = g / gw, xg = g - y * gw;
        const size_t p = (a.height - 1 - y) * a.width + xg * 4;
        const uint2 *c2 = reinterpret_cast<const uint2 *>(a.plane[0] + p * PTM_NCOEF);
        const uint32_t *r32 = reinterpret_cast<const uint32_t *>(a.plane[1] + p * 3);
        const uint2 cw[3] = {__ldg(c2), __ldg(c2 + 1), __ldg(c2 + 2)};
        const uint32_t rw[3] = {__ldg(r32), __ldg(r32 + 1), __ldg(r32 + 2)};
        const uint32_t cb[6] = {cw[0].x, cw[0].y, cw[1].x, cw[1].y, cw[2].x, cw[2].y};
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
        // pixel pairs (0, 1) and (2, 3)
        f32x2 U[2][PTM_NCOEF], C[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                U[h][k] = pk(u[2 * h][k], u[2 * h + 1][k]);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                C[h][ch] = pk(col[2 * h][ch], col[2 * h + 1][ch]);
        }

        uint32_t *dst = reinterpret_cast<uint32_t *>(a.out + g * 12);
        const size_t dir_words = pixels * 3 / 4;
        for (unsigned d = 0; d < a.n_dirs; ++d) {
            const ulonglong2 *lp = reinterpret_cast<const ulonglong2 *>(light_s + d * 12);
            const ulonglong2 l01 = lp[0], l23 = lp[1];
            const f32x2 l4 = *reinterpret_cast<const f32x2 *>(light_s + d * 12 + 8);
            unsigned b[12];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // the reference's association (rti-builder/ptmlib.c:58-63), two pixels per instruction
                f32x2 s = mul2(U[h][0], l01.x);
                s = add2(s, mul2(U[h][1], l01.y));
                s = add2(s, mul2(U[h][2], l23.x));
                s = add2(s, mul2(U[h][3], l23.y));
                s = add2(s, mul2(U[h][4], l4));
                s = add2(s, U[h][5]);
                // s / 255.0f as in div255(): -q0 * 255 == q0 * -255 exactly
                const f32x2 q0 = mul2(s, kR);
                const f32x2 e = fma2(q0, kM255, s);
                const f32x2 q = fma2(e, kR, q0);
                float La, Lb;
                unpk(q, La, Lb);
                const f32x2 L = pk(fmaxf(La, 0.0f), fmaxf(Lb, 0.0f));
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float ta, tb;
                    unpk(mul2(L, C[h][ch]), ta, tb);
                    float za, zb;
                    unpk(add2_rz(pk(fminf(ta, 255.0f), fminf(tb, 255.0f)), kMagic), za, zb);
                    b[(2 * h) * 3 + ch] = __float_as_uint(za);
                    b[(2 * h + 1) * 3 + ch] = __float_as_uint(zb);
                }
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
// RGB formats, packed form: as relight_rgb_x4_kernel with the multiplies and the truncating add on
// pixel pairs.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kRelightThreads, 2)
relight_rgb_x4p_kernel(RelightArgs a) {
    extern __shared__ __align__(16) float light_s[];  // [n_dirs][12]: (u^2,u^2) (v^2,v^2) (uv,uv) (u,u) (v,v) pad
    for (unsigned i = threadIdx.x; i < a.n_dirs * 12; i += blockDim.x) {
        const unsigned d = i / 12, j = (i % 12) >> 1;
        light_s[i] = j < 5 ? a.light[d * PTM_NCOEF + j] : 0.f;
    }
    __syncthreads();

    const size_t pixels = a.width * a.height;
    const size_t groups = pixels / 4;
    const size_t gw = a.width / 4;
    const f32x2 kMagic = pk(8388608.0f, 8388608.0f);
    float kb[PTM_NCOEF];
#pragma unroll
    for (int k = 0; k < PTM_NCOEF; ++k)
        kb[k] = 8388608.0f + (float) a.bias[k];
    for (size_t g = (size_t) blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (size_t) gridDim.x * blockDim.x) {
        const size_t y = g / gw, xg = g - y * gw;
        const size_t p = (a.height - 1 - y) * a.width + xg * 4;
        f32x2 U[3][2][PTM_NCOEF];   // [channel][pixel pair][coefficient]
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const uint2 *c2 = reinterpret_cast<const uint2 *>(a.plane[ch] + p * PTM_NCOEF);
            const uint2 cw[3] = {__ldg(c2), __ldg(c2 + 1), __ldg(c2 + 2)};
            const uint32_t cb[6] = {cw[0].x, cw[0].y, cw[1].x, cw[1].y, cw[2].x, cw[2].y};
            float u[4][PTM_NCOEF];
#pragma unroll
            for (int j = 0; j < 24; ++j) {
                const float m = __uint_as_float(__byte_perm(cb[j >> 2], 0x4B000000u, 0x7650u + (j & 3)));
                u[j / 6][j % 6] = __fmul_rn(a.scale[j % 6], __fsub_rn(m, kb[j % 6]));
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < PTM_NCOEF; ++k)
                    U[ch][h][k] = pk(u[2 * h][k], u[2 * h + 1][k]);
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(a.out + g * 12);
        const size_t dir_words = pixels * 3 / 4;
        for (unsigned d = 0; d < a.n_dirs; ++d) {
            const ulonglong2 *lp = reinterpret_cast<const ulonglong2 *>(light_s + d * 12);
            const ulonglong2 l01 = lp[0], l23 = lp[1];
            const f32x2 l4 = *reinterpret_cast<const f32x2 *>(light_s + d * 12 + 8);
            unsigned b[12];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    f32x2 s = mul2(U[ch][h][0], l01.x);
                    s = add2(s, mul2(U[ch][h][1], l01.y));
                    s = add2(s, mul2(U[ch][h][2], l23.x));
                    s = add2(s, mul2(U[ch][h][3], l23.y));
                    s = add2(s, mul2(U[ch][h][4], l4));
                    s = add2(s, U[ch][h][5]);
                    float ta, tb, za, zb;
                    unpk(s, ta, tb);
                    unpk(add2_rz(pk(fminf(fmaxf(ta, 0.0f), 255.0f), fminf(fmaxf(tb, 0.0f), 255.0f)), kMagic), za, zb);
                    b[(2 * h) * 3 + ch] = __float_as_uint(za);
                    b[(2 * h + 1) * 3 + ch] = __float_as_uint(zb);
                }
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
        static int packed = -1;
        if (packed < 0) {
            const char *v = getenv("PTM_RELIGHT_PACKED");   // tuning aid: 0 = scalar FP32 instructions
            packed = v ? atoi(v) : 1;
        }
        if (packed && a.n_dirs >= 4)   // a single direction is memory bound: the scalar kernel has less set-up
            relight_lrgb_x4p_kernel<<<grid, kRelightThreads, (size_t) a.n_dirs * 12 * sizeof(float), st>>>(a);
        else
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
        static int packed3 = -1;
        if (packed3 < 0) {
            const char *v = getenv("PTM_RELIGHT_PACKED");
            packed3 = v ? atoi(v) : 1;
        }
        if (packed3 && a.n_dirs >= 4)
            relight_rgb_x4p_kernel<<<grid4, kRelightThreads, (size_t) a.n_dirs * 12 * sizeof(float), st>>>(a);
        else
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

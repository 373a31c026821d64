// The remaining per-pixel pieces of rti-builder's PTM path, one kernel each.  They are
// small, memory bound and off the headline path, so they are written plainly: one pixel per thread,
// coalesced accesses, the reference's arithmetic statement by statement (unfused, -fmad=false).
//
//   normal_map_kernel     ptm_normal (rti-builder/ptmlib.c:808-814, [Malzbender2001] eq. 16-17) for every
//                         pixel of a float coefficient plane, or of a quantised block (UNSCALE, :54, first)
//   relight_lum_kernel    the PTM_FORMAT_LUM branch of the relight loop (rti-builder/ptmlib.c:601-609)
//   fit_lsum_kernel       the encoder's disabled "L = R + G + B" LRGB branch (rti-builder/ptm-encoder.c:356-400)
//                         with the mean chroma it computes and, as the option its FIXME asks for (:383), the median
#include "ptm_device.cuh"
#include "ptm_kernels.h"

namespace ptm {

namespace {

constexpr int kThreads = 256;

// rti-builder/ptmlib.c:808-814.  The sign inside the square root is the reference's (1 - nu^2 + nv^2).
__device__ __forceinline__ void normal_of(const float *c, float &nu, float &nv, float &nw) {
    const float divisor = __fsub_rn(__fmul_rn(__fmul_rn(4.0f, c[0]), c[1]), __fmul_rn(c[2], c[2]));
    nu = __fdiv_rn(__fsub_rn(__fmul_rn(c[2], c[4]), __fmul_rn(__fmul_rn(2.0f, c[1]), c[3])), divisor);
    nv = __fdiv_rn(__fsub_rn(__fmul_rn(c[2], c[3]), __fmul_rn(__fmul_rn(2.0f, c[0]), c[4])), divisor);
    nw = __fsqrt_rn(__fadd_rn(__fsub_rn(1.0f, __fmul_rn(nu, nu)), __fmul_rn(nv, nv)));
}

// coeffs != nullptr: float plane [pixels][6]; else bytes [pixels][6] unscaled with scale / bias first.
__global__ void __launch_bounds__(kThreads)
normal_map_kernel(const float *__restrict__ coeffs, const uint8_t *__restrict__ bytes, NormalArgs a, size_t pixels,
                  float *__restrict__ normals) {
    for (size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (size_t) gridDim.x * blockDim.x) {
        float c[PTM_NCOEF];
        if (coeffs) {
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                c[k] = coeffs[p * PTM_NCOEF + k];
        } else {
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)   // UNSCALE: scale * (byte - bias), the subtraction in int
                c[k] = __fmul_rn(a.scale[k], (float) ((int) bytes[p * PTM_NCOEF + k] - a.bias[k]));
        }
        float nu, nv, nw;
        normal_of(c, nu, nv, nw);
        normals[p * 3 + 0] = nu;
        normals[p * 3 + 1] = nv;
        normals[p * 3 + 2] = nw;
    }
}

__device__ __forceinline__ unsigned clip_trunc(float f) {
    // rti-builder/ptmlib.h:156-158 (NaN: the comparisons fail and x86's conversion yields byte 0)
    return f > 255.0f ? 255u : (f < 0.0f ? 0u : (f == f ? (unsigned) (unsigned char) (int) f : 0u));
}

// rti-builder/ptmlib.c:601-609.  Block 1 is read as 2-byte (cr, cb) records at pixel index yoffs + x, exactly
// as the reference's pointer arithmetic on crcb_coefficients_t does; the output triplet is (Y, Cb, Cr).
__global__ void __launch_bounds__(kThreads)
relight_lum_kernel(RelightArgs a) {
    extern __shared__ float light_s[];  // [n_dirs][6]
    for (unsigned i = threadIdx.x; i < a.n_dirs * PTM_NCOEF; i += blockDim.x)
        light_s[i] = a.light[i];
    __syncthreads();
    const size_t pixels = a.width * a.height;
    for (size_t o = (size_t) blockIdx.x * blockDim.x + threadIdx.x; o < pixels; o += (size_t) gridDim.x * blockDim.x) {
        const size_t y = o / a.width, x = o - y * a.width;
        const size_t p = (a.height - 1 - y) * a.width + x;
        float u[PTM_NCOEF];
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            u[k] = __fmul_rn(a.scale[k], (float) ((int) a.plane[0][p * PTM_NCOEF + k] - a.bias[k]));
        const uint8_t cr = a.plane[1][p * 2], cb = a.plane[1][p * 2 + 1];
        for (unsigned d = 0; d < a.n_dirs; ++d) {
            const float *l = light_s + d * PTM_NCOEF;
            float s = __fmul_rn(u[0], l[0]);
            s = __fadd_rn(s, __fmul_rn(u[1], l[1]));
            s = __fadd_rn(s, __fmul_rn(u[2], l[2]));
            s = __fadd_rn(s, __fmul_rn(u[3], l[3]));
            s = __fadd_rn(s, __fmul_rn(u[4], l[4]));
            s = __fadd_rn(s, u[5]);
            uint8_t *dst = a.out + ((size_t) d * pixels + o) * 3;
            dst[0] = (uint8_t) clip_trunc(s);
            dst[1] = cb;
            dst[2] = cr;
        }
    }
}

// k-th smallest (0-based) of the n bytes of channel `ch` of pixel p across the lights: radix select from the
// top bit down, 8 passes; after the first pass the block's samples come from L1.
__device__ __forceinline__ unsigned select_byte(const uint8_t *stack, size_t image_stride, size_t off, unsigned n,
                                                unsigned k) {
    unsigned prefix = 0, mask = 0;
    for (int bit = 7; bit >= 0; --bit) {
        unsigned zeros = 0;
        for (unsigned i = 0; i < n; ++i) {
            const unsigned v = __ldg(stack + (size_t) i * image_stride + off);
            zeros += ((v & mask) == prefix) && !((v >> bit) & 1u);
        }
        if (k >= zeros) {
            k -= zeros;
            prefix |= 1u << bit;
        }
        mask |= 1u << bit;
    }
    return prefix;
}

// rti-builder/ptm-encoder.c:356-400 (disabled there: color_components == 666).  Per pixel:
//   L_n = R_n + G_n + B_n (:366-372), fitted like ptm_fit_poly_uint (:376-380): coefficients UNnormalised;
//   chroma_c = truncated mean over the lights (ptm_cbcr_avg, :384-386) or, kMedian, the median the FIXME at
//              :383 asks for ((lower + upper middle) / 2 for an even count);
//   colour_c = (JSAMPLE) (256 * (int) chroma_c / L_0) with L_0 the sum of the FIRST light (:391-399, `l`
//              walks L[0 .. pixels)).  L_0 == 0 divides by zero in the reference; 0 is stored here.
template <bool kMedian>
__global__ void __launch_bounds__(kThreads)
fit_lsum_kernel(FitArgs a) {
    extern __shared__ float m_gen[];  // [n][6]
    __shared__ int keys_s[2 * PTM_NCOEF];
    for (unsigned i = threadIdx.x; i < a.n_lights * PTM_NCOEF; i += blockDim.x)
        m_gen[i] = a.M[(size_t) (i % PTM_NCOEF) * a.n_lights + i / PTM_NCOEF];
    __syncthreads();
    Extrema ext;
    ext.init();
    const unsigned n = a.n_lights;
    for (size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x; p < a.pixels; p += (size_t) gridDim.x * blockDim.x) {
        float acc[PTM_NCOEF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        unsigned s[3] = {0, 0, 0}, l0 = 0;
        for (unsigned i = 0; i < n; ++i) {
            const uint8_t *t = a.stack + (size_t) i * a.image_stride + p * 3;
            const unsigned r = __ldg(t), g = __ldg(t + 1), b = __ldg(t + 2);
            const unsigned l = r + g + b;
            if (i == 0)
                l0 = l;
            const float y = (float) l;
#pragma unroll
            for (int k = 0; k < PTM_NCOEF; ++k)
                acc[k] = fmaf(m_gen[i * PTM_NCOEF + k], y, acc[k]);
            s[0] += r;
            s[1] += g;
            s[2] += b;
        }
        ext.add(acc);
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            a.coeffs[p * PTM_NCOEF + k] = acc[k];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            unsigned c;
            if (kMedian) {
                const unsigned lo = select_byte(a.stack, a.image_stride, p * 3 + ch, n, (n - 1) / 2);
                const unsigned hi = (n & 1) ? lo : select_byte(a.stack, a.image_stride, p * 3 + ch, n, n / 2);
                c = (lo + hi) / 2;
            } else {
                c = s[ch] / n;
            }
            a.rgb[p * 3 + ch] = l0 ? (uint8_t) (256u * c / l0) : (uint8_t) 0;
        }
    }
    ext.flush(a.keys, keys_s);
}

unsigned grid_for_px(size_t pixels, int sm_count, int per_sm) {
    const size_t want = (pixels + kThreads - 1) / kThreads;
    const size_t cap = (size_t) sm_count * per_sm;
    return (unsigned) (want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

cudaError_t launch_normal_map(const float *coeffs, const uint8_t *bytes, const NormalArgs &a, size_t pixels,
                              float *normals, int sm_count, cudaStream_t st) {
    if (pixels == 0)
        return cudaSuccess;
    normal_map_kernel<<<grid_for_px(pixels, sm_count, 8), kThreads, 0, st>>>(coeffs, bytes, a, pixels, normals);
    return cudaGetLastError();
}

cudaError_t launch_relight_lum(const RelightArgs &a, int sm_count, cudaStream_t st) {
    const size_t pixels = a.width * a.height;
    if (pixels == 0 || a.n_dirs == 0)
        return cudaSuccess;
    relight_lum_kernel<<<grid_for_px(pixels, sm_count, 8), kThreads, (size_t) a.n_dirs * PTM_NCOEF * sizeof(float), st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fit_lsum(const FitArgs &a, bool median, int sm_count, cudaStream_t st) {
    if (a.pixels == 0)
        return cudaSuccess;
    static bool allowed[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    const size_t smem = (size_t) a.n_lights * PTM_NCOEF * sizeof(float);
    if (dev >= 0 && dev < 64 && !allowed[dev]) {
        if ((e = cudaFuncSetAttribute(fit_lsum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 24)) != cudaSuccess ||
            (e = cudaFuncSetAttribute(fit_lsum_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 24)) != cudaSuccess)
            return e;
        allowed[dev] = true;
    }
    if (median)
        fit_lsum_kernel<true><<<grid_for_px(a.pixels, sm_count, 8), kThreads, smem, st>>>(a);
    else
        fit_lsum_kernel<false><<<grid_for_px(a.pixels, sm_count, 8), kThreads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ptm

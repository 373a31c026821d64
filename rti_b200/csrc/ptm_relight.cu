// K3 -- relighting: the pixel loop of ptm_write_jpeg (rti-builder/ptmlib.c:553-562,
// 587-621) evaluated for a batch of light directions.
//
// POLY / UNSCALE (rti-builder/ptmlib.c:54-63): (scale_k * (float)(byte_k - bias_k)) is
// direction independent, so it is computed once per pixel; per direction the five
// products are added left to right with separate multiplies and adds, exactly the
// reference's association.  LRGB: L = POLY / 255.0f (a true division, :595), then
// CLIP(L * colour byte).  Output row y reads stored row height-1-y (:587-589).
#include "ptm_device.cuh"
#include "ptm_kernels.h"

namespace ptm {

constexpr int kRelightThreads = 256;

__device__ __forceinline__ float poly_eval(const float *a, const float *l) {
    float s = __fmul_rn(a[0], l[0]);
    s = __fadd_rn(s, __fmul_rn(a[1], l[1]));
    s = __fadd_rn(s, __fmul_rn(a[2], l[2]));
    s = __fadd_rn(s, __fmul_rn(a[3], l[3]));
    s = __fadd_rn(s, __fmul_rn(a[4], l[4]));
    return __fadd_rn(s, a[5]);
}

__device__ __forceinline__ void unscale6(const uint8_t *c, const RelightArgs &a, float *out) {
#pragma unroll
    for (int k = 0; k < PTM_NCOEF; ++k)
        out[k] = __fmul_rn(a.scale[k], (float) ((int) c[k] - a.bias[k]));
}

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
                const float L = __fdiv_rn(poly_eval(u, light_s + d * PTM_NCOEF), 255.0f);
                uint8_t *dst = a.out + ((size_t) d * pixels + o) * 3;
                dst[0] = (uint8_t) clip_u8(__fmul_rn(L, r));
                dst[1] = (uint8_t) clip_u8(__fmul_rn(L, g));
                dst[2] = (uint8_t) clip_u8(__fmul_rn(L, b));
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
                    dst[ch] = (uint8_t) clip_u8(poly_eval(u[ch], light_s + d * PTM_NCOEF));
            }
        }
    }
}

cudaError_t launch_relight(const RelightArgs &a, bool lrgb, int sm_count, cudaStream_t st) {
    const size_t pixels = a.width * a.height;
    if (pixels == 0 || a.n_dirs == 0)
        return cudaSuccess;
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

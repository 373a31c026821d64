// Synthetic stack generator (device side).  Integer-only and counter based: the
// triplet of (pixel p, light n) depends only on (seed, p, n, the light's Q14
// direction), so every GPU can generate its own row slab and the host statement
// (oracle/synth.c) reproduces the same bytes for the parity tests.
//
// Shape: pseudo-random normal and albedo per pixel,
// Lambertian shading, a little per-sample noise, clamped to >= 1 so the averaged
// luma never hits the reference's 256/0 (rti-builder/ptm-encoder.c:342).
#include "ptm_device.cuh"
#include "ptm_kernels.h"

namespace ptm {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <typename T>
__global__ void __launch_bounds__(256)
synth_kernel(uint32_t seed, int mode, uint32_t pixel0, size_t pixels, unsigned n_lights,
             const int32_t *__restrict__ lq, uint8_t *stack, size_t image_stride) {
    constexpr bool k16 = sizeof(T) == 2;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < pixels;
         i += (size_t) gridDim.x * blockDim.x) {
        const uint32_t p = pixel0 + (uint32_t) i;
        const uint32_t h1 = mix32(seed ^ mix32(p));
        const uint32_t h2 = mix32(h1 + 0x9E3779B9U);
        const int nx = (((int) (h1 & 0xFFFFU) - 32768) * 3) / 8;
        const int ny = (((int) (h1 >> 16) - 32768) * 3) / 8;
        const int nz = 32768 - ((nx * nx + ny * ny) >> 16);
        int amp[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            amp[ch] = 60 + (int) (((h2 >> (8 * ch)) & 0xFFU) * 3U) / 4;
        for (unsigned n = 0; n < n_lights; ++n) {
            const uint32_t h3 = mix32(h2 ^ (n * 0x85EBCA6BU + 0xC2B2AE35U));
            const int dot = nx * lq[3 * n] + ny * lq[3 * n + 1] + nz * lq[3 * n + 2];
            const int shade = dot > 0 ? dot >> 14 : 0;
            T *t = reinterpret_cast<T *>(stack + (size_t) n * image_stride) + i * 3;
            if (!k16) {
                if (mode == 0) {
                    const int y = ((shade * amp[0]) >> 15) + (int) ((h3 & 0xFFU) % 7U) - 3;
                    t[0] = (T) clampi(y, 1, 255);
                    t[1] = (T) (128 + ((int) ((h2 >> 8) & 63U) - 32) + ((int) ((h3 >> 8) & 3U) - 1));
                    t[2] = (T) (128 + ((int) ((h2 >> 16) & 63U) - 32) + ((int) ((h3 >> 16) & 3U) - 1));
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const int c = ((shade * amp[ch]) >> 15) + (int) (((h3 >> (8 * ch)) & 0xFFU) % 7U) - 3;
                        t[ch] = (T) clampi(c, 1, 255);
                    }
                }
            } else {
                if (mode == 0) {
                    const int y = ((shade * amp[0]) >> 7) + (int) (h3 & 0x3FFU) - 512;
                    t[0] = (T) clampi(y, 1, 65535);
                    t[1] = (T) (32768 + (((int) ((h2 >> 8) & 63U) - 32) << 8) + ((int) ((h3 >> 10) & 0xFFU) - 128));
                    t[2] = (T) (32768 + (((int) ((h2 >> 16) & 63U) - 32) << 8) + ((int) ((h3 >> 18) & 0xFFU) - 128));
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const int c = ((shade * amp[ch]) >> 7) + (int) ((h3 >> (10 * ch)) & 0x3FFU) - 512;
                        t[ch] = (T) clampi(c, 1, 65535);
                    }
                }
            }
        }
    }
}

cudaError_t launch_synth(uint32_t seed, int mode, int sample_type, uint32_t pixel0, size_t pixels,
                         unsigned n_lights, const int32_t *lights_q14_dev, void *stack, size_t image_stride,
                         int sm_count, cudaStream_t st) {
    if (pixels == 0)
        return cudaSuccess;
    size_t want = (pixels + 255) / 256;
    size_t cap = (size_t) sm_count * 8;
    unsigned grid = (unsigned) (want < cap ? want : cap);
    if (sample_type == 0)
        synth_kernel<uint8_t><<<grid, 256, 0, st>>>(seed, mode, pixel0, pixels, n_lights, lights_q14_dev,
                                                    (uint8_t *) stack, image_stride);
    else if (sample_type == 1)
        synth_kernel<uint16_t><<<grid, 256, 0, st>>>(seed, mode, pixel0, pixels, n_lights, lights_q14_dev,
                                                     (uint8_t *) stack, image_stride);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace ptm

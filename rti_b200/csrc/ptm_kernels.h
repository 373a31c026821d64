// Internal: kernel argument blocks and launchers shared by the .cu files.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace ptm {

constexpr int kFitThreads = 256;

struct FitArgs {
    const uint8_t *stack;   // [n_lights][pixels][3] samples, device
    size_t image_stride;    // bytes between two lights' images
    size_t pixels;
    unsigned n_lights;
    unsigned magic;         // floor(2^32 / n_lights) + 1, 0 when n_lights == 1
    const float *M;         // [6][n_lights], device
    float *coeffs;          // [pixels][6] (LRGB) or 3 planes (RGB); may be NULL for LRGB
    size_t plane_stride;    // floats between RGB coefficient planes
    uint8_t *rgb;           // [pixels][3] (LRGB only)
    int *keys;              // [12]
};

struct RelightArgs {
    const uint8_t *plane[3];  // LRGB: coef, rgb, unused; RGB: r, g, b coefficient planes
    size_t width, height;
    float scale[6];
    int bias[6];
    const float *light;       // device [n_dirs][6]: u^2 v^2 uv u v 1
    unsigned n_dirs;
    uint8_t *out;             // [n_dirs][height][width][3]
};

cudaError_t launch_keys_init(int *keys, cudaStream_t st);
cudaError_t launch_fit_lrgb(const FitArgs &a, int sample_type, int sm_count, cudaStream_t st);
cudaError_t launch_fit_lrgb_tma(const FitArgs &a, int sm_count, cudaStream_t st, size_t *done);
cudaError_t launch_fit_rgb(const FitArgs &a, int sample_type, int sm_count, cudaStream_t st);
cudaError_t launch_fit_channel(const void *samples, int sample_type, size_t pixel_stride, size_t image_stride,
                               size_t pixels, unsigned n_lights, const float *M, float *coeffs, int sm_count,
                               cudaStream_t st);
cudaError_t launch_chroma_avg(const uint8_t *stack, size_t image_stride, size_t pixels, unsigned n_lights,
                              uint8_t *avg, int sm_count, cudaStream_t st);
cudaError_t launch_lrgb_colour_normalise(uint8_t *ycc, float *coeffs, size_t pixels, int sm_count,
                                         cudaStream_t st);
cudaError_t launch_minmax(const float *coeffs, size_t pixels, int *keys, int sm_count, cudaStream_t st);
cudaError_t launch_quantise(const float *coeffs, size_t pixels, const int *keys, uint8_t *bytes, void *sb,
                            bool consume, int sm_count, cudaStream_t st);
cudaError_t launch_relight(const RelightArgs &a, bool lrgb, int sm_count, cudaStream_t st);
cudaError_t launch_div255_selftest(unsigned long long first, unsigned long long count,
                                   unsigned long long *mismatches_dev, int sm_count, cudaStream_t st);
cudaError_t launch_synth(uint32_t seed, int mode, int sample_type, uint32_t pixel0, size_t pixels,
                         unsigned n_lights, const int32_t *lights_q14_dev, void *stack, size_t image_stride,
                         int sm_count, cudaStream_t st);

}  // namespace ptm

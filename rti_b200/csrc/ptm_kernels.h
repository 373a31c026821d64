// Internal: kernel argument blocks and launchers shared by the .cu files.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

namespace ptm {

constexpr int kFitThreads = 256;

// Mailbox used to hand the extrema keys from the fit to the quantiser, on one GPU or across
// GPUs (ptm_xchg.cu): slots[2][kMaxRanks][16] ints, flags[2][kMaxRanks] ints, status[4] ints.
constexpr int kMaxRanks = 16;
constexpr int kSlotInts = 16;  // 12 keys, padded to 64 bytes
struct XchgLayout {
    static constexpr size_t slots_off = 0;
    static constexpr size_t flags_off = 2 * kMaxRanks * kSlotInts * sizeof(int);
    static constexpr size_t status_off = flags_off + 2 * kMaxRanks * sizeof(int);
    static constexpr size_t bytes = status_off + 4 * sizeof(int);
};

// What the LAST block of a fit launch does with the merged keys: deliver them to the mailbox of
// every rank (its own included) under `epoch`, and reset keys + ticket for the next launch.
struct KeyHandoff {
    unsigned *ticket;           // device counter, 0 between launches; nullptr = no hand-off
    unsigned *tile_counter;     // K1's dynamic tile counter, reset by the last block (may be nullptr)
    int *mailbox[kMaxRanks];
    int world, rank;
    unsigned epoch;
};

// Where K2 takes the extrema from: plain keys, or the local mailbox (waits for all ranks' flags).
struct KeySource {
    const int *keys;
    int *mailbox;               // nullptr = use keys
    int world;
    unsigned epoch;
};

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
    KeyHandoff handoff;     // optional (ticket == nullptr: none)
    // tensor-core engine (ptm_fit_mma.cu): fixed-point digits of M as the B operand image, device
    const int8_t *digits;   // nullptr = not available for this matrix
    float digit_scale[6];   // 2^-F_k
    unsigned kpad;          // lights rounded up to 32
    unsigned *tile_counter; // dynamic tile scheduling of the LRGB ring kernel: device counter, 0 at launch (nullptr = static)
    // L2 policy of K1's coefficient stores: 0 = plain .cg stores, 1 = evict_first, 2 = evict_last (default: the
    // float plane is what K2 reads next; kept with priority, part of it is still in L2 then)
    int store_hint;
    int engine;             // PTMB_ENGINE_*: 0 = default choice, 1 = FFMA kernels, 2 = tensor-core kernel
    // LRGB-shaped fits only: kFitLum = PTM_FORMAT_LUM (no colour conversion, no normalisation: the colour block
    // is the averaged YCbCr triplet), kFitRgbIn = the stack holds RGB triplets that are converted to YCbCr per
    // sample (libjpeg's integer formula) before anything else.  Both are served by the FFMA kernels.
    int mode;
    // the caller guarantees that nothing queued earlier on the stream writes the stack: the ring kernel's
    // producer may then start copying before the programmatic dependency on the previous grid has resolved
    int stack_stable;
};
constexpr int kFitLum = 1;
constexpr int kFitRgbIn = 2;

struct RelightArgs {
    const uint8_t *plane[3];  // LRGB: coef, rgb, unused; RGB: r, g, b coefficient planes
    size_t width, height;
    float scale[6];
    int bias[6];
    const float *light;       // device [n_dirs][6]: u^2 v^2 uv u v 1
    unsigned n_dirs;
    uint8_t *out;             // [n_dirs][height][width][3]
};

struct NormalArgs {
    float scale[6];
    int bias[6];
};
cudaError_t launch_normal_map(const float *coeffs, const uint8_t *bytes, const NormalArgs &a, size_t pixels,
                              float *normals, int sm_count, cudaStream_t st);
cudaError_t launch_relight_lum(const RelightArgs &a, int sm_count, cudaStream_t st);
cudaError_t launch_relight_rgb_mma(const RelightArgs &a, float max_term, int sm_count, cudaStream_t st);
cudaError_t launch_fit_lsum(const FitArgs &a, bool median, int sm_count, cudaStream_t st);

bool pack_mma_digits(const float *M, unsigned n_lights, std::vector<int8_t> &image, float *scale, unsigned *kpad_out);
// per-format default engines (1 = FFMA kernels, 2 = tensor-core kernel), see DESIGN.md
constexpr int kDefaultEngineLrgbU8 = 1;
constexpr int kDefaultEngineRgbU8 = 2;
constexpr int kDefaultEngineRgbU16 = 2;
constexpr int kDefaultEngineLrgbU16 = 1;
int resolve_engine(int requested, int format_default);
cudaError_t launch_keys_init(int *keys, cudaStream_t st);
cudaError_t launch_fit_lrgb(const FitArgs &a, int sample_type, int sm_count, cudaStream_t st);
cudaError_t launch_fit_lrgb_tma(const FitArgs &a, int sm_count, cudaStream_t st, size_t *done);
cudaError_t launch_fit_u16_tma(const FitArgs &a, bool rgb, int sm_count, cudaStream_t st, size_t *done);
unsigned long long mma_launch_count();
cudaError_t launch_fit_mma(const FitArgs &a, int mode, int sm_count, cudaStream_t st, size_t *done);
cudaError_t launch_fit_rgb_tma(const FitArgs &a, int sm_count, cudaStream_t st, size_t *done);
cudaError_t launch_fit_rgb(const FitArgs &a, int sample_type, int sm_count, cudaStream_t st);
cudaError_t launch_fit_channel(const void *samples, int sample_type, size_t pixel_stride, size_t image_stride,
                               size_t pixels, unsigned n_lights, const float *M, float *coeffs, int sm_count,
                               cudaStream_t st);
cudaError_t launch_chroma_avg(const uint8_t *stack, size_t image_stride, size_t pixels, unsigned n_lights,
                              uint8_t *avg, int sm_count, cudaStream_t st);
cudaError_t launch_lrgb_colour_normalise(uint8_t *ycc, float *coeffs, size_t pixels, int sm_count,
                                         cudaStream_t st);
cudaError_t launch_minmax(const float *coeffs, size_t pixels, int *keys, int sm_count, cudaStream_t st);
cudaError_t launch_quantise(const float *coeffs, size_t pixels, const KeySource &src, uint8_t *bytes, void *sb,
                            bool consume, int sm_count, cudaStream_t st);
cudaError_t launch_relight(const RelightArgs &a, bool lrgb, int sm_count, cudaStream_t st);
// tolerance form (ptm_relight_fast.cu); cudaErrorNotSupported = layout not taken, use launch_relight
cudaError_t launch_relight_fast(const RelightArgs &a, bool lrgb, int sm_count, cudaStream_t st);
cudaError_t launch_div255_selftest(unsigned long long first, unsigned long long count,
                                   unsigned long long *mismatches_dev, int sm_count, cudaStream_t st);
cudaError_t launch_synth(uint32_t seed, int mode, int sample_type, uint32_t pixel0, size_t pixels,
                         unsigned n_lights, const int32_t *lights_q14_dev, void *stack, size_t image_stride,
                         int sm_count, cudaStream_t st);

}  // namespace ptm

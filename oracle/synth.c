/*
 * synth.c -- host-side statement of the synthetic stack generator
 * (TEST INFRASTRUCTURE; the device-side generator the benchmark uses lives in
 * rti_b200/csrc/ptm_synth.cu and must produce the same bytes -- see
 * tests/test_gpu_parity.py::test_synth_matches_host).
 *
 * Integer-only, counter-based: sample (pixel p, light n) depends on nothing
 * but (seed, p, n, the light's fixed-point direction), so any row slab can be
 * generated independently on any GPU or on the host.
 *
 * Shape: a pseudo-random surface normal and albedo per
 * pixel, Lambertian shading under light n, a little per-sample noise, values
 * clamped to [1, max] so the averaged luma is never 0 (the reference divides by
 * it, rti-builder/ptm-encoder.c:342).
 *
 *   mode 0 (YCbCr): byte 0 = shaded luma, bytes 1, 2 = per-pixel chroma near 128
 *   mode 1 (RGB)  : three shaded channels with per-channel albedo
 */
#include <stddef.h>
#include <stdint.h>

static inline uint32_t mix32 (uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

static inline int clampi (int v, int lo, int hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

/* Q14 light direction from floats, round to nearest (ties away from zero). */
void synth_light_q14 (const float *uvw, int n, int32_t *q) {
    for (int i = 0; i < 3 * n; ++i) {
        float s = uvw[i] * 16384.0f;
        q[i] = (int32_t) (s >= 0.0f ? s + 0.5f : s - 0.5f);
    }
}

static inline void synth_pixel_state (uint32_t seed, uint32_t p, uint32_t *h1, uint32_t *h2,
                                      int *nx, int *ny, int *nz) {
    *h1 = mix32 (seed ^ mix32 (p));
    *h2 = mix32 (*h1 + 0x9E3779B9U);
    *nx = (((int) (*h1 & 0xFFFFU) - 32768) * 3) / 8;
    *ny = (((int) (*h1 >> 16) - 32768) * 3) / 8;
    *nz = 32768 - ((*nx * *nx + *ny * *ny) >> 16);
}

static inline int synth_shade (int nx, int ny, int nz, const int32_t *l) {
    int dot = nx * l[0] + ny * l[1] + nz * l[2];
    return dot > 0 ? dot >> 14 : 0;
}

static inline int synth_amp (uint32_t h2, int ch) {
    return 60 + (int) (((h2 >> (8 * ch)) & 0xFFU) * 3U) / 4;
}

/* 8-bit triplets for pixels [p0, p0+count) of light n. */
void synth_u8 (uint32_t seed, int mode, uint32_t p0, size_t count, uint32_t n,
               const int32_t *light_q14, uint8_t *out) {
    const int32_t *l = light_q14 + 3 * (size_t) n;
    for (size_t i = 0; i < count; ++i) {
        uint32_t p = p0 + (uint32_t) i, h1, h2;
        int nx, ny, nz;
        synth_pixel_state (seed, p, &h1, &h2, &nx, &ny, &nz);
        uint32_t h3 = mix32 (h2 ^ (n * 0x85EBCA6BU + 0xC2B2AE35U));
        int shade = synth_shade (nx, ny, nz, l);
        uint8_t *t = out + i * 3;
        if (mode == 0) {
            int y = ((shade * synth_amp (h2, 0)) >> 15) + (int) ((h3 & 0xFFU) % 7U) - 3;
            t[0] = (uint8_t) clampi (y, 1, 255);
            t[1] = (uint8_t) (128 + ((int) ((h2 >> 8) & 63U) - 32) + ((int) ((h3 >> 8) & 3U) - 1));
            t[2] = (uint8_t) (128 + ((int) ((h2 >> 16) & 63U) - 32) + ((int) ((h3 >> 16) & 3U) - 1));
        } else {
            for (int ch = 0; ch < 3; ++ch) {
                int c = ((shade * synth_amp (h2, ch)) >> 15)
                      + (int) (((h3 >> (8 * ch)) & 0xFFU) % 7U) - 3;
                t[ch] = (uint8_t) clampi (c, 1, 255);
            }
        }
    }
}

/* 16-bit linear triplets (the RAW-derived configuration), same shape. */
void synth_u16 (uint32_t seed, int mode, uint32_t p0, size_t count, uint32_t n,
                const int32_t *light_q14, uint16_t *out) {
    const int32_t *l = light_q14 + 3 * (size_t) n;
    for (size_t i = 0; i < count; ++i) {
        uint32_t p = p0 + (uint32_t) i, h1, h2;
        int nx, ny, nz;
        synth_pixel_state (seed, p, &h1, &h2, &nx, &ny, &nz);
        uint32_t h3 = mix32 (h2 ^ (n * 0x85EBCA6BU + 0xC2B2AE35U));
        int shade = synth_shade (nx, ny, nz, l);
        uint16_t *t = out + i * 3;
        if (mode == 0) {
            int y = ((shade * synth_amp (h2, 0)) >> 7) + (int) (h3 & 0x3FFU) - 512;
            t[0] = (uint16_t) clampi (y, 1, 65535);
            t[1] = (uint16_t) (32768 + (((int) ((h2 >> 8) & 63U) - 32) << 8) + ((int) ((h3 >> 10) & 0xFFU) - 128));
            t[2] = (uint16_t) (32768 + (((int) ((h2 >> 16) & 63U) - 32) << 8) + ((int) ((h3 >> 18) & 0xFFU) - 128));
        } else {
            for (int ch = 0; ch < 3; ++ch) {
                int c = ((shade * synth_amp (h2, ch)) >> 7)
                      + (int) ((h3 >> (10 * ch)) & 0x3FFU) - 512;
                t[ch] = (uint16_t) clampi (c, 1, 65535);
            }
        }
    }
}

/* Whole stack [n_lights][pixels][3] for pixels [p0, p0+pixels). */
void synth_stack_u8 (uint32_t seed, int mode, uint32_t p0, size_t pixels, uint32_t n_lights,
                     const int32_t *light_q14, uint8_t *out) {
    #pragma omp parallel for schedule(static)
    for (uint32_t n = 0; n < n_lights; ++n)
        synth_u8 (seed, mode, p0, pixels, n, light_q14, out + (size_t) n * pixels * 3);
}

void synth_stack_u16 (uint32_t seed, int mode, uint32_t p0, size_t pixels, uint32_t n_lights,
                      const int32_t *light_q14, uint16_t *out) {
    #pragma omp parallel for schedule(static)
    for (uint32_t n = 0; n < n_lights; ++n)
        synth_u16 (seed, mode, p0, pixels, n, light_q14, out + (size_t) n * pixels * 3);
}

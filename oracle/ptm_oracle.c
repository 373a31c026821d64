/*
 * ptm_oracle.c -- CPU restatement of the cceh/rti PTM fit / quantise / relight
 * arithmetic, used ONLY as the parity checker for the CUDA path.
 *
 * TEST INFRASTRUCTURE.  Nothing under rti_b200/ may call, link or import this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and there only as the checker or the timed CPU
 * baseline.  The product path has no CPU fallback.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 8c).  This restatement is therefore pinned by (1) the one
 * numeric example in the reference's documentation
 * (doc_src/draw_light_illustrations.m:1-59, see tests/golden/octave_kat.json)
 * and (2) byte-for-byte comparison against the reference's own ptmlib.c and
 * command-line mains, compiled unmodified into oracle/_ref/ by oracle/Makefile
 * and run in the build container (tests/test_oracle.py::test_oracle_vs_reference_live).
 *
 * Every function cites the reference lines it follows.  Build flags mirror the
 * reference's (rti-builder/Makefile:5-7: gcc -O -std=c11 -fopenmp, no -march,
 * no -ffast-math), so a*b+c is two roundings here exactly as it is there.
 *
 * All paths are relative to /root/reference/.
 */

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cblas.h>
#include <lapacke.h>

#define NCOEF 6

/* rti-builder/ptmlib.h:156-158 -- clip, then C float->uchar truncation. */
static inline uint8_t clip_u8 (float f) {
    if (f > 255.0f)
        return 255;
    if (f < 0.0f)
        return 0;
    return (uint8_t) f;
}

/* ------------------------------------------------------------------ pinv
 * rti-builder/ptmlib.c:630-689 (ptm_svd).  Row n of the design matrix is
 * [u^2 v^2 uv u v 1] (:640-651); thin SVD through LAPACKE_sgesdd (:657-661);
 * M = V * diag(1/s) * U' by two sgemm calls (:670-684), 6 x n row major.
 * Returns 0 on success, 1 when the SVD did not converge (:662-664).
 */
int orc_pinv (const float *u, const float *v, int n, float *M) {
    float *A = calloc ((size_t) n * NCOEF, sizeof (float));
    float *U = calloc ((size_t) n * NCOEF, sizeof (float));
    float S[NCOEF * NCOEF] = { 0 };
    float VT[NCOEF * NCOEF] = { 0 };
    float sv[NCOEF] = { 0 };
    float M1[NCOEF * NCOEF] = { 0 };

    for (int i = 0; i < n; ++i) {
        float *a = A + (size_t) i * NCOEF;
        a[0] = u[i] * u[i];
        a[1] = v[i] * v[i];
        a[2] = u[i] * v[i];
        a[3] = u[i];
        a[4] = v[i];
        a[5] = 1.0f;
    }
    lapack_int info = LAPACKE_sgesdd (LAPACK_ROW_MAJOR, 'S', n, NCOEF, A, NCOEF,
                                      sv, U, NCOEF, VT, NCOEF);
    if (info > 0) {
        free (A);
        free (U);
        return 1;
    }
    for (int i = 0; i < NCOEF; ++i)
        S[i * NCOEF + i] = 1.0f / sv[i];
    cblas_sgemm (CblasRowMajor, CblasTrans, CblasNoTrans, NCOEF, NCOEF, NCOEF,
                 1.0f, VT, NCOEF, S, NCOEF, 0.0f, M1, NCOEF);
    cblas_sgemm (CblasRowMajor, CblasNoTrans, CblasTrans, NCOEF, n, NCOEF,
                 1.0f, M1, NCOEF, U, NCOEF, 0.0f, M, n);
    free (A);
    free (U);
    return 0;
}

/* Singular values only (for the KAT in SURVEY.md section 4). */
int orc_singular_values (const float *A_in, int n, float *sv) {
    float *A = malloc ((size_t) n * NCOEF * sizeof (float));
    float *U = calloc ((size_t) n * NCOEF, sizeof (float));
    float VT[NCOEF * NCOEF];
    memcpy (A, A_in, (size_t) n * NCOEF * sizeof (float));
    lapack_int info = LAPACKE_sgesdd (LAPACK_ROW_MAJOR, 'S', n, NCOEF, A, NCOEF,
                                      sv, U, NCOEF, VT, NCOEF);
    free (A);
    free (U);
    return info > 0;
}

/* Same as orc_pinv but from an explicit n x 6 design matrix (the Octave
 * example in doc_src/draw_light_illustrations.m gives A, not u and v). */
int orc_pinv_from_design (const float *A_in, int n, float *M) {
    float *A = malloc ((size_t) n * NCOEF * sizeof (float));
    float *U = calloc ((size_t) n * NCOEF, sizeof (float));
    float S[NCOEF * NCOEF] = { 0 };
    float VT[NCOEF * NCOEF] = { 0 };
    float sv[NCOEF] = { 0 };
    float M1[NCOEF * NCOEF] = { 0 };
    memcpy (A, A_in, (size_t) n * NCOEF * sizeof (float));
    lapack_int info = LAPACKE_sgesdd (LAPACK_ROW_MAJOR, 'S', n, NCOEF, A, NCOEF,
                                      sv, U, NCOEF, VT, NCOEF);
    if (info > 0) {
        free (A);
        free (U);
        return 1;
    }
    for (int i = 0; i < NCOEF; ++i)
        S[i * NCOEF + i] = 1.0f / sv[i];
    cblas_sgemm (CblasRowMajor, CblasTrans, CblasNoTrans, NCOEF, NCOEF, NCOEF,
                 1.0f, VT, NCOEF, S, NCOEF, 0.0f, M1, NCOEF);
    cblas_sgemm (CblasRowMajor, CblasNoTrans, CblasTrans, NCOEF, n, NCOEF,
                 1.0f, M1, NCOEF, U, NCOEF, 0.0f, M, n);
    free (A);
    free (U);
    return 0;
}

/* ------------------------------------------------------------------- fit
 * rti-builder/ptmlib.c:692-725 (ptm_fit_poly_jsample): for each pixel gather
 * the n samples (stride = one whole image, :712-715) and compute M * b
 * (:717-720).  The reference leaves the summation order to its BLAS; this
 * restatement sums in ascending light order with separate multiply and add.
 * `samples` may be offset by a channel, `pixel_stride` is in samples.
 */
void orc_fit_u8 (const uint8_t *samples, size_t width, size_t height,
                 unsigned n_lights, size_t pixel_stride, const float *M, float *out) {
    const size_t row_stride = width * pixel_stride;
    const size_t image_stride = height * row_stride;

    #pragma omp parallel for schedule(dynamic)
    for (size_t y = 0; y < height; ++y) {
        float *o = out + y * width * NCOEF;
        for (size_t x = 0; x < width; ++x, o += NCOEF) {
            const uint8_t *s = samples + y * row_stride + x * pixel_stride;
            float acc[NCOEF] = { 0 };
            for (unsigned n = 0; n < n_lights; ++n) {
                float b = (float) s[(size_t) n * image_stride];
                for (int k = 0; k < NCOEF; ++k)
                    acc[k] = acc[k] + M[(size_t) k * n_lights + n] * b;
            }
            memcpy (o, acc, sizeof (acc));
        }
    }
}

/* The same fit, but with the per-pixel product handed to cblas_sgemv exactly
 * as the reference does (rti-builder/ptmlib.c:707-720).  Bit-identical to the
 * reference on the same machine and BLAS build; used to show that everything
 * around the dot product in this restatement is exact, and what the BLAS
 * summation order alone changes. */
void orc_fit_u8_blas (const uint8_t *samples, size_t width, size_t height,
                      unsigned n_lights, size_t pixel_stride, const float *M, float *out) {
    const size_t row_stride = width * pixel_stride;
    const size_t image_stride = height * row_stride;

    #pragma omp parallel for schedule(dynamic)
    for (size_t y = 0; y < height; ++y) {
        float *b = calloc (n_lights, sizeof (float));
        float *o = out + y * width * NCOEF;
        for (size_t x = 0; x < width; ++x, o += NCOEF) {
            const uint8_t *s = samples + y * row_stride + x * pixel_stride;
            for (unsigned n = 0; n < n_lights; ++n)
                b[n] = (float) s[(size_t) n * image_stride];
            cblas_sgemv (CblasRowMajor, CblasNoTrans, NCOEF, (int) n_lights,
                         1.0f, M, (int) n_lights, b, 1, 0.0f, o, 1);
        }
        free (b);
    }
}

/* rti-builder/ptmlib.c:727-760 (ptm_fit_poly_uint): the same for unsigned int
 * samples -- the reference's only entry point for data wider than 8 bits. */
void orc_fit_u32 (const uint32_t *samples, size_t width, size_t height,
                  unsigned n_lights, size_t pixel_stride, const float *M, float *out) {
    const size_t row_stride = width * pixel_stride;
    const size_t image_stride = height * row_stride;

    #pragma omp parallel for schedule(dynamic)
    for (size_t y = 0; y < height; ++y) {
        float *o = out + y * width * NCOEF;
        for (size_t x = 0; x < width; ++x, o += NCOEF) {
            const uint32_t *s = samples + y * row_stride + x * pixel_stride;
            float acc[NCOEF] = { 0 };
            for (unsigned n = 0; n < n_lights; ++n) {
                float b = (float) s[(size_t) n * image_stride];
                for (int k = 0; k < NCOEF; ++k)
                    acc[k] = acc[k] + M[(size_t) k * n_lights + n] * b;
            }
            memcpy (o, acc, sizeof (acc));
        }
    }
}

/* ---------------------------------------------------------------- chroma
 * rti-builder/ptmlib.c:762-802 (ptm_cbcr_avg): unsigned 32-bit sums of all
 * three bytes over the lights (:774-787), truncating integer division by the
 * light count (:789-800).  Y is averaged too.
 */
void orc_chroma_avg (const uint8_t *stack, size_t pixels, unsigned n_lights, uint8_t *out) {
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < pixels; ++p) {
        unsigned int s0 = 0, s1 = 0, s2 = 0;
        for (unsigned n = 0; n < n_lights; ++n) {
            const uint8_t *t = stack + ((size_t) n * pixels + p) * 3;
            s0 += t[0];
            s1 += t[1];
            s2 += t[2];
        }
        out[p * 3 + 0] = (uint8_t) (s0 / n_lights);
        out[p * 3 + 1] = (uint8_t) (s1 / n_lights);
        out[p * 3 + 2] = (uint8_t) (s2 / n_lights);
    }
}

/* --------------------------------------------------- LRGB colour + normalise
 * rti-builder/ptm-encoder.c:327-353 (the loop in the encoder's main): the
 * averaged (Y, Cb, Cr) triplet is turned into RGB in place with the JFIF float
 * formulas (:337-339; Cb and Cr are centred by an int subtraction first,
 * :333-334), and the six coefficients are multiplied by 256/Y (:342-348).
 */
void orc_lrgb_colour_normalise (uint8_t *ycc_to_rgb, float *coeffs, size_t pixels) {
    for (size_t p = 0; p < pixels; ++p) {
        uint8_t *t = ycc_to_rgb + p * 3;
        float *c = coeffs + p * NCOEF;
        float y = t[0];
        float cb = t[1] - 128;
        float cr = t[2] - 128;
        t[0] = clip_u8 (y + 1.40200f * cr);
        t[1] = clip_u8 (y - 0.34414f * cb - 0.71414f * cr);
        t[2] = clip_u8 (y + 1.77200f * cb);
        float norm = 256.0f / y;
        for (int k = 0; k < NCOEF; ++k)
            c[k] *= norm;
    }
}

/* ---------------------------------------------------------- scale + bias
 * rti-builder/ptmlib.c:816-881 (ptm_scale_coefficients).
 *   phase A (:826-847) per-coefficient min and max with fminf/fmaxf over the
 *           FIRST coefficient plane only (the row pointer has no block index);
 *   phase B (:858-863) scale = (max-min)/256, bias = (int)(-256/(max-min)*min),
 *           inv_scale = 1/scale;
 *   phase C (:867-880) byte = CLIP(c * inv_scale + (float) bias) for every
 *           coefficient plane (`planes` = 1 for LRGB, 3 for RGB formats).
 * `minmax_out` (optional) receives min[6] then max[6].
 */
void orc_scale (const float *coeffs, size_t pixels, int planes,
                float *scale, int *bias, uint8_t *const *blocks, float *minmax_out) {
    float mn[NCOEF], mx[NCOEF];
    for (int k = 0; k < NCOEF; ++k) {
        mn[k] = FLT_MAX;
        mx[k] = -FLT_MAX;
    }
    for (size_t p = 0; p < pixels; ++p) {
        const float *c = coeffs + p * NCOEF;
        for (int k = 0; k < NCOEF; ++k) {
            mn[k] = fminf (mn[k], c[k]);
            mx[k] = fmaxf (mx[k], c[k]);
        }
    }
    if (minmax_out) {
        memcpy (minmax_out, mn, sizeof (mn));
        memcpy (minmax_out + NCOEF, mx, sizeof (mx));
    }
    float inv_scale[NCOEF];
    for (int k = 0; k < NCOEF; ++k) {
        scale[k] = (mx[k] - mn[k]) / 256.0f;
        bias[k] = -256.0f / (mx[k] - mn[k]) * mn[k];
        inv_scale[k] = 1.0f / scale[k];
    }
    for (int b = 0; b < planes; ++b) {
        const float *src = coeffs + (size_t) b * pixels * NCOEF;
        uint8_t *dst = blocks[b];
        #pragma omp parallel for schedule(static)
        for (size_t p = 0; p < pixels; ++p) {
            for (int k = 0; k < NCOEF; ++k)
                dst[p * NCOEF + k] = clip_u8 ((src[p * NCOEF + k] * inv_scale[k]) + bias[k]);
        }
    }
}

/* ---------------------------------------------------------------- relight
 * rti-builder/ptmlib.c:553-562 (light vector), :54-63 (UNSCALE / POLY: the
 * byte minus the bias as an int, converted to float, times the scale; products
 * with the light terms added left to right), :587-600 (LRGB: L = POLY / 255,
 * each colour byte times L, clipped) and :610-619 (RGB: POLY of each plane).
 * Output row y is stored row height-1-y (:587-589).  `out` is an interleaved
 * height x width x 3 raster, top row first -- what the reference hands to the
 * JPEG writer one scanline at a time.
 */
static inline float orc_poly (const uint8_t *c, const float *scale, const int *bias,
                              const float *light) {
    return scale[0] * (c[0] - bias[0]) * light[0]
         + scale[1] * (c[1] - bias[1]) * light[1]
         + scale[2] * (c[2] - bias[2]) * light[2]
         + scale[3] * (c[3] - bias[3]) * light[3]
         + scale[4] * (c[4] - bias[4]) * light[4]
         + scale[5] * (c[5] - bias[5]);
}

void orc_relight_lrgb (const uint8_t *coef, const uint8_t *rgb, size_t width, size_t height,
                       const float *scale, const int *bias, float u, float v, uint8_t *out) {
    const float light[NCOEF] = { u * u, v * v, u * v, u, v, 1.0f };
    #pragma omp parallel for schedule(static)
    for (size_t y = 0; y < height; ++y) {
        const size_t src = (height - 1 - y) * width;
        const uint8_t *c = coef + src * NCOEF;
        const uint8_t *t = rgb + src * 3;
        uint8_t *o = out + y * width * 3;
        for (size_t x = 0; x < width; ++x, c += NCOEF, t += 3, o += 3) {
            float L = orc_poly (c, scale, bias, light) / 255.0f;
            o[0] = clip_u8 (L * t[0]);
            o[1] = clip_u8 (L * t[1]);
            o[2] = clip_u8 (L * t[2]);
        }
    }
}

void orc_relight_rgb (const uint8_t *r, const uint8_t *g, const uint8_t *b,
                      size_t width, size_t height,
                      const float *scale, const int *bias, float u, float v, uint8_t *out) {
    const float light[NCOEF] = { u * u, v * v, u * v, u, v, 1.0f };
    #pragma omp parallel for schedule(static)
    for (size_t y = 0; y < height; ++y) {
        const size_t src = (height - 1 - y) * width * NCOEF;
        uint8_t *o = out + y * width * 3;
        for (size_t x = 0; x < width; ++x, o += 3) {
            o[0] = clip_u8 (orc_poly (r + src + x * NCOEF, scale, bias, light));
            o[1] = clip_u8 (orc_poly (g + src + x * NCOEF, scale, bias, light));
            o[2] = clip_u8 (orc_poly (b + src + x * NCOEF, scale, bias, light));
        }
    }
}

/* ------------------------------------------------------ whole LRGB encode
 * The call sequence of rti-builder/ptm-encoder.c:307-353,406 for an LRGB
 * format: fit the Y channel, average the triplets, colour + normalise, scale.
 * `coeffs` (pixels x 6 floats) is caller-provided scratch and holds the
 * normalised float coefficients on return.
 */
void orc_encode_lrgb (const uint8_t *stack, size_t width, size_t height, unsigned n_lights,
                      const float *M, float *coeffs, uint8_t *coef_block, uint8_t *rgb_block,
                      float *scale, int *bias, float *minmax_out) {
    const size_t pixels = width * height;
    orc_fit_u8 (stack, width, height, n_lights, 3, M, coeffs);
    orc_chroma_avg (stack, pixels, n_lights, rgb_block);
    orc_lrgb_colour_normalise (rgb_block, coeffs, pixels);
    uint8_t *blocks[1] = { coef_block };
    orc_scale (coeffs, pixels, 1, scale, bias, blocks, minmax_out);
}

/* The RGB-format sequence of rti-builder/ptm-encoder.c:272-284,406: one fit per
 * colour channel into three coefficient planes, one shared scale/bias. */
void orc_encode_rgb (const uint8_t *stack, size_t width, size_t height, unsigned n_lights,
                     const float *M, float *coeffs, uint8_t *const *blocks,
                     float *scale, int *bias, float *minmax_out) {
    const size_t pixels = width * height;
    for (int ch = 0; ch < 3; ++ch)
        orc_fit_u8 (stack + ch, width, height, n_lights, 3, M, coeffs + (size_t) ch * pixels * NCOEF);
    orc_scale (coeffs, pixels, 3, scale, bias, blocks, minmax_out);
}

/* EXTENSION -- 16-bit LRGB.  The reference has no such path (its ptm_cbcr_avg and colour loop
 * are 8 bit only), so this has NO reference to be pinned against; it states the semantics the
 * CUDA path implements: fit on the 16-bit luma exactly like ptm_fit_poly_uint
 * (rti-builder/ptmlib.c:727-760), truncated integer means at 16 bits (same formula as :774-800),
 * the colour conversion of rti-builder/ptm-encoder.c:331-339 on their top 8 bits, and
 * norm = 256 / (16-bit mean luma).  stack: uint16 [n][pixels][3]. */
void orc_lrgb16_finish (const uint16_t *stack, size_t pixels, unsigned n_lights, float *coeffs, uint8_t *rgb) {
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < pixels; ++p) {
        unsigned int s0 = 0, s1 = 0, s2 = 0;
        for (unsigned n = 0; n < n_lights; ++n) {
            const uint16_t *t = stack + ((size_t) n * pixels + p) * 3;
            s0 += t[0];
            s1 += t[1];
            s2 += t[2];
        }
        const unsigned ya = s0 / n_lights, cba = s1 / n_lights, cra = s2 / n_lights;
        float y = (float) (ya >> 8);
        float cb = (int) (cba >> 8) - 128;
        float cr = (int) (cra >> 8) - 128;
        rgb[p * 3 + 0] = clip_u8 (y + 1.40200f * cr);
        rgb[p * 3 + 1] = clip_u8 (y - 0.34414f * cb - 0.71414f * cr);
        rgb[p * 3 + 2] = clip_u8 (y + 1.77200f * cb);
        float norm = 256.0f / (float) ya;
        for (int k = 0; k < NCOEF; ++k)
            coeffs[p * NCOEF + k] *= norm;
    }
}

/* rti-builder/ptmlib.c:808-814 (ptm_normal), including its sign quirk in the
 * last line (1 - nu^2 + nv^2). */
void orc_normal (const float *c, float *nu, float *nv, float *nw) {
    float divisor = 4 * c[0] * c[1] - c[2] * c[2];
    *nu = (c[2] * c[4] - 2 * c[1] * c[3]) / divisor;
    *nv = (c[2] * c[3] - 2 * c[0] * c[4]) / divisor;
    *nw = sqrtf (1.0f - *nu * *nu + *nv * *nv);
}

/* ptm_normal over a whole plane: from float coefficients [pixels][6], or (coeffs == NULL) from a quantised
 * block unscaled with UNSCALE (rti-builder/ptmlib.c:54: scale * (byte - bias), int subtraction). */
void orc_normal_map (const float *coeffs, const uint8_t *bytes, const float *scale, const int *bias,
                     size_t pixels, float *normals) {
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < pixels; ++p) {
        float c[NCOEF];
        for (int k = 0; k < NCOEF; ++k)
            c[k] = coeffs ? coeffs[p * NCOEF + k] : scale[k] * (bytes[p * NCOEF + k] - bias[k]);
        orc_normal (c, normals + p * 3, normals + p * 3 + 1, normals + p * 3 + 2);
    }
}

/* rti-builder/ptmlib.c:601-609 -- the PTM_FORMAT_LUM branch of the relight loop.  Block 1 is addressed as
 * crcb_coefficients_t (rti-builder/ptmlib.h: JSAMPLE cr, cb -- 2 bytes) at pixel index yoffs + x, exactly
 * like the reference's pointer arithmetic; the raster holds (Y, Cb, Cr) triplets (JCS_YCbCr, :575). */
void orc_relight_lum (const uint8_t *coef, const uint8_t *crcb, size_t width, size_t height,
                      const float *scale, const int *bias, float u, float v, uint8_t *out) {
    const float light[NCOEF] = { u * u, v * v, u * v, u, v, 1.0f };
    #pragma omp parallel for schedule(static)
    for (size_t y = 0; y < height; ++y) {
        const size_t src = (height - 1 - y) * width;
        uint8_t *o = out + y * width * 3;
        for (size_t x = 0; x < width; ++x, o += 3) {
            o[0] = clip_u8 (orc_poly (coef + (src + x) * NCOEF, scale, bias, light));
            o[1] = clip_u8 (crcb[(src + x) * 2 + 1]);   /* cb */
            o[2] = clip_u8 (crcb[(src + x) * 2]);       /* cr */
        }
    }
}

/* RGB -> YCbCr per sample, in place: libjpeg's integer JFIF conversion (jccolor.c rgb_ycc_convert;
 * third party, absent from /root/reference: FIX(x) = (int) (x * 65536 + 0.5), ONE_HALF = 32768,
 * CBCR_OFFSET = 128 << 16).  The reference never runs it (its decoder delivers YCbCr,
 * rti-builder/ptm-encoder.c:188): this is the host statement of the optional RGB-input front end
 *  -- PARITY UNPINNED for this function: neither the reference nor this image
 * exposes libjpeg's converter on its own; tests/test_oracle.py checks it against the real-valued JFIF
 * matrix (within 1) and against known triplets. */
void orc_rgb_to_ycbcr (uint8_t *samples, size_t n_triplets) {
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_triplets; ++i) {
        uint8_t *t = samples + i * 3;
        const unsigned r = t[0], g = t[1], b = t[2];
        t[0] = (uint8_t) ((19595u * r + 38470u * g + 7471u * b + 32768u) >> 16);
        t[1] = (uint8_t) ((8388608u + 32767u + 32768u * b - 11059u * r - 21709u * g) >> 16);
        t[2] = (uint8_t) ((8388608u + 32767u + 32768u * r - 27439u * g - 5329u * b) >> 16);
    }
}

static int cmp_u8 (const void *a, const void *b) {
    return (int) *(const uint8_t *) a - (int) *(const uint8_t *) b;
}

/* rti-builder/ptm-encoder.c:356-400 -- the disabled LRGB branch (color_components == 666).
 *   :366-372  L = R + G + B for every sample
 *   :376-380  ptm_fit_poly_uint on L (pixel stride 1): unnormalised coefficients
 *   :384-386  ptm_cbcr_avg over the RGB triplets (truncated mean) -- "FIXME should use median here!" (:383);
 *             median != 0 computes that median instead ((lower + upper middle) / 2; EXTENSION, the
 *             reference has no arithmetic for it)
 *   :391-399  colour = 256 * (int) mean / *l, stored into a JSAMPLE, with l walking L[0 .. pixels): the sums
 *             of the FIRST light.  *l == 0 is a division by zero there; 0 is stored here.
 * stack: uint8 [n][pixels][3] RGB. */
void orc_fit_lsum (const uint8_t *stack, size_t pixels, unsigned n_lights, const float *M, int median,
                   float *coeffs, uint8_t *colour) {
    unsigned int *L = malloc (sizeof (unsigned int) * (size_t) n_lights * pixels);
    for (size_t i = 0; i < (size_t) n_lights * pixels; ++i)
        L[i] = (unsigned int) stack[i * 3] + stack[i * 3 + 1] + stack[i * 3 + 2];
    orc_fit_u32 (L, pixels, 1, n_lights, 1, M, coeffs);
    orc_chroma_avg (stack, pixels, n_lights, colour);
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < pixels; ++p) {
        for (int ch = 0; ch < 3; ++ch) {
            unsigned int c = colour[p * 3 + ch];
            if (median) {
                uint8_t v[2048];
                for (unsigned n = 0; n < n_lights; ++n)
                    v[n] = stack[((size_t) n * pixels + p) * 3 + ch];
                qsort (v, n_lights, 1, cmp_u8);
                c = ((unsigned int) v[(n_lights - 1) / 2] + v[n_lights / 2]) / 2;
            }
            colour[p * 3 + ch] = L[p] ? (uint8_t) (256 * (int) c / L[p]) : 0;
        }
    }
    free (L);
}

const char *orc_blas_corename (void) {
    return openblas_get_corename ();
}

void orc_blas_set_threads (int n) {
    openblas_set_num_threads (n);
}

/* OpenMP team size for every later parallel region of this process (this library's and the reference's
 * libptmref.so share one libgomp).  An OMP_NUM_THREADS written into the environment after libgomp has been
 * loaded is never read; launchers such as torchrun export OMP_NUM_THREADS=1 before the process starts. */
#include <omp.h>
void orc_omp_set_threads (int n) {
    omp_set_dynamic (0);
    omp_set_num_threads (n > 0 ? n : 1);
}

int orc_omp_max_threads (void) {
    return omp_get_max_threads ();
}

/* threads a parallel region really gets (the number a benchmark should report) */
int orc_omp_team_size (void) {
    int n = 1;
    #pragma omp parallel
    {
        #pragma omp master
        n = omp_get_num_threads ();
    }
    return n;
}

/*
 * ptmlib_host.c -- the ptmlib.h interface (include/ptmlib.h) on top of the B200 CUDA library.
 *
 * Drop-in for the reference's rti-builder/ptmlib.c: same exported functions, same ownership
 * and error conventions (NULL returns, messages on stderr, no error codes), same PTM_1.2 file
 * layout.  What differs is where the arithmetic runs: ptm_fit_poly_*, ptm_cbcr_avg,
 * ptm_scale_coefficients and the pixel loop of ptm_write_jpeg are GPU kernels reached through
 * include/ptm_b200.h.  ptm_svd stays on the host in LAPACKE, as in the reference.  There is no
 * CPU fallback: without a GPU these calls print the CUDA error and exit(1).
 *
 * "ref :N" = line N of rti-builder/ptmlib.c in the reference tree.
 */
#define _POSIX_C_SOURCE 200809L

#include <assert.h>
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ptm_b200.h"
#include "../../include/ptmlib.h"
#include "ptm_codec.h"
#include "ptm_lapack.h"

/* external definition for the header's inline CLIP (the reference has none and only links
 * when optimising) */
extern inline JSAMPLE CLIP (const float f);

/* ref :44-51 -- id, blocks, ptm_blocks, color_components, jpeg_streams, name */
const ptm_format_t ptm_formats[6] = {
    { PTM_FORMAT_RGB,       3, 3, 0,  0, "PTM_FORMAT_RGB"       },
    { PTM_FORMAT_JPEG_RGB,  3, 3, 0, 18, "PTM_FORMAT_JPEG_RGB"  },
    { PTM_FORMAT_LRGB,      2, 1, 3,  0, "PTM_FORMAT_LRGB"      },
    { PTM_FORMAT_JPEG_LRGB, 2, 1, 3,  9, "PTM_FORMAT_JPEG_LRGB" },
    { PTM_FORMAT_LUM,       2, 1, 2,  0, "PTM_FORMAT_LUM"       },
    { 0,                    0, 0, 0,  0, NULL },
};

/* ref :66-71 -- the colour matrix a LUM header carries (gray only) */
static const float LUM_COLOR_MATRIX[16] = {
    0, 1, 0, 0,
    0, 1, 0, 0,
    0, 1, 0, 0,
    0, 0, 0, 1
};

/* ------------------------------------------------------------------ GPU context */

static ptmb_ctx_t *g_ctx = NULL;

static void gpu_die (const char *what) {
    fprintf (stderr, "ptmlib (B200): %s failed: %s\n", what, ptmb_last_error ());
    exit (1);
}

static ptmb_ctx_t *gpu (void) {
    if (!g_ctx) {
        /* PTM_B200_DEVICE selects the GPU (default 0); the ordinal is passed on unchanged and the
         * process environment is left alone (a library must not narrow CUDA_VISIBLE_DEVICES under a
         * host that also runs torch / NCCL -- the command-line mains do that for themselves, before
         * their first CUDA call, see cli/cli_common.h). */
        const char *dev = getenv ("PTM_B200_DEVICE");
        const int ordinal = dev ? atoi (dev) : 0;
        if (ptmb_create (ordinal, NULL, &g_ctx))
            gpu_die ("ptmb_create");
        /* the drop-in reproduces the reference decoder's rasters byte for byte unless told otherwise
         * (PTM_RELIGHT=fast: every byte within +-1, sweeps at the memory rate) */
        if (!getenv ("PTM_RELIGHT"))
            ptmb_set_relight_mode (g_ctx, PTMB_RELIGHT_EXACT);
    }
    return g_ctx;
}

/* PTM_B200_DEVICES=0,1,...: ptm_encode_stack splits the image into one row slab per listed GPU (the reference
 * spreads the same rows over host cores, ref :704).  NULL when the variable is not set or names one GPU. */
static ptmb_multi_t *g_multi = NULL;

static ptmb_multi_t *gpu_multi (void) {
    static int tried = 0;
    if (tried)
        return g_multi;
    tried = 1;
    const char *list = getenv ("PTM_B200_DEVICES");
    if (!list || !*list)
        return NULL;
    int devices[16], n = 0;
    for (const char *p = list; *p && n < 16;) {
        char *end = NULL;
        const long d = strtol (p, &end, 10);
        if (end == p)
            break;
        devices[n++] = (int) d;
        p = *end == ',' ? end + 1 : end;
    }
    if (n < 2)
        return NULL;
    if (ptmb_multi_create (devices, n, &g_multi))
        gpu_die ("ptmb_multi_create (PTM_B200_DEVICES)");
    return g_multi;
}

/* for ptm_codec.c and the command lines (not part of ptmlib.h) */
ptmb_ctx_t *ptm_internal_gpu (void) {
    return gpu ();
}

static int sample_bytes (const ptm_header_t *h, int n_block) {
    /* ref :286-290 -- colour blocks are always sized for 3 components */
    return n_block < h->format->ptm_blocks ? PTM_COEFFICIENTS : RGB_COEFFICIENTS;
}

/* ------------------------------------------------------------------ formats, memory */

const ptm_format_t *ptm_get_format (const char *format_name) {
    for (const ptm_format_t *f = ptm_formats; f->name; ++f) {
        if (strcmp (format_name, f->name) == 0)
            return f;
    }
    return NULL;
}

ptm_header_t *ptm_alloc_header () {
    return calloc (1, sizeof (ptm_header_t));
}

ptm_block_t *ptm_alloc_blocks (const ptm_header_t *h) {
    const size_t pixels = h->dimen[0] * h->dimen[1];
    ptm_block_t *blocks = calloc (h->format->blocks, sizeof (ptm_block_t));
    for (int i = 0; i < h->format->blocks; ++i)
        blocks[i] = malloc ((size_t) sample_bytes (h, i) * pixels);
    return blocks;
}

void ptm_free_blocks (const ptm_header_t *h, ptm_block_t *blocks) {
    for (int i = 0; i < h->format->blocks; ++i)
        free (blocks[i]);
    free (blocks);
}

/* ------------------------------------------------------------------ PTM_1.2 header */

static void put_ints (FILE *fp, int n, const int *v) {
    for (int i = 0; i < n; ++i)
        fprintf (fp, i ? " %d" : "%d", v[i]);
    fputc ('\n', fp);
}

static void put_sizes (FILE *fp, int n, const size_t *v) {
    for (int i = 0; i < n; ++i)
        fprintf (fp, i ? " %lu" : "%lu", (unsigned long) v[i]);
    fputc ('\n', fp);
}

static void put_floats (FILE *fp, int n, const float *v) {
    for (int i = 0; i < n; ++i)
        fprintf (fp, i ? " %f" : "%f", v[i]);
    fputc ('\n', fp);
}

/* Numbers are read WITHOUT the reference's trailing blank in the format ("%d ", ref :120-136): that
 * blank swallows every white-space byte after the last header field, i.e. also leading payload bytes
 * 0x09-0x0D / 0x20 of an uncompressed block or stream -- ordinary coefficient values -- and shifts the
 * whole payload.  Each conversion skips the white space in front of it, and header_end() consumes the
 * one line end that terminates the header (deliberate deviation from the reference). */
static void get_ints (FILE *fp, int n, int *v) {
    for (int i = 0; i < n; ++i)
        if (fscanf (fp, "%d", v + i) != 1)
            v[i] = 0;
}

static void get_sizes (FILE *fp, int n, size_t *v) {
    for (int i = 0; i < n; ++i) {
        unsigned long t = 0;
        if (fscanf (fp, "%lu", &t) != 1)
            t = 0;
        v[i] = t;
    }
}

static void get_floats (FILE *fp, int n, float *v) {
    for (int i = 0; i < n; ++i)
        if (fscanf (fp, "%f", v + i) != 1)
            v[i] = 0.0f;
}

/* blanks up to and including exactly one '\n' (the writers end every header line with one) */
static void header_end (FILE *fp) {
    int ch;
    while ((ch = fgetc (fp)) == ' ' || ch == '\t' || ch == '\r')
        ;
    if (ch != '\n' && ch != EOF)
        ungetc (ch, fp);
}

static char *read_trimmed_line (FILE *fp) {
    char *line = NULL;
    size_t cap = 0;
    ssize_t got = getline (&line, &cap, fp);
    if (got < 0) {
        free (line);
        return NULL;
    }
    while (got > 0 && isspace ((unsigned char) line[got - 1]))
        line[--got] = '\0';
    return line;
}

/* ref :176-216 */
ptm_header_t *ptm_read_header (FILE *fp) {
    char *line = read_trimmed_line (fp);
    if (!line || strcmp (line, "PTM_1.2") != 0) {
        fprintf (stderr, "not a PTM file\n");
        free (line);
        return NULL;
    }
    free (line);
    line = read_trimmed_line (fp);
    const ptm_format_t *format = line ? ptm_get_format (line) : NULL;
    if (!format) {
        fprintf (stderr, "unsupported PTM format: %s\n", line ? line : "");
        free (line);
        return NULL;
    }
    free (line);

    ptm_header_t *h = ptm_alloc_header ();
    h->format = format;
    get_sizes (fp, 2, h->dimen);
    get_floats (fp, PTM_COEFFICIENTS, h->scale);
    get_ints (fp, PTM_COEFFICIENTS, h->bias);
    if (format->id == PTM_FORMAT_LUM) {
        /* ptm_write_header emits the 4 x 4 colour matrix for LUM (ref :225-227) but the reference's reader
         * never consumes it (its LUM round trip is broken); read it here so that
         * the payload starts where the writer put it */
        float matrix[16];
        get_floats (fp, 16, matrix);
    }
    if (format->jpeg_streams) {
        const int n = format->jpeg_streams;
        get_ints (fp, 1, h->compression_param);
        get_ints (fp, n, h->transforms);
        get_ints (fp, n, h->motion_vector_x);
        get_ints (fp, n, h->motion_vector_y);
        get_ints (fp, n, h->order);
        get_ints (fp, n, h->reference_planes);
        get_sizes (fp, n, h->compressed_size);
        get_sizes (fp, n, h->side_info_sizes);
    } else {
        h->compression_param[0] = 90;   /* ref :211 -- quality used when relit images are written */
    }
    header_end (fp);
    return h;
}

/* ref :218-239 */
void ptm_write_header (FILE *fp, const ptm_header_t *h) {
    fprintf (fp, "PTM_1.2\n%s\n", h->format->name);
    fprintf (fp, "%lu\n%lu\n", (unsigned long) h->dimen[0], (unsigned long) h->dimen[1]);
    put_floats (fp, PTM_COEFFICIENTS, h->scale);
    put_ints (fp, PTM_COEFFICIENTS, h->bias);
    if (h->format->id == PTM_FORMAT_LUM)
        put_floats (fp, 16, LUM_COLOR_MATRIX);
    if (h->format->jpeg_streams) {
        const int n = h->format->jpeg_streams;
        put_ints (fp, 1, h->compression_param);
        put_ints (fp, n, h->transforms);
        put_ints (fp, n, h->motion_vector_x);
        put_ints (fp, n, h->motion_vector_y);
        put_ints (fp, n, h->order);
        put_ints (fp, n, h->reference_planes);
        put_sizes (fp, n, h->compressed_size);
        put_sizes (fp, n, h->side_info_sizes);
    }
}

/* ------------------------------------------------------------------ block payload */

/* ref :319-324 */
static void read_plain_blocks (FILE *fp, const ptm_header_t *h, ptm_block_t *blocks) {
    const size_t pixels = h->dimen[0] * h->dimen[1];
    if (h->format->id == PTM_FORMAT_LUM) {
        /* the writer's layout (ref :333-347): per pixel six coefficients, then Cr, Cb.  The reference reads
         * this as two planar blocks (ref :319-324), which cannot work; here the records are taken apart:
         * block 0 = coefficients, block 1 = 2-byte (cr, cb) records -- the layout the relight loop
         * (ref :601-609) addresses block 1 with */
        JSAMPLE *c = blocks[0], *crcb = blocks[1];
        for (size_t p = 0; p < pixels; ++p, c += PTM_COEFFICIENTS, crcb += CBCR_COEFFICIENTS) {
            JSAMPLE rec[PTM_COEFFICIENTS + CBCR_COEFFICIENTS];
            if (fread (rec, sizeof (rec), 1, fp) != 1) {
                fprintf (stderr, "short read in PTM block 0\n");
                break;
            }
            memcpy (c, rec, PTM_COEFFICIENTS);
            crcb[0] = rec[PTM_COEFFICIENTS];
            crcb[1] = rec[PTM_COEFFICIENTS + 1];
        }
        return;
    }
    for (int i = 0; i < h->format->blocks; ++i) {
        if (fread (blocks[i], sample_bytes (h, i), pixels, fp) != pixels)
            fprintf (stderr, "short read in PTM block %d\n", i);
    }
}

/* ref :333-356 -- LUM interleaves the six coefficients with (Cr, Cb) per pixel */
static void write_plain_blocks (FILE *fp, const ptm_header_t *h, ptm_block_t *blocks) {
    const size_t pixels = h->dimen[0] * h->dimen[1];
    if (h->format->id == PTM_FORMAT_LUM) {
        const JSAMPLE *c = blocks[0];
        const JSAMPLE *ycc = blocks[1];
        for (size_t p = 0; p < pixels; ++p, c += PTM_COEFFICIENTS, ycc += 3) {
            JSAMPLE rec[PTM_COEFFICIENTS + CBCR_COEFFICIENTS];
            memcpy (rec, c, PTM_COEFFICIENTS);
            rec[PTM_COEFFICIENTS] = ycc[2];       /* Cr first */
            rec[PTM_COEFFICIENTS + 1] = ycc[1];
            fwrite (rec, sizeof (rec), 1, fp);
        }
        return;
    }
    for (int i = 0; i < h->format->blocks; ++i)
        fwrite (blocks[i], sample_bytes (h, i), pixels, fp);
}

/* ref :249-260 -- a plane predicted from another: add (src or 255 - src) - 128, byte wrap */
static void add_reference_plane (JSAMPLE *dst, const JSAMPLE *src, size_t n, int invert) {
    for (size_t i = 0; i < n; ++i)
        dst[i] = (JSAMPLE) (dst[i] + (invert ? 255 - src[i] : src[i]) - 128);
}

/* ref :262-284 -- 5-byte patches: big-endian offset, then the sample */
static void apply_patches (JSAMPLE *plane, size_t plane_size, const unsigned char *info, size_t size) {
    for (size_t i = 0; i + 5 <= size; i += 5) {
        const size_t off = ((size_t) info[i] << 24) | ((size_t) info[i + 1] << 16) |
                           ((size_t) info[i + 2] << 8) | (size_t) info[i + 3];
        if (off < plane_size)
            plane[off] = info[i + 4];
    }
}

/* ref :367-458 -- one grayscale stream per coefficient plane, optional prediction + patches,
 * then back to the interleaved block layout */
static void read_stream_blocks (FILE *fp, const ptm_header_t *h, ptm_block_t *blocks) {
    const int n = h->format->jpeg_streams;
    const size_t width = h->dimen[0], height = h->dimen[1], pixels = width * height;
    JSAMPLE *planes[MAX_JPEG_STREAMS] = { 0 };
    unsigned char *side[MAX_JPEG_STREAMS] = { 0 };

    for (int i = 0; i < n; ++i) {
        unsigned char *packed = malloc (h->compressed_size[i] ? h->compressed_size[i] : 1);
        if (fread (packed, 1, h->compressed_size[i], fp) != h->compressed_size[i])
            fprintf (stderr, "short read in PTM stream %d\n", i);
        unsigned int pw = 0, ph = 0, pc = 0;
        if (ptm_codec_decode_mem (packed, h->compressed_size[i], &pw, &ph, &pc, &planes[i])) {
            fprintf (stderr, "cannot decode PTM stream %d\n", i);
            exit (1);
        }
        assert (pw == width);
        assert (ph == height);
        assert (pc == 1);
        free (packed);
        if (h->side_info_sizes[i] > 0) {
            side[i] = malloc (h->side_info_sizes[i]);
            if (fread (side[i], 1, h->side_info_sizes[i], fp) != h->side_info_sizes[i])
                fprintf (stderr, "short read in PTM side info %d\n", i);
        }
    }
    for (int i = 0; i < n; ++i) {
        int comp = -1;
        for (int j = 0; j < n; ++j) {
            if (h->order[j] == i) {
                comp = j;
                break;
            }
        }
        assert (0 <= comp && comp < n);
        const int ref = h->reference_planes[comp];
        assert (-1 <= ref && ref < n);
        if (ref > -1)
            add_reference_plane (planes[comp], planes[ref], pixels, (h->transforms[comp] & 1) > 0);
        if (side[comp])
            apply_patches (planes[comp], pixels, side[comp], h->side_info_sizes[comp]);
    }
    for (int i = 0; i < n; ++i) {
        const int b = i / PTM_COEFFICIENTS, k = i % PTM_COEFFICIENTS;
        const int step = sample_bytes (h, b);
        JSAMPLE *dst = blocks[b] + k;
        for (size_t p = 0; p < pixels; ++p)
            dst[p * step] = planes[i][p];
        free (planes[i]);
        free (side[i]);
    }
}

/* ref :468-522,532-546 */
static void write_stream_blocks (FILE *fp, ptm_header_t *h, ptm_block_t *blocks) {
    const int n = h->format->jpeg_streams;
    const size_t width = h->dimen[0], height = h->dimen[1], pixels = width * height;
    unsigned char *streams[MAX_JPEG_STREAMS] = { 0 };
    h->compression_param[0] = 90;

    #pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < n; ++i) {
        const int b = i / PTM_COEFFICIENTS, k = i % PTM_COEFFICIENTS;
        const int step = sample_bytes (h, b);
        const JSAMPLE *src = blocks[b] + k;
        JSAMPLE *plane = malloc (pixels ? pixels : 1);
        for (size_t p = 0; p < pixels; ++p)
            plane[p] = src[p * step];
        size_t size = 0;
        if (ptm_codec_encode_mem (plane, (unsigned int) width, (unsigned int) height, h->compression_param[0],
                                  &streams[i], &size)) {
            fprintf (stderr, "cannot encode PTM stream %d\n", i);
            exit (1);
        }
        h->compressed_size[i] = size;
        free (plane);
    }
    for (int i = 0; i < n; ++i) {
        h->order[i] = i;
        h->reference_planes[i] = -1;
    }
    ptm_write_header (fp, h);
    for (int i = 0; i < n; ++i) {
        fwrite (streams[i], h->compressed_size[i], 1, fp);
        free (streams[i]);
    }
}

/* ref :524-530 */
void ptm_read_ptm (FILE *fp, const ptm_header_t *h, ptm_block_t *blocks) {
    if (h->format->jpeg_streams > 0)
        read_stream_blocks (fp, h, blocks);
    else
        read_plain_blocks (fp, h, blocks);
}

/* ref :532-550 */
void ptm_write_ptm (FILE *fp, ptm_header_t *h, ptm_block_t *blocks) {
    if (h->format->jpeg_streams > 0) {
        write_stream_blocks (fp, h, blocks);
    } else {
        ptm_write_header (fp, h);
        write_plain_blocks (fp, h, blocks);
    }
}

/* ------------------------------------------------------------------ pseudo-inverse */

/* ref :630-689 -- A[n] = {u^2 v^2 uv u v 1}, thin SVD, M = V diag(1/s) U' (6 x n, row major) */
float *ptm_svd_uv (const float *u, const float *v, int n) {
    const int k = PTM_COEFFICIENTS;
    float *A = calloc ((size_t) n * k, sizeof (float));
    float *U = calloc ((size_t) n * k, sizeof (float));
    float S[PTM_COEFFICIENTS * PTM_COEFFICIENTS] = { 0 };
    float VT[PTM_COEFFICIENTS * PTM_COEFFICIENTS] = { 0 };
    float M1[PTM_COEFFICIENTS * PTM_COEFFICIENTS] = { 0 };
    float sv[PTM_COEFFICIENTS] = { 0 };
    float *M = NULL;

    for (int i = 0; i < n; ++i) {
        float *a = A + (size_t) i * k;
        a[0] = u[i] * u[i];
        a[1] = v[i] * v[i];
        a[2] = u[i] * v[i];
        a[3] = u[i];
        a[4] = v[i];
        a[5] = 1.0f;
    }
    const int info = LAPACKE_sgesdd (PTM_LAPACK_ROW_MAJOR, 'S', n, k, A, k, sv, U, k, VT, k);
    if (info <= 0) {   /* the reference only rejects info > 0 (:662) */
        for (int i = 0; i < k; ++i)
            S[i * k + i] = 1.0f / sv[i];
        M = calloc ((size_t) k * n, sizeof (float));
        cblas_sgemm (PTM_CBLAS_ROW_MAJOR, PTM_CBLAS_TRANS, PTM_CBLAS_NO_TRANS, k, k, k,
                     1.0f, VT, k, S, k, 0.0f, M1, k);
        cblas_sgemm (PTM_CBLAS_ROW_MAJOR, PTM_CBLAS_NO_TRANS, PTM_CBLAS_TRANS, k, n, k,
                     1.0f, M1, k, U, k, 0.0f, M, n);
    }
    free (A);
    free (U);
    return M;
}

float *ptm_svd (decoder_t **decoders, int n_decoders) {
    float *u = malloc (sizeof (float) * (n_decoders > 0 ? n_decoders : 1));
    float *v = malloc (sizeof (float) * (n_decoders > 0 ? n_decoders : 1));
    for (int i = 0; i < n_decoders; ++i) {
        u[i] = decoders[i]->u;
        v[i] = decoders[i]->v;
    }
    float *M = ptm_svd_uv (u, v, n_decoders);
    free (u);
    free (v);
    return M;
}

/* ------------------------------------------------------------------ GPU-backed arithmetic */

/* ref :692-725 */
void ptm_fit_poly_jsample (const ptm_image_info_t *info, const JSAMPLE *buffer, size_t pixel_stride,
                           const float *M, ptm_unscaled_coefficients_t *output) {
    if (ptmb_fit_channel_host (gpu (), buffer, PTMB_U8, pixel_stride, info->width, info->height,
                               info->n_decoders, M, (float *) output))
        gpu_die ("ptm_fit_poly_jsample");
}

/* ref :727-760 */
void ptm_fit_poly_uint (const ptm_image_info_t *info, const unsigned int *buffer, size_t pixel_stride,
                        const float *M, ptm_unscaled_coefficients_t *output) {
    if (ptmb_fit_channel_host (gpu (), buffer, PTMB_U32, pixel_stride, info->width, info->height,
                               info->n_decoders, M, (float *) output))
        gpu_die ("ptm_fit_poly_uint");
}

/* ref :762-802 */
void ptm_cbcr_avg (const ptm_image_info_t *info, const ycbcr_coefficients_t *buffer,
                   ycbcr_coefficients_t *block) {
    if (ptmb_chroma_avg_host (gpu (), (const uint8_t *) buffer, info->pixels, info->n_decoders,
                              (uint8_t *) block))
        gpu_die ("ptm_cbcr_avg");
}

/* rti-builder/ptm-encoder.c:327-353 */
void ptm_lrgb_finish (const ptm_image_info_t *info, ycbcr_coefficients_t *ycbcr_to_rgb,
                      ptm_unscaled_coefficients_t *coeffs) {
    if (ptmb_lrgb_colour_normalise_host (gpu (), (uint8_t *) ycbcr_to_rgb, (float *) coeffs, info->pixels))
        gpu_die ("ptm_lrgb_finish");
}

/* ref :816-881 */
void ptm_scale_coefficients (ptm_header_t *h, const ptm_unscaled_coefficients_t *unscaled,
                             ptm_block_t *scaled) {
    ptmb_scale_bias_t sb;
    if (ptmb_scale_host (gpu (), (const float *) unscaled, h->dimen[0] * h->dimen[1], h->format->ptm_blocks,
                         scaled, &sb))
        gpu_die ("ptm_scale_coefficients");
    for (int i = 0; i < PTM_COEFFICIENTS; ++i) {
        h->scale[i] = sb.scale[i];
        h->bias[i] = sb.bias[i];
    }
}

void ptm_encode_stack (ptm_header_t *h, const ptm_image_info_t *info, const JSAMPLE *buffer,
                       const float *M, ptm_block_t *blocks) {
    ptmb_scale_bias_t sb;
    /* LUM (rti-builder/ptm-encoder.c:286-305): fit Y, average the triplets, no colour conversion and no
     * normalisation -- the same fused pass with PTMB_FIT_LUM */
    const int lum = h->format->color_components == 2;
    ptmb_multi_t *multi = gpu_multi ();
    if (multi ? ptmb_multi_set_fit_mode (multi, lum ? PTMB_FIT_LUM : 0) : ptmb_set_fit_mode (gpu (), lum ? PTMB_FIT_LUM : 0))
        gpu_die ("ptmb_set_fit_mode");
    if (multi) {
        if (ptmb_multi_encode_host (multi, buffer, PTMB_U8, info->width, info->height, info->n_decoders, M,
                                    h->format->ptm_blocks, blocks, NULL, &sb))
            gpu_die ("ptm_encode_stack (multi-GPU)");
    } else if (ptmb_encode_host (gpu (), buffer, PTMB_U8, info->width, info->height, info->n_decoders, M,
                                 h->format->ptm_blocks, blocks, NULL, &sb))
        gpu_die ("ptm_encode_stack");
    for (int i = 0; i < PTM_COEFFICIENTS; ++i) {
        h->scale[i] = sb.scale[i];
        h->bias[i] = sb.bias[i];
    }
}

/* Encode straight from JPEG files: every image is decoded by nvJPEG INTO the device-resident stack
 * (bottom row first, like rti-builder/ptm-encoder.c:252-256), so only the compressed bytes cross
 * PCIe; then the fit / quantise kernels run on the stack and the blocks come back.  `jpegs[i]` /
 * `lens[i]` are the n_decoders compressed files in memory.  Returns 0 on success, non-zero if the
 * inputs are not usable this way (the caller then takes the host-buffer path). */
int ptm_encode_jpeg_files (ptm_header_t *h, const ptm_image_info_t *info, const unsigned char *const *jpegs,
                           const size_t *lens, const float *M, ptm_block_t *blocks) {
    ptmb_ctx_t *ctx = gpu ();
    const size_t pixels = info->pixels, n = info->n_decoders;
    const int planes = h->format->ptm_blocks;
    if (h->format->color_components == 2)
        return 1;   /* LUM goes through the step-by-step path */
    void *stack = NULL, *coeffs = NULL, *bytes = NULL, *sb = NULL, *keys = NULL;
    /* image stride rounded up to 16 bytes keeps every light on the bulk-copy alignment */
    const size_t stride = (pixels * 3 + 15) & ~(size_t) 15;
    int rc = 1;
    if (ptmb_malloc (ctx, stride * n, &stack) || ptmb_malloc (ctx, pixels * 24 * planes, &coeffs) ||
        ptmb_malloc (ctx, pixels * 6 * planes + pixels * 3, &bytes) || ptmb_malloc (ctx, sizeof (ptmb_scale_bias_t), &sb) ||
        ptmb_malloc (ctx, PTMB_KEYS * sizeof (int32_t), &keys))
        goto done;
    if (ptmb_set_matrix (ctx, M, (unsigned) n) || ptmb_set_fit_mode (ctx, 0))
        goto done;
    for (size_t i = 0; i < n; ++i) {
        unsigned w = 0, hgt = 0, c = 0;
        if (ptmb_jpeg_info (ctx, jpegs[i], lens[i], &w, &hgt, &c) || w != info->width || hgt != info->height || c != 3)
            goto done;
        if (ptmb_jpeg_decode (ctx, jpegs[i], lens[i], h->format->color_components == 0, 1,
                              (unsigned char *) stack + i * stride, 1))
            goto done;
    }
    unsigned char *rgb_dev = (unsigned char *) bytes + pixels * 6 * planes;
    if (planes == 1) {
        if (ptmb_encode_lrgb_dev (ctx, NULL, stack, PTMB_U8, stride, pixels, coeffs, rgb_dev, bytes, sb))
            goto done;
    } else {
        if (ptmb_keys_init (ctx, keys) || ptmb_fit_rgb (ctx, stack, PTMB_U8, stride, pixels, coeffs, pixels * 6, keys))
            goto done;
        for (int b = 0; b < planes; ++b)
            if (ptmb_quantise (ctx, (float *) coeffs + (size_t) b * pixels * 6, pixels, keys,
                               (unsigned char *) bytes + (size_t) b * pixels * 6, b == 0 ? sb : NULL))
                goto done;
    }
    for (int b = 0; b < planes; ++b)
        if (ptmb_download (ctx, blocks[b], (unsigned char *) bytes + (size_t) b * pixels * 6, pixels * 6))
            goto done;
    if (planes == 1 && ptmb_download (ctx, blocks[1], rgb_dev, pixels * 3))
        goto done;
    ptmb_scale_bias_t sbh;
    if (ptmb_download (ctx, &sbh, sb, sizeof (sbh)))
        goto done;
    for (int i = 0; i < PTM_COEFFICIENTS; ++i) {
        h->scale[i] = sbh.scale[i];
        h->bias[i] = sbh.bias[i];
    }
    rc = 0;
done:
    if (rc)
        fprintf (stderr, "ptmlib (B200): device JPEG path not used: %s\n", ptmb_last_error ());
    ptmb_free (ctx, stack);
    ptmb_free (ctx, coeffs);
    ptmb_free (ctx, bytes);
    ptmb_free (ctx, sb);
    ptmb_free (ctx, keys);
    return rc;
}

/* the pixel loop of ref :587-621, for a batch of light directions */
void ptm_relight_batch (const ptm_header_t *h, ptm_block_t *blocks, const float *uv, unsigned int n_dirs,
                        JSAMPLE *out) {
    int32_t bias[PTM_COEFFICIENTS];
    for (int i = 0; i < PTM_COEFFICIENTS; ++i)
        bias[i] = h->bias[i];
    /* PTM_FORMAT_LUM (ref :601-609): a YCbCr raster, block 1 addressed as (cr, cb) records */
    const int planes = h->format->color_components == 2 ? 2 : h->format->ptm_blocks;
    if (ptmb_relight_host (gpu (), (const uint8_t *const *) blocks, planes, h->dimen[0],
                           h->dimen[1], h->scale, bias, uv, n_dirs, out))
        gpu_die ("ptm_relight");
}

void ptm_relight (const ptm_header_t *h, ptm_block_t *blocks, float u, float v, JSAMPLE *out) {
    const float uv[2] = { u, v };
    ptm_relight_batch (h, blocks, uv, 1, out);
}

/* ref :553-628 -- relight, then hand the raster to the image writer */
void ptm_write_jpeg (FILE *fp, const ptm_header_t *h, ptm_block_t *blocks, float u, float v) {
    const size_t pixels = h->dimen[0] * h->dimen[1];
    JSAMPLE *raster = malloc (pixels > 0 ? pixels * 3 : 1);
    ptm_relight (h, blocks, u, v, raster);
    /* quality from the header; RGB colour space, YCbCr for LUM (ref :575-577) */
    if (ptm_codec_write (fp, raster, (unsigned int) h->dimen[0], (unsigned int) h->dimen[1], 3,
                         h->format->color_components == 2, h->compression_param[0]))
        fprintf (stderr, "cannot write image\n");
    free (raster);
}

/* ------------------------------------------------------------------ small helpers */

/* ref :808-814, including the sign of the last term as the reference has it */
void ptm_normal (const ptm_unscaled_coefficients_t *c, float *nu, float *nv, float *nw) {
    const float divisor = 4 * c->cu2 * c->cv2 - c->cuv * c->cuv;
    *nu = (c->cuv * c->cv - 2 * c->cv2 * c->cu) / divisor;
    *nv = (c->cuv * c->cu - 2 * c->cu2 * c->cv) / divisor;
    *nw = sqrtf (1.0f - *nu * *nu + *nv * *nv);
}

void ptm_normal_map (const ptm_header_t *h, ptm_block_t *blocks, float *normals) {
    int32_t bias[PTM_COEFFICIENTS];
    for (int i = 0; i < PTM_COEFFICIENTS; ++i)
        bias[i] = h->bias[i];
    if (ptmb_normal_map_host (gpu (), NULL, blocks[0], h->scale, bias, h->dimen[0] * h->dimen[1], normals))
        gpu_die ("ptm_normal_map");
}

/* ref :883-891 */
void ptm_print_matrix (const char *name, const float *M, int m, int n) {
    fprintf (stderr, "\n%s = [", name);
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j)
            fprintf (stderr, " %3.6f", M[i * n + j]);
        fprintf (stderr, ";\n");
    }
    fprintf (stderr, "]\n");
}

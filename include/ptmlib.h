/*
 * ptmlib.h -- the PTM library interface of cceh/rti's rti-builder, served by the B200 build.
 *
 * Source compatible with the reference header (rti-builder/ptmlib.h): same type names,
 * struct layouts, constants and the same 15 functions + ptm_formats[], so that
 * ptm-encoder.c / ptm-decoder.c / ptm-exploder.c style callers compile and link against
 * libptm.so (rti_b200/csrc/ptmlib_host.c) instead of ptmlib.c.  The per-pixel arithmetic
 * behind these calls runs on the GPU through include/ptm_b200.h; the pseudo-inverse
 * (ptm_svd) stays on the host in LAPACKE as in the reference.
 *
 * Layout notes are given as "ref :line" = rti-builder/ptmlib.h line of the reference.
 *
 * JPEG: when built with -DPTM_HAVE_LIBJPEG the real <jpeglib.h> is included, exactly like the
 * reference.  Without it (this image has no libjpeg headers) JSAMPLE is declared here and
 * image I/O uses the lossless raw container of ptm_codec.h; only decoder_t's tail differs.
 */
#ifndef PTMLIB_H
#define PTMLIB_H

#include <stddef.h>
#include <stdio.h>

#ifdef PTM_HAVE_LIBJPEG
#include <jpeglib.h>
#else
#ifndef PTM_JSAMPLE_DEFINED
#define PTM_JSAMPLE_DEFINED
typedef unsigned char JSAMPLE;
#endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ref :16-34 */
#define PTM_COEFFICIENTS      6      /* polynomial terms per pixel: u^2 v^2 uv u v 1 */
#define RGB_COEFFICIENTS      3      /* extra colour bytes per pixel in LRGB files */
#define CBCR_COEFFICIENTS     2      /* extra chroma bytes per pixel in LUM files */
#define MAX_PTM_BLOCKS        3      /* coefficient blocks: 1 (LRGB, LUM) or 3 (RGB) */
#define MAX_COLOR_COMPONENTS  RGB_COEFFICIENTS
#define MAX_JPEG_STREAMS      RGB_COEFFICIENTS * PTM_COEFFICIENTS   /* unparenthesised, as in ref :34 */

/* ref :37-45 -- quantised coefficients of one pixel, u^2 first */
typedef struct {
    JSAMPLE cu2, cv2, cuv, cu, cv, c1;
} ptm_coefficients_t;

/* ref :48-55 -- float coefficients of one pixel, same order */
typedef struct {
    float cu2, cv2, cuv, cu, cv, c1;
} ptm_unscaled_coefficients_t;

/* ref :58-62, :65-69, :72-75 -- colour triplets / pairs as stored */
typedef struct {
    JSAMPLE r, g, b;
} rgb_coefficients_t;

typedef struct {
    JSAMPLE y, cb, cr;
} ycbcr_coefficients_t;

typedef struct {
    JSAMPLE cr, cb;    /* note the order: Cr first in LUM files */
} crcb_coefficients_t;

/* ref :78-84 */
typedef enum {
    PTM_FORMAT_RGB = 1,
    PTM_FORMAT_LUM,
    PTM_FORMAT_LRGB,
    PTM_FORMAT_JPEG_RGB,
    PTM_FORMAT_JPEG_LRGB
} ptm_formats_enum_t;

/* ref :87-99 */
typedef struct {
    ptm_formats_enum_t id;
    int blocks;              /* blocks to allocate: coefficient blocks + 0/1 colour block */
    int ptm_blocks;          /* coefficient blocks: 1 = LRGB / LUM, 3 = RGB */
    int color_components;    /* channels not PTM-encoded: 0 = RGB, 2 = LUM, 3 = LRGB */
    int jpeg_streams;        /* JPEG streams in the file, 0 if uncompressed */
    const char *name;        /* e.g. "PTM_FORMAT_LRGB" */
} ptm_format_t;

/* ref :102 -- five formats and a NULL-named terminator */
extern const ptm_format_t ptm_formats[6];

/* ref :105-115 */
typedef struct {
    size_t width;
    size_t height;
    size_t pixels;           /* width * height */
    size_t row_stride;       /* samples from one row to the next in the stack */
    size_t decoder_stride;   /* samples from one image to the next in the stack */
    unsigned int n_decoders; /* number of images = number of lights */
} ptm_image_info_t;

/* ref :123-138 -- everything that goes into the PTM_1.2 text header */
typedef struct {
    const ptm_format_t *format;
    size_t dimen           [2];                  /* width, height */
    float scale            [PTM_COEFFICIENTS];
    int bias               [PTM_COEFFICIENTS];
    /* compressed formats only */
    int compression_param  [1];                  /* JPEG quality */
    int transforms         [MAX_JPEG_STREAMS];
    int motion_vector_x    [MAX_JPEG_STREAMS];
    int motion_vector_y    [MAX_JPEG_STREAMS];
    int order              [MAX_JPEG_STREAMS];
    int reference_planes   [MAX_JPEG_STREAMS];
    size_t compressed_size [MAX_JPEG_STREAMS];
    size_t side_info_sizes [MAX_JPEG_STREAMS];
} ptm_header_t;

/* ref :142 */
typedef JSAMPLE *ptm_block_t;

/* ref :145-151 -- one input image.  ptm_svd reads only u and v. */
typedef struct {
    FILE *fp;
    char *filename;
    float u, v, w;
#ifdef PTM_HAVE_LIBJPEG
    struct jpeg_decompress_struct dinfo;
#else
    struct {
        unsigned int output_width, output_height;
        int output_components;
        unsigned int output_scanline;
    } dinfo;                 /* raw-container decoder state (ptm_codec.h) */
#endif
} decoder_t;

/* ref :156-158 -- clip to 0..255, then truncate.  The reference relies on an inline-only
 * definition; libptm.so also provides the external one. */
inline JSAMPLE CLIP (const float f) {
    return f > 255.0f ? 255 : (f < 0.0f ? 0 : f);
}

/* ref :160 */
#define LUM(r,g,b) ((unsigned int) r + (unsigned int) g + (unsigned int) b)

/* ref :170 -- format by name, NULL if unknown */
const ptm_format_t *ptm_get_format (const char *format_name);

/* ref :177 -- zeroed header, release with free() */
ptm_header_t *ptm_alloc_header ();

/* ref :187, :195 -- format->blocks pointers, 6 or 3 bytes per pixel each */
ptm_block_t *ptm_alloc_blocks (const ptm_header_t *ptm_header);
void ptm_free_blocks (const ptm_header_t *ptm_header, ptm_block_t *blocks);

/* ref :204, :212 -- PTM_1.2 text header */
ptm_header_t *ptm_read_header (FILE *fp);
void ptm_write_header (FILE *fp, const ptm_header_t *ptm_header);

/* ref :223, :234 -- block payload, with per-plane (de)compression for the JPEG formats */
void ptm_read_ptm  (FILE *fp, const ptm_header_t *ptm_header, ptm_block_t *blocks);
void ptm_write_ptm (FILE *fp, ptm_header_t *ptm_header, ptm_block_t *blocks);

/* ref :245 -- relight for light (u, v) and write the image to fp */
void ptm_write_jpeg (FILE *fp, const ptm_header_t *ptm_header, ptm_block_t *blocks, float u, float v);

/* ref :255 -- 6 x n_decoders pseudo-inverse of the light matrix, calloc'd; NULL on failure */
float *ptm_svd (decoder_t **decoders, int n_decoders);

/* ref :271-281 -- per-pixel fit of one channel; buffer = [image][y][x][pixel_stride] */
void ptm_fit_poly_jsample (const ptm_image_info_t *info,
                           const JSAMPLE *buffer,
                           size_t pixel_stride,
                           const float *M,
                           ptm_unscaled_coefficients_t *output);

void ptm_fit_poly_uint (const ptm_image_info_t *info,
                        const unsigned int *buffer,
                        size_t pixel_stride,
                        const float *M,
                        ptm_unscaled_coefficients_t *output);

/* ref :295-297 -- truncated integer mean of every byte over the images */
void ptm_cbcr_avg (const ptm_image_info_t *info,
                   const ycbcr_coefficients_t *buffer,
                   ycbcr_coefficients_t *block);

/* ref :310-312 -- scale / bias into the header, bytes into the blocks */
void ptm_scale_coefficients (ptm_header_t *ptm_header,
                             const ptm_unscaled_coefficients_t *unscaled,
                             ptm_block_t *scaled);

/* ref :322, :333 */
void ptm_normal (const ptm_unscaled_coefficients_t *coeffs, float *nu, float *nv, float *nw);
void ptm_print_matrix (const char *name, const float* M, int m, int n);

/* ---- additions of the B200 build (not in the reference header) ---------------------- */

/* ptm_svd from plain arrays (what ptm_svd does after reading decoder->u / ->v). */
float *ptm_svd_uv (const float *u, const float *v, int n_lights);

/* The LRGB colour + normalise loop that the reference keeps in the encoder's main
 * (rti-builder/ptm-encoder.c:327-353), as a library call. */
void ptm_lrgb_finish (const ptm_image_info_t *info, ycbcr_coefficients_t *ycbcr_to_rgb,
                      ptm_unscaled_coefficients_t *coeffs);

/* Whole encode of a decoded stack in one GPU pass (fit + chroma + colour + scale):
 * replaces the call sequence rti-builder/ptm-encoder.c:270-353,406.  Fills header->scale /
 * bias and the blocks. */
void ptm_encode_stack (ptm_header_t *ptm_header, const ptm_image_info_t *info, const JSAMPLE *buffer,
                       const float *M, ptm_block_t *blocks);

/* The same from JPEG files held in memory: nvJPEG decodes every image straight into the stack on the
 * GPU (only compressed bytes cross PCIe).  Returns 0 on success; non-zero means "use ptm_encode_stack". */
int ptm_encode_jpeg_files (ptm_header_t *ptm_header, const ptm_image_info_t *info,
                           const unsigned char *const *jpegs, const size_t *lens, const float *M,
                           ptm_block_t *blocks);

/* Relight without the codec: RGB raster [height][width][3] into `out`; the batch form takes
 * n_dirs (u, v) pairs and fills n_dirs rasters in one GPU pass (what ptm-exploder needs). */
void ptm_relight (const ptm_header_t *ptm_header, ptm_block_t *blocks, float u, float v, JSAMPLE *out);
void ptm_relight_batch (const ptm_header_t *ptm_header, ptm_block_t *blocks, const float *uv,
                        unsigned int n_dirs, JSAMPLE *out);

/* ptm_normal for every pixel of a PTM (first coefficient block, unscaled with the header's scale / bias):
 * normals float [height * width][3] = (nu, nv, nw) per stored pixel. */
void ptm_normal_map (const ptm_header_t *ptm_header, ptm_block_t *blocks, float *normals);

#ifdef __cplusplus
}
#endif

#endif

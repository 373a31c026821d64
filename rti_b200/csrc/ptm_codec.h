/* Image codec used by the host library and the command-line mains.
 *
 * The reference reads and writes JPEG through libjpeg (rti-builder/ptm-encoder.c:181-190,
 * 257-261; rti-builder/ptmlib.c:371-412,473-520,565-627).  Here two containers are understood,
 * told apart by their first bytes:
 *
 *   JPEG  (FF D8)  decoded / encoded on the GPU with nvJPEG (ptmb_jpeg_* in include/ptm_b200.h),
 *                  colour handling restated from libjpeg (YCbCr passed through with libjpeg's chroma
 *                  upsampling, or its integer YCbCr -> RGB);
 *   RAWJ           a lossless raw container for pre-decoded stacks: "RAWJ", u32 width, u32 height,
 *                  u32 components (little endian), then the scanlines, top row first, interleaved.
 *                  A RAWJ file holds whatever triplets its writer put there -- what libjpeg hands the
 *                  reference with out_color_space set (rti-builder/ptm-encoder.c:188).  Parity tests
 *                  use it, because it takes the codec out of the comparison.
 *
 * Output format: JPEG unless the environment says PTM_IMAGE_CODEC=raw.
 */
#ifndef PTM_CODEC_H
#define PTM_CODEC_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTM_CODEC_RAW  0
#define PTM_CODEC_JPEG 1

typedef struct {
    unsigned int width, height, components;
    int kind;                 /* PTM_CODEC_RAW or PTM_CODEC_JPEG */
    unsigned char *data;      /* JPEG: the whole compressed file, owned until ptm_codec_release */
    size_t len;
} ptm_image_dims_t;

/* Sniff the stream at fp and read its header (JPEG: the whole file).  Returns 0 on success. */
int ptm_codec_read_dims (FILE *fp, ptm_image_dims_t *dims);

/* Deliver `rows` scanlines through row pointers (so the caller can flip, like
 * rti-builder/ptm-encoder.c:252-259).  want_rgb selects libjpeg's JCS_RGB or JCS_YCbCr behaviour for
 * JPEG input and is ignored for RAWJ.  Returns 0 on success. */
int ptm_codec_read_rows (FILE *fp, ptm_image_dims_t *dims, unsigned char **row_pointers, unsigned int rows,
                         int want_rgb);

void ptm_codec_release (ptm_image_dims_t *dims);

/* Write a whole image, top row first.  components 1 = gray, 3 = RGB (is_ycbcr = 0) or YCbCr (1). */
int ptm_codec_write (FILE *fp, const unsigned char *samples, unsigned int width, unsigned int height,
                     unsigned int components, int is_ycbcr, int quality);

/* Memory forms for the per-plane grayscale streams of the compressed PTM formats. */
int ptm_codec_encode_mem (const unsigned char *samples, unsigned int width, unsigned int height, int quality,
                          unsigned char **out, size_t *out_size);
int ptm_codec_decode_mem (const unsigned char *buf, size_t size, unsigned int *width, unsigned int *height,
                          unsigned int *components, unsigned char **samples);

/* 1 when output should be the raw container (PTM_IMAGE_CODEC=raw). */
int ptm_codec_output_is_raw (void);

#ifdef __cplusplus
}
#endif

#endif

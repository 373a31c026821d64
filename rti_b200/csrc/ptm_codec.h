/* Image container used by the host library and the command-line mains.
 *
 * The reference reads and writes JPEG through libjpeg (rti-builder/ptm-encoder.c:181-190,
 * 257-261; rti-builder/ptmlib.c:371-412,473-520,565-627).  Entropy coding is outside the PTM
 * hot path and this image has no libjpeg headers, so the B200 build carries a lossless raw
 * container for pre-decoded stacks:
 *
 *     "RAWJ"  u32 width  u32 height  u32 components  (little endian)  then the scanlines,
 *     top row first, components interleaved.
 *
 * A RAWJ file holds whatever triplets its writer put there -- YCbCr for the LRGB path, RGB for
 * the RGB path -- i.e. what libjpeg hands the reference with out_color_space set
 * (rti-builder/ptm-encoder.c:188).
 */
#ifndef PTM_CODEC_H
#define PTM_CODEC_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    unsigned int width, height, components;
} ptm_image_dims_t;

/* Read the stream header from fp.  Returns 0 on success. */
int ptm_codec_read_dims (FILE *fp, ptm_image_dims_t *dims);

/* Read `rows` scanlines into row pointers (so the caller can flip, like
 * rti-builder/ptm-encoder.c:252-259).  Returns 0 on success. */
int ptm_codec_read_rows (FILE *fp, const ptm_image_dims_t *dims, unsigned char **row_pointers, unsigned int rows);

/* Write a whole image, top row first.  Returns 0 on success. */
int ptm_codec_write (FILE *fp, const unsigned char *samples, const ptm_image_dims_t *dims);

/* Memory forms for the per-plane streams of the compressed PTM formats. */
int ptm_codec_encode_mem (const unsigned char *samples, const ptm_image_dims_t *dims,
                          unsigned char **out, size_t *out_size);
int ptm_codec_decode_mem (const unsigned char *buf, size_t size, ptm_image_dims_t *dims, unsigned char **samples);

#ifdef __cplusplus
}
#endif

#endif

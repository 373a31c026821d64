/* Stand-in "jpeglib.h" for building the reference PTM sources as a test oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The image ships no libjpeg development files, and
 * entropy coding is outside the PTM hot path anyway.  This header declares the
 * small slice of the libjpeg API that the reference's ptmlib.c and its three
 * command-line mains touch, and rawjpeg.c implements it as a LOSSLESS raw
 * container ("RAWJ": magic, width, height, components, then the scanlines).
 * With it the reference's C files compile unmodified and run end to end on
 * pre-decoded stacks, so every arithmetic step of the reference (including the
 * LRGB loop that lives in the encoder's main and the relight loop that is
 * fused with the JPEG writer) is exercised for parity -- with the codec taken
 * out of the comparison.
 *
 * It is NOT ABI compatible with any real libjpeg. */
#ifndef ORACLE_SHIM_JPEGLIB_H
#define ORACLE_SHIM_JPEGLIB_H

#include <stddef.h>
#include <stdio.h>

typedef unsigned char JSAMPLE;
typedef JSAMPLE *JSAMPROW;
typedef JSAMPROW *JSAMPARRAY;
typedef unsigned char JOCTET;
typedef unsigned int JDIMENSION;
typedef int boolean;

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define CENTERJSAMPLE 128
#define MAXJSAMPLE 255

#define JPOOL_PERMANENT 0
#define JPOOL_IMAGE 1

typedef enum {
    JCS_UNKNOWN,
    JCS_GRAYSCALE,
    JCS_RGB,
    JCS_YCbCr,
    JCS_CMYK,
    JCS_YCCK
} J_COLOR_SPACE;

struct jpeg_error_mgr {
    int unused;
};

struct jpeg_common_struct;
typedef struct jpeg_common_struct *j_common_ptr;

struct jpeg_memory_mgr {
    JSAMPARRAY (*alloc_sarray) (j_common_ptr cinfo, int pool_id,
                                JDIMENSION samplesperrow, JDIMENSION numrows);
    void *(*alloc_large) (j_common_ptr cinfo, int pool_id, size_t sizeofobject);
};

/* Fields shared by both directions; must stay first in both structs below. */
#define RAWJ_COMMON_FIELDS        \
    struct jpeg_error_mgr *err;   \
    struct jpeg_memory_mgr *mem;  \
    void *rawj_priv

struct jpeg_common_struct {
    RAWJ_COMMON_FIELDS;
};

struct jpeg_decompress_struct {
    RAWJ_COMMON_FIELDS;
    JDIMENSION image_width;
    JDIMENSION image_height;
    int num_components;
    J_COLOR_SPACE jpeg_color_space;
    J_COLOR_SPACE out_color_space;
    JDIMENSION output_width;
    JDIMENSION output_height;
    int output_components;
    JDIMENSION output_scanline;
};

struct jpeg_compress_struct {
    RAWJ_COMMON_FIELDS;
    JDIMENSION image_width;
    JDIMENSION image_height;
    int input_components;
    J_COLOR_SPACE in_color_space;
    JDIMENSION next_scanline;
};

typedef struct jpeg_decompress_struct *j_decompress_ptr;
typedef struct jpeg_compress_struct *j_compress_ptr;

struct jpeg_error_mgr *jpeg_std_error (struct jpeg_error_mgr *err);

void jpeg_create_decompress (j_decompress_ptr dinfo);
void jpeg_stdio_src (j_decompress_ptr dinfo, FILE *fp);
void jpeg_mem_src (j_decompress_ptr dinfo, const unsigned char *buf, unsigned long size);
int jpeg_read_header (j_decompress_ptr dinfo, boolean require_image);
void jpeg_calc_output_dimensions (j_decompress_ptr dinfo);
boolean jpeg_start_decompress (j_decompress_ptr dinfo);
JDIMENSION jpeg_read_scanlines (j_decompress_ptr dinfo, JSAMPARRAY rows, JDIMENSION max_lines);
boolean jpeg_finish_decompress (j_decompress_ptr dinfo);
void jpeg_destroy_decompress (j_decompress_ptr dinfo);

void jpeg_create_compress (j_compress_ptr cinfo);
void jpeg_stdio_dest (j_compress_ptr cinfo, FILE *fp);
void jpeg_mem_dest (j_compress_ptr cinfo, unsigned char **outbuffer, unsigned long *outsize);
void jpeg_set_defaults (j_compress_ptr cinfo);
void jpeg_set_quality (j_compress_ptr cinfo, int quality, boolean force_baseline);
void jpeg_start_compress (j_compress_ptr cinfo, boolean write_all_tables);
JDIMENSION jpeg_write_scanlines (j_compress_ptr cinfo, JSAMPARRAY rows, JDIMENSION num_lines);
void jpeg_finish_compress (j_compress_ptr cinfo);
void jpeg_destroy_compress (j_compress_ptr cinfo);

#endif

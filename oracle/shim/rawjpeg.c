/* Lossless raw-container implementation of the libjpeg API slice declared in
 * the stand-in jpeglib.h next to this file.  TEST INFRASTRUCTURE ONLY.
 *
 * "RAWJ" stream layout (little endian):
 *     bytes 0..3   'R' 'A' 'W' 'J'
 *     u32 width, u32 height, u32 components
 *     height * width * components sample bytes, top scanline first
 *
 * The colour space fields are carried but never acted on: a RAWJ file holds
 * whatever triplets its writer put there (YCbCr for the LRGB encoder path,
 * RGB for the RGB path), exactly like a libjpeg "null conversion".
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "jpeglib.h"

typedef struct rawj_alloc {
    struct rawj_alloc *next;
} rawj_alloc_t;

typedef struct {
    struct jpeg_memory_mgr mem;
    rawj_alloc_t *allocs;
    /* source */
    FILE *src_fp;
    const unsigned char *src_buf;
    size_t src_size;
    size_t src_pos;
    /* destination */
    FILE *dst_fp;
    unsigned char **dst_buf;
    unsigned long *dst_size;
    unsigned char *dst_data;
    size_t dst_cap;
    size_t dst_len;
} rawj_priv_t;

static void rawj_die (const char *msg) {
    fprintf (stderr, "rawjpeg: %s\n", msg);
    exit (2);
}

static void *rawj_pool_alloc (rawj_priv_t *p, size_t size) {
    rawj_alloc_t *a = malloc (sizeof (rawj_alloc_t) + 16 + size);
    if (!a)
        rawj_die ("out of memory");
    a->next = p->allocs;
    p->allocs = a;
    return (unsigned char *) a + 16;   /* keep 16-byte alignment of the payload */
}

static void *rawj_alloc_large (j_common_ptr cinfo, int pool_id, size_t size) {
    (void) pool_id;
    return rawj_pool_alloc ((rawj_priv_t *) cinfo->rawj_priv, size);
}

static JSAMPARRAY rawj_alloc_sarray (j_common_ptr cinfo, int pool_id,
                                     JDIMENSION samplesperrow, JDIMENSION numrows) {
    (void) pool_id;
    rawj_priv_t *p = (rawj_priv_t *) cinfo->rawj_priv;
    JSAMPARRAY rows = rawj_pool_alloc (p, (size_t) numrows * sizeof (JSAMPROW));
    JSAMPLE *data = rawj_pool_alloc (p, (size_t) numrows * samplesperrow);
    for (JDIMENSION y = 0; y < numrows; ++y)
        rows[y] = data + (size_t) y * samplesperrow;
    return rows;
}

static rawj_priv_t *rawj_new_priv (void) {
    rawj_priv_t *p = calloc (1, sizeof (*p));
    if (!p)
        rawj_die ("out of memory");
    p->mem.alloc_sarray = rawj_alloc_sarray;
    p->mem.alloc_large = rawj_alloc_large;
    return p;
}

static void rawj_free_priv (rawj_priv_t *p) {
    while (p->allocs) {
        rawj_alloc_t *next = p->allocs->next;
        free (p->allocs);
        p->allocs = next;
    }
    free (p->dst_data);
    free (p);
}

struct jpeg_error_mgr *jpeg_std_error (struct jpeg_error_mgr *err) {
    err->unused = 0;
    return err;
}

/* ---------------------------------------------------------------- decoder */

void jpeg_create_decompress (j_decompress_ptr dinfo) {
    struct jpeg_error_mgr *err = dinfo->err;
    memset (dinfo, 0, sizeof (*dinfo));
    dinfo->err = err;
    rawj_priv_t *p = rawj_new_priv ();
    dinfo->rawj_priv = p;
    dinfo->mem = &p->mem;
}

void jpeg_stdio_src (j_decompress_ptr dinfo, FILE *fp) {
    rawj_priv_t *p = dinfo->rawj_priv;
    p->src_fp = fp;
    p->src_buf = NULL;
}

void jpeg_mem_src (j_decompress_ptr dinfo, const unsigned char *buf, unsigned long size) {
    rawj_priv_t *p = dinfo->rawj_priv;
    p->src_fp = NULL;
    p->src_buf = buf;
    p->src_size = size;
    p->src_pos = 0;
}

static void rawj_read (rawj_priv_t *p, void *dst, size_t n) {
    if (p->src_fp) {
        if (fread (dst, 1, n, p->src_fp) != n)
            rawj_die ("short read");
    } else {
        if (!p->src_buf || p->src_pos + n > p->src_size)
            rawj_die ("short read from memory source");
        memcpy (dst, p->src_buf + p->src_pos, n);
        p->src_pos += n;
    }
}

int jpeg_read_header (j_decompress_ptr dinfo, boolean require_image) {
    (void) require_image;
    rawj_priv_t *p = dinfo->rawj_priv;
    unsigned char magic[4];
    uint32_t dims[3];
    rawj_read (p, magic, 4);
    if (memcmp (magic, "RAWJ", 4))
        rawj_die ("not a RAWJ stream");
    rawj_read (p, dims, sizeof (dims));
    dinfo->image_width = dims[0];
    dinfo->image_height = dims[1];
    dinfo->num_components = (int) dims[2];
    dinfo->jpeg_color_space = dims[2] == 1 ? JCS_GRAYSCALE : JCS_YCbCr;
    dinfo->out_color_space = dims[2] == 1 ? JCS_GRAYSCALE : JCS_RGB;
    dinfo->output_scanline = 0;
    return 1;
}

void jpeg_calc_output_dimensions (j_decompress_ptr dinfo) {
    dinfo->output_width = dinfo->image_width;
    dinfo->output_height = dinfo->image_height;
    dinfo->output_components = dinfo->num_components;
}

boolean jpeg_start_decompress (j_decompress_ptr dinfo) {
    jpeg_calc_output_dimensions (dinfo);
    dinfo->output_scanline = 0;
    return TRUE;
}

JDIMENSION jpeg_read_scanlines (j_decompress_ptr dinfo, JSAMPARRAY rows, JDIMENSION max_lines) {
    rawj_priv_t *p = dinfo->rawj_priv;
    size_t row_bytes = (size_t) dinfo->output_width * dinfo->output_components;
    JDIMENSION done = 0;
    /* like libjpeg, hand back only a few lines per call so callers must loop */
    if (max_lines > 8)
        max_lines = 8;
    while (done < max_lines && dinfo->output_scanline < dinfo->output_height) {
        rawj_read (p, rows[done], row_bytes);
        ++done;
        ++dinfo->output_scanline;
    }
    return done;
}

boolean jpeg_finish_decompress (j_decompress_ptr dinfo) {
    (void) dinfo;
    return TRUE;
}

void jpeg_destroy_decompress (j_decompress_ptr dinfo) {
    if (dinfo->rawj_priv)
        rawj_free_priv (dinfo->rawj_priv);
    dinfo->rawj_priv = NULL;
    dinfo->mem = NULL;
}

/* ---------------------------------------------------------------- encoder */

void jpeg_create_compress (j_compress_ptr cinfo) {
    struct jpeg_error_mgr *err = cinfo->err;
    memset (cinfo, 0, sizeof (*cinfo));
    cinfo->err = err;
    rawj_priv_t *p = rawj_new_priv ();
    cinfo->rawj_priv = p;
    cinfo->mem = &p->mem;
}

void jpeg_stdio_dest (j_compress_ptr cinfo, FILE *fp) {
    rawj_priv_t *p = cinfo->rawj_priv;
    p->dst_fp = fp;
    p->dst_buf = NULL;
}

void jpeg_mem_dest (j_compress_ptr cinfo, unsigned char **outbuffer, unsigned long *outsize) {
    rawj_priv_t *p = cinfo->rawj_priv;
    p->dst_fp = NULL;
    p->dst_buf = outbuffer;
    p->dst_size = outsize;
}

void jpeg_set_defaults (j_compress_ptr cinfo) {
    (void) cinfo;
}

void jpeg_set_quality (j_compress_ptr cinfo, int quality, boolean force_baseline) {
    (void) cinfo;
    (void) quality;
    (void) force_baseline;
}

static void rawj_write (rawj_priv_t *p, const void *src, size_t n) {
    if (p->dst_fp) {
        if (fwrite (src, 1, n, p->dst_fp) != n)
            rawj_die ("short write");
        return;
    }
    if (p->dst_len + n > p->dst_cap) {
        size_t cap = p->dst_cap ? p->dst_cap : 4096;
        while (cap < p->dst_len + n)
            cap *= 2;
        p->dst_data = realloc (p->dst_data, cap);
        if (!p->dst_data)
            rawj_die ("out of memory");
        p->dst_cap = cap;
    }
    memcpy (p->dst_data + p->dst_len, src, n);
    p->dst_len += n;
}

void jpeg_start_compress (j_compress_ptr cinfo, boolean write_all_tables) {
    (void) write_all_tables;
    rawj_priv_t *p = cinfo->rawj_priv;
    uint32_t dims[3] = { cinfo->image_width, cinfo->image_height,
                         (uint32_t) cinfo->input_components };
    p->dst_len = 0;
    cinfo->next_scanline = 0;
    rawj_write (p, "RAWJ", 4);
    rawj_write (p, dims, sizeof (dims));
}

JDIMENSION jpeg_write_scanlines (j_compress_ptr cinfo, JSAMPARRAY rows, JDIMENSION num_lines) {
    rawj_priv_t *p = cinfo->rawj_priv;
    size_t row_bytes = (size_t) cinfo->image_width * cinfo->input_components;
    JDIMENSION done = 0;
    while (done < num_lines && cinfo->next_scanline < cinfo->image_height) {
        rawj_write (p, rows[done], row_bytes);
        ++done;
        ++cinfo->next_scanline;
    }
    return done;
}

void jpeg_finish_compress (j_compress_ptr cinfo) {
    rawj_priv_t *p = cinfo->rawj_priv;
    if (p->dst_fp) {
        fflush (p->dst_fp);
    } else if (p->dst_buf) {
        *p->dst_buf = p->dst_data;
        *p->dst_size = p->dst_len;
    }
}

void jpeg_destroy_compress (j_compress_ptr cinfo) {
    if (cinfo->rawj_priv)
        rawj_free_priv (cinfo->rawj_priv);
    cinfo->rawj_priv = NULL;
    cinfo->mem = NULL;
}

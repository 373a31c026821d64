/* Image codec front end (see ptm_codec.h): raw container on the host, JPEG through nvJPEG. */
#define _POSIX_C_SOURCE 200809L

#include "ptm_codec.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ptm_b200.h"

/* ptmlib_host.c: the process-wide GPU context */
ptmb_ctx_t *ptm_internal_gpu (void);

static const unsigned char RAW_MAGIC[4] = { 'R', 'A', 'W', 'J' };

int ptm_codec_output_is_raw (void) {
    const char *e = getenv ("PTM_IMAGE_CODEC");
    return e && strcmp (e, "raw") == 0;
}

static int slurp (FILE *fp, const unsigned char *head, size_t head_len, unsigned char **out, size_t *len) {
    size_t cap = 1 << 20, n = head_len;
    unsigned char *buf = malloc (cap);
    if (!buf)
        return 1;
    memcpy (buf, head, head_len);
    for (;;) {
        if (n == cap) {
            cap *= 2;
            unsigned char *nb = realloc (buf, cap);
            if (!nb) {
                free (buf);
                return 1;
            }
            buf = nb;
        }
        size_t got = fread (buf + n, 1, cap - n, fp);
        n += got;
        if (got == 0)
            break;
    }
    *out = buf;
    *len = n;
    return 0;
}

int ptm_codec_read_dims (FILE *fp, ptm_image_dims_t *dims) {
    unsigned char magic[4];
    memset (dims, 0, sizeof (*dims));
    if (fread (magic, 1, 4, fp) != 4)
        return 1;
    if (memcmp (magic, RAW_MAGIC, 4) == 0) {
        uint32_t d[3];
        if (fread (d, sizeof (uint32_t), 3, fp) != 3)
            return 1;
        dims->kind = PTM_CODEC_RAW;
        dims->width = d[0];
        dims->height = d[1];
        dims->components = d[2];
        return (d[0] == 0 || d[1] == 0 || d[2] == 0 || d[2] > 4) ? 1 : 0;
    }
    if (magic[0] == 0xFF && magic[1] == 0xD8) {
        if (slurp (fp, magic, 4, &dims->data, &dims->len))
            return 1;
        dims->kind = PTM_CODEC_JPEG;
        if (ptmb_jpeg_info (ptm_internal_gpu (), dims->data, dims->len, &dims->width, &dims->height,
                            &dims->components)) {
            fprintf (stderr, "JPEG header: %s\n", ptmb_last_error ());
            ptm_codec_release (dims);
            return 1;
        }
        return 0;
    }
    return 1;
}

void ptm_codec_release (ptm_image_dims_t *dims) {
    free (dims->data);
    dims->data = NULL;
    dims->len = 0;
}

int ptm_codec_read_rows (FILE *fp, ptm_image_dims_t *dims, unsigned char **row_pointers, unsigned int rows,
                         int want_rgb) {
    const size_t row_bytes = (size_t) dims->width * dims->components;
    if (dims->kind == PTM_CODEC_RAW) {
        for (unsigned int y = 0; y < rows; ++y) {
            if (fread (row_pointers[y], 1, row_bytes, fp) != row_bytes)
                return 1;
        }
        return 0;
    }
    unsigned char *img = malloc (row_bytes * dims->height);
    if (!img)
        return 1;
    if (ptmb_jpeg_decode (ptm_internal_gpu (), dims->data, dims->len, want_rgb, 0, img, 0)) {
        fprintf (stderr, "JPEG decode: %s\n", ptmb_last_error ());
        free (img);
        return 1;
    }
    for (unsigned int y = 0; y < rows && y < dims->height; ++y)
        memcpy (row_pointers[y], img + (size_t) y * row_bytes, row_bytes);
    free (img);
    return 0;
}

static void raw_header (unsigned char *h, unsigned int w, unsigned int hgt, unsigned int c) {
    const uint32_t d[3] = { w, hgt, c };
    memcpy (h, RAW_MAGIC, 4);
    memcpy (h + 4, d, 12);
}

int ptm_codec_encode_mem (const unsigned char *samples, unsigned int width, unsigned int height, int quality,
                          unsigned char **out, size_t *out_size) {
    const size_t bytes = (size_t) width * height;
    if (ptm_codec_output_is_raw ()) {
        unsigned char *buf = malloc (16 + bytes);
        if (!buf)
            return 1;
        raw_header (buf, width, height, 1);
        memcpy (buf + 16, samples, bytes);
        *out = buf;
        *out_size = 16 + bytes;
        return 0;
    }
    if (ptmb_jpeg_encode (ptm_internal_gpu (), samples, width, height, 1, 0, quality, out, out_size)) {
        fprintf (stderr, "JPEG encode: %s\n", ptmb_last_error ());
        return 1;
    }
    return 0;
}

int ptm_codec_decode_mem (const unsigned char *buf, size_t size, unsigned int *width, unsigned int *height,
                          unsigned int *components, unsigned char **samples) {
    if (size >= 16 && memcmp (buf, RAW_MAGIC, 4) == 0) {
        uint32_t d[3];
        memcpy (d, buf + 4, 12);
        const size_t bytes = (size_t) d[0] * d[1] * d[2];
        if (size < 16 + bytes)
            return 1;
        *width = d[0];
        *height = d[1];
        *components = d[2];
        *samples = malloc (bytes ? bytes : 1);
        if (!*samples)
            return 1;
        memcpy (*samples, buf + 16, bytes);
        return 0;
    }
    if (size >= 4 && buf[0] == 0xFF && buf[1] == 0xD8) {
        ptmb_ctx_t *ctx = ptm_internal_gpu ();
        if (ptmb_jpeg_info (ctx, buf, size, width, height, components))
            return 1;
        *samples = malloc ((size_t) *width * *height * *components);
        if (!*samples)
            return 1;
        if (ptmb_jpeg_decode (ctx, buf, size, 0, 0, *samples, 0)) {
            fprintf (stderr, "JPEG decode: %s\n", ptmb_last_error ());
            free (*samples);
            *samples = NULL;
            return 1;
        }
        return 0;
    }
    return 1;
}

int ptm_codec_write (FILE *fp, const unsigned char *samples, unsigned int width, unsigned int height,
                     unsigned int components, int is_ycbcr, int quality) {
    const size_t bytes = (size_t) width * height * components;
    if (ptm_codec_output_is_raw ()) {
        unsigned char h[16];
        raw_header (h, width, height, components);
        if (fwrite (h, 1, 16, fp) != 16 || fwrite (samples, 1, bytes, fp) != bytes)
            return 1;
        return fflush (fp) != 0;
    }
    unsigned char *jpg = NULL;
    size_t len = 0;
    if (ptmb_jpeg_encode (ptm_internal_gpu (), samples, width, height, components, is_ycbcr, quality, &jpg, &len)) {
        fprintf (stderr, "JPEG encode: %s\n", ptmb_last_error ());
        return 1;
    }
    const int bad = fwrite (jpg, 1, len, fp) != len;
    free (jpg);
    return bad || fflush (fp) != 0;
}

/* Raw image container (see ptm_codec.h). */
#include "ptm_codec.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const char MAGIC[4] = { 'R', 'A', 'W', 'J' };

int ptm_codec_read_dims (FILE *fp, ptm_image_dims_t *dims) {
    char magic[4];
    uint32_t d[3];
    if (fread (magic, 1, 4, fp) != 4 || memcmp (magic, MAGIC, 4) != 0)
        return 1;
    if (fread (d, sizeof (uint32_t), 3, fp) != 3)
        return 1;
    dims->width = d[0];
    dims->height = d[1];
    dims->components = d[2];
    return (d[0] == 0 || d[1] == 0 || d[2] == 0 || d[2] > 4) ? 1 : 0;
}

int ptm_codec_read_rows (FILE *fp, const ptm_image_dims_t *dims, unsigned char **row_pointers, unsigned int rows) {
    const size_t row_bytes = (size_t) dims->width * dims->components;
    for (unsigned int y = 0; y < rows; ++y) {
        if (fread (row_pointers[y], 1, row_bytes, fp) != row_bytes)
            return 1;
    }
    return 0;
}

int ptm_codec_write (FILE *fp, const unsigned char *samples, const ptm_image_dims_t *dims) {
    const uint32_t d[3] = { dims->width, dims->height, dims->components };
    const size_t bytes = (size_t) dims->width * dims->height * dims->components;
    if (fwrite (MAGIC, 1, 4, fp) != 4 || fwrite (d, sizeof (uint32_t), 3, fp) != 3)
        return 1;
    if (fwrite (samples, 1, bytes, fp) != bytes)
        return 1;
    return fflush (fp) != 0;
}

int ptm_codec_encode_mem (const unsigned char *samples, const ptm_image_dims_t *dims,
                          unsigned char **out, size_t *out_size) {
    const uint32_t d[3] = { dims->width, dims->height, dims->components };
    const size_t bytes = (size_t) dims->width * dims->height * dims->components;
    unsigned char *buf = malloc (16 + bytes);
    if (!buf)
        return 1;
    memcpy (buf, MAGIC, 4);
    memcpy (buf + 4, d, 12);
    memcpy (buf + 16, samples, bytes);
    *out = buf;
    *out_size = 16 + bytes;
    return 0;
}

int ptm_codec_decode_mem (const unsigned char *buf, size_t size, ptm_image_dims_t *dims, unsigned char **samples) {
    uint32_t d[3];
    if (size < 16 || memcmp (buf, MAGIC, 4) != 0)
        return 1;
    memcpy (d, buf + 4, 12);
    const size_t bytes = (size_t) d[0] * d[1] * d[2];
    if (size < 16 + bytes)
        return 1;
    dims->width = d[0];
    dims->height = d[1];
    dims->components = d[2];
    *samples = malloc (bytes ? bytes : 1);
    if (!*samples)
        return 1;
    memcpy (*samples, buf + 16, bytes);
    return 0;
}

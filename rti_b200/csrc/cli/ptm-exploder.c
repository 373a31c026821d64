/*
 * ptm-exploder (B200 build) -- same command line as the reference's rti-builder/ptm-exploder.c:
 *
 *     ptm-exploder filename.ptm filename.lp filename.out
 *
 * One relit image per "name u v" line of the .lp file (lines that do not parse to exactly three
 * fields are skipped, ref :81), written to <filename.out up to its last dot>NNN.jpeg with NNN
 * counting from 001 (ref :71-83), and the reference's progress line on stderr (ref :85).
 *
 * All directions are relit in ONE batched GPU pass (the PTM is uploaded and unscaled once);
 * the files are then written in order.
 */
#define _POSIX_C_SOURCE 200809L

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/ptm_b200.h"
#include "../../../include/ptmlib.h"
#include "../ptm_codec.h"
#include "cli_common.h"

int main (int argc, char *argv[]) {
    if (argc != 4) {
        fprintf (stderr, "Usage: %s filename.ptm filename.lp filename.out\n", argv[0]);
        fprintf (stderr, "       Explodes a PTM into JPEGs\n");
        return 1;
    }
    cli_select_devices ();
    FILE *fp = fopen (argv[1], "rb");
    if (fp == NULL) {
        fprintf (stderr, "can't open %s\n", argv[1]);
        return 1;
    }
    ptm_header_t *header = ptm_read_header (fp);
    if (header == NULL)
        return 1;
    ptm_block_t *blocks = ptm_alloc_blocks (header);
    ptm_read_ptm (fp, header, blocks);
    fclose (fp);

    FILE *fp_lp = fopen (argv[2], "rb");
    if (fp_lp == NULL) {
        fprintf (stderr, "can't open %s\n", argv[2]);
        return 1;
    }
    lp_entry_t *entries = NULL;
    const size_t n = lp_read (fp_lp, 1, &entries);
    fclose (fp_lp);
    float *uv = malloc ((n ? n : 1) * 2 * sizeof (float));
    for (size_t i = 0; i < n; ++i) {
        uv[2 * i] = entries[i].u;
        uv[2 * i + 1] = entries[i].v;
        free (entries[i].name);
    }
    free (entries);

    char *filename = malloc (strlen (argv[3]) + 20);
    strcpy (filename, argv[3]);
    char *ext = strrchr (filename, '.');
    if (!ext)
        ext = filename + strlen (argv[3]);

    const size_t pixels = header->dimen[0] * header->dimen[1];
    /* batches bound host memory: about 1 GiB of rasters at a time */
    size_t batch = pixels ? (1u << 30) / (pixels * 3) : 1;
    if (batch < 1)
        batch = 1;
    JSAMPLE *rasters = malloc ((batch < n ? batch : (n ? n : 1)) * pixels * 3 + 1);
    for (size_t d0 = 0; d0 < n; d0 += batch) {
        const size_t nd = n - d0 < batch ? n - d0 : batch;
        ptm_relight_batch (header, blocks, uv + 2 * d0, (unsigned int) nd, rasters);
        for (size_t d = 0; d < nd; ++d) {
            sprintf (ext, "%03d.jpeg", (int) (d0 + d + 1));
            fprintf (stderr, "writing %s %f %f ...\n", filename, (double) uv[2 * (d0 + d)], (double) uv[2 * (d0 + d) + 1]);
            fflush (stderr);
            FILE *out = fopen (filename, "wb");
            if (out == NULL) {
                fprintf (stderr, "can't open %s\n", filename);
                return 1;
            }
            if (ptm_codec_write (out, rasters + d * pixels * 3, (unsigned int) header->dimen[0],
                                 (unsigned int) header->dimen[1], 3, 0, header->compression_param[0]))
                fprintf (stderr, "cannot write image\n");
            fclose (out);
        }
    }
    free (rasters);
    free (uv);
    free (filename);
    ptm_free_blocks (header, blocks);
    free (header);
    return 0;
}

/*
 * ptm-decoder (B200 build) -- same command line as the reference's rti-builder/ptm-decoder.c:
 *
 *     ptm-decoder filename.ptm [u v] > image
 *
 * Relights the PTM for light (u, v) (default 0 0, ref :46-47) on the GPU and writes the image to
 * stdout through ptm_write_jpeg.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>

#include "../../../include/ptmlib.h"
#include "cli_common.h"

int main (int argc, char *argv[]) {
    if (argc != 2 && argc != 4) {
        fprintf (stderr, "Usage: %s filename.ptm [u v]\n", argv[0]);
        fprintf (stderr, "       Outputs JPEG to stdout\n");
        return 1;
    }
    float u = 0.0f, v = 0.0f;
    if (argc == 4) {
        sscanf (argv[2], "%f", &u);
        sscanf (argv[3], "%f", &v);
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

    ptm_write_jpeg (stdout, header, blocks, u, v);

    ptm_free_blocks (header, blocks);
    free (header);
    return 0;
}

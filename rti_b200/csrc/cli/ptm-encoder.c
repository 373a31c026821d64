/*
 * ptm-encoder (B200 build) -- same command line as the reference's rti-builder/ptm-encoder.c:
 *
 *     ptm-encoder [-f FORMAT] [-l] [-o FILE] [-v] FILENAME.LP
 *
 * FILENAME.LP lists "filename u v [w]" lines (paths relative to the .lp file); any line with
 * fewer than three fields, such as a leading count, is skipped (ref :152-157).  At least 12
 * images are required (ref :204).  Default format PTM_FORMAT_JPEG_RGB, default output stdout
 * (ref :120-121).  -v prints the reference's phase lines on stderr (wall-clock milliseconds here;
 * the reference prints summed CPU time).
 *
 * What differs from the reference's main: the decoded stack goes through ONE fused GPU pass
 * (ptm_encode_stack) instead of fit / average / colour loop / scale, and images are read
 * through ptm_codec.h (lossless raw container in this image).
 */
#define _POSIX_C_SOURCE 200809L

#include <argp.h>
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../../include/ptmlib.h"
#include "../ptm_codec.h"
#include "cli_common.h"

const char *argp_program_version = "PTM Encoder 0.1 (B200)";

static struct argp_option options[] = {
    { "format",  'f', "FORMAT", 0, "Which PTM format to output (default: PTM_FORMAT_JPEG_RGB).", 0 },
    { "list",    'l', 0,        0, "List supported PTM formats.",                                0 },
    { "output",  'o', "FILE",   0, "Output to FILE instead of STDOUT.",                          1 },
    { "verbose", 'v', 0,        0, "Produce verbose output.",                                    2 },
    { 0 }
};

struct arguments {
    const ptm_format_t *format;
    const char *filename_lp;
    const char *filename_ptm;
    int verbose;
};

/* option handling; messages and exit codes as the reference prints them (ref :64-104) */
static void list_formats_and_exit (void) {
    printf ("Supported formats:\n");
    for (const ptm_format_t *f = ptm_formats; f->name; ++f)
        printf ("%s\n", f->name);
    exit (0);
}

static error_t parse_opt (int key, char *arg, struct argp_state *state) {
    struct arguments *a = state->input;
    if (key == 'l')
        list_formats_and_exit ();
    if (key == 'v') {
        a->verbose = 1;
    } else if (key == 'o') {
        a->filename_ptm = arg;
    } else if (key == 'f') {
        if (!(a->format = ptm_get_format (arg))) {
            fprintf (stderr, "No format by that name: %s\n", arg);
            exit (0);          /* the reference exits 0 here too (ref :69-72) */
        }
    } else if (key == ARGP_KEY_ARG) {
        if (state->arg_num > 0)
            argp_usage (state);
        a->filename_lp = arg;
    } else if (key == ARGP_KEY_END) {
        if (state->arg_num == 0)
            argp_usage (state);
    } else {
        return ARGP_ERR_UNKNOWN;
    }
    return 0;
}

static struct argp argp = { options, parse_opt, "FILENAME.LP", "Encode a series of JPEG files into a PTM file.",
                            NULL, NULL, NULL };

static double now_ms (void) {
    struct timespec ts;
    clock_gettime (CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

int main (int argc, char *argv[]) {
    struct arguments args;
    args.verbose = 0;
    args.format = ptm_get_format ("PTM_FORMAT_JPEG_RGB");
    args.filename_ptm = "-";
    argp_parse (&argp, argc, argv, 0, 0, &args);

    cli_select_devices ();
    double t0 = now_ms ();

    /* 1. the light list, 2. one image header per entry (paths are relative to the .lp file) */
    FILE *fp_lp = fopen (args.filename_lp, "rb");
    if (fp_lp == NULL) {
        fprintf (stderr, "can't open %s\n", args.filename_lp);
        return 1;
    }
    lp_entry_t *entries = NULL;
    const size_t n = lp_read (fp_lp, 0, &entries);
    fclose (fp_lp);

    ptm_header_t *header = ptm_alloc_header ();
    header->format = args.format;
    decoder_t **decoders = malloc ((n ? n : 1) * sizeof (*decoders));
    ptm_image_dims_t *all_dims = malloc ((n ? n : 1) * sizeof (*all_dims));
    ptm_image_dims_t dims0;
    memset (&dims0, 0, sizeof (dims0));
    char *lp_copy = strdup (args.filename_lp);
    const char *dir = dirname (lp_copy);
    for (size_t i = 0; i < n; ++i) {
        char *path = malloc (strlen (dir) + strlen (entries[i].name) + 2);
        sprintf (path, "%s/%s", dir, entries[i].name);
        FILE *fp = fopen (path, "rb");
        if (fp == NULL) {
            fprintf (stderr, "can't open %s\n", path);
            return 1;
        }
        free (path);
        if (ptm_codec_read_dims (fp, &all_dims[i])) {
            fprintf (stderr, "can't read image header of %s\n", entries[i].name);
            return 1;
        }
        if (i == 0)
            dims0 = all_dims[0];
        if (all_dims[i].width != dims0.width || all_dims[i].height != dims0.height ||
            all_dims[i].components != dims0.components) {
            fprintf (stderr, "%s: all images must have the same size\n", entries[i].name);   /* ref :238-243 asserts */
            return 1;
        }
        decoder_t *d = calloc (1, sizeof (*d));
        d->fp = fp;
        d->filename = entries[i].name;
        d->u = entries[i].u;
        d->v = entries[i].v;
        d->w = entries[i].w;
        d->dinfo.output_width = all_dims[i].width;
        d->dinfo.output_height = all_dims[i].height;
        d->dinfo.output_components = (int) all_dims[i].components;
        decoders[i] = d;
    }
    free (entries);
    free (lp_copy);

    if (n < 12) {
        fprintf (stderr, "not enough jpegs %u < %d\n", (unsigned) n, 12);
        return 1;
    }
    if (args.verbose)
        fprintf (stderr, "%u JPEGs opened ...\n", (unsigned) n);

    float *M = ptm_svd (decoders, (int) n);
    if (M == NULL) {
        fprintf (stderr, "Error in Singular Value Decomposition\n");
        return 1;
    }
    if (args.verbose)
        fprintf (stderr, "done SVD\n");

    if (dims0.components != 3) {
        fprintf (stderr, "input images must have 3 components\n");
        return 1;
    }
    ptm_image_info_t info;
    info.n_decoders = (unsigned int) n;
    info.width = dims0.width;
    info.height = dims0.height;
    info.pixels = info.width * info.height;
    info.row_stride = info.width * dims0.components;
    info.decoder_stride = info.height * info.row_stride;
    header->dimen[0] = info.width;
    header->dimen[1] = info.height;

    ptm_block_t *blocks = ptm_alloc_blocks (header);
    double t1 = now_ms (), t2;
    int all_jpeg = 1;
    for (size_t i = 0; i < n; ++i)
        all_jpeg = all_jpeg && all_dims[i].kind == PTM_CODEC_JPEG;
    int encoded = 0;
    if (all_jpeg && header->format->color_components != 2) {
        /* JPEG inputs: decode on the GPU straight into the device-resident stack */
        const unsigned char **jp = malloc (n * sizeof (*jp));
        size_t *ln = malloc (n * sizeof (*ln));
        for (size_t i = 0; i < n; ++i) {
            jp[i] = all_dims[i].data;
            ln[i] = all_dims[i].len;
        }
        encoded = ptm_encode_jpeg_files (header, &info, jp, ln, M, blocks) == 0;
        free (jp);
        free (ln);
        if (encoded && args.verbose)
            fprintf (stderr, "time for decoding %u JPEGs = %lums\n", (unsigned) n, 0ul);
    }
    if (!encoded) {
        /* decode every image into the host stack, bottom row first (ref :245-264) */
        JSAMPLE *buffer = calloc (n * info.decoder_stride, sizeof (JSAMPLE));
        int failed = 0;
        #pragma omp parallel for schedule(dynamic)
        for (size_t i = 0; i < n; ++i) {
            unsigned char **rows = calloc (info.height, sizeof (*rows));
            for (size_t y = 0; y < info.height; ++y)
                rows[y] = buffer + i * info.decoder_stride + (info.height - y - 1) * info.row_stride;
            /* YCbCr for the LUM / LRGB formats, RGB for the RGB formats (ref :186-188) */
            if (ptm_codec_read_rows (decoders[i]->fp, &all_dims[i], rows, (unsigned int) info.height,
                                     header->format->color_components == 0))
                failed = 1;
            free (rows);
        }
        if (failed) {
            fprintf (stderr, "short read while decoding the input images\n");
            return 1;
        }
        t1 = now_ms ();
        if (args.verbose)
            fprintf (stderr, "time for decoding %u JPEGs = %lums\n", (unsigned) n, (unsigned long) (t1 - t0));
        ptm_encode_stack (header, &info, buffer, M, blocks);
        free (buffer);
    }
    for (size_t i = 0; i < n; ++i) {
        ptm_codec_release (&all_dims[i]);
        fclose (decoders[i]->fp);
    }
    t2 = now_ms ();
    if (args.verbose) {
        /* the reference prints two lines (fit, scale); the fused pass covers both */
        fprintf (stderr, "time for ptm_fit_poly_jsample = %lums\n", (unsigned long) (t2 - t1));
        fprintf (stderr, "time for ptm_scale_coefficients = %lums\n", 0ul);
    }

    FILE *fp_ptm = stdout;
    if (strcmp (args.filename_ptm, "-") != 0) {
        fp_ptm = fopen (args.filename_ptm, "wb");
        if (fp_ptm == NULL) {
            fprintf (stderr, "can't open %s\n", args.filename_ptm);
            return 1;
        }
    }
    ptm_write_ptm (fp_ptm, header, blocks);
    double t3 = now_ms ();
    if (args.verbose)
        fprintf (stderr, "time for ptm_write_ptm = %lums\n", (unsigned long) (t3 - t2));
    fclose (fp_ptm);

    free (M);
    ptm_free_blocks (header, blocks);
    for (size_t i = 0; i < n; ++i) {
        free (decoders[i]->filename);
        free (decoders[i]);
    }
    free (decoders);
    free (all_dims);
    free (header);
    return 0;
}

/*
 * Shared by the three command-line mains (B200 build).
 *
 *   cli_select_devices   a one-shot command pays for CUDA start-up on every run and the driver
 *                        initialises every visible GPU (about 1 s on an 8-GPU box), so the mains
 *                        narrow CUDA_VISIBLE_DEVICES to the GPUs they will use -- PTM_B200_DEVICES
 *                        ("0,1,2,3": row slabs of one encode across several GPUs) or PTM_B200_DEVICE
 *                        (default 0) -- BEFORE their first CUDA call, and renumber the selection from 0.
 *                        This lives in the mains, not in libptm.so: a library must not edit the
 *                        environment of a host process that may already run CUDA.
 *   lp_read              the .lp light-position list: "name u v [w]" per line; lines with fewer than
 *                        `min_fields` fields (e.g. the leading count) are skipped
 *                        (rti-builder/ptm-encoder.c:152-157, rti-builder/ptm-exploder.c:81).
 */
#ifndef PTM_CLI_COMMON_H
#define PTM_CLI_COMMON_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline void cli_select_devices (void) {
    if (getenv ("CUDA_VISIBLE_DEVICES"))
        return;   /* the user's own restriction stands; ordinals are relative to it */
    const char *many = getenv ("PTM_B200_DEVICES");
    const char *one = getenv ("PTM_B200_DEVICE");
    if (many && *many) {
        setenv ("CUDA_VISIBLE_DEVICES", many, 1);
        /* renumber: the n selected GPUs are now 0 .. n-1 */
        size_t n = 1;
        for (const char *p = many; *p; ++p)
            n += *p == ',';
        char *list = malloc (n * 12 + 1), *w = list;
        for (size_t i = 0; i < n; ++i)
            w += sprintf (w, i ? ",%u" : "%u", (unsigned) i);
        setenv ("PTM_B200_DEVICES", list, 1);
        free (list);
    } else {
        setenv ("CUDA_VISIBLE_DEVICES", one && *one ? one : "0", 1);
        setenv ("PTM_B200_DEVICE", "0", 1);
    }
}

typedef struct {
    char *name;
    float u, v, w;
    int fields;     /* 3 or 4 */
} lp_entry_t;

/* Reads every usable line of an .lp stream.  exact_three: accept only lines that parse to exactly
 * "name u v" (the exploder's rule); otherwise three or four fields (the encoder's rule).
 * Returns the number of entries; *out is malloc'd (free each name, then the array). */
static inline size_t lp_read (FILE *fp, int exact_three, lp_entry_t **out) {
    size_t cap = 64, n = 0;
    lp_entry_t *e = malloc (cap * sizeof (*e));
    char *line = NULL;
    size_t len = 0;
    while (getline (&line, &len, fp) != -1) {
        lp_entry_t cur = { malloc (len + 1), 0.0f, 0.0f, 0.0f, 0 };
        cur.fields = exact_three ? sscanf (line, "%s %f %f", cur.name, &cur.u, &cur.v)
                                 : sscanf (line, "%s %f %f %f", cur.name, &cur.u, &cur.v, &cur.w);
        if (exact_three ? cur.fields != 3 : cur.fields < 3) {
            free (cur.name);
            continue;
        }
        if (n == cap)
            e = realloc (e, (cap *= 2) * sizeof (*e));
        e[n++] = cur;
    }
    free (line);
    *out = e;
    return n;
}

#endif

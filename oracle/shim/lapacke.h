/* Minimal hand-written LAPACKE declarations (test infrastructure only):
 * the thin SVD driver the PTM pseudo-inverse uses. */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H

typedef int lapack_int;

#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102

lapack_int LAPACKE_sgesdd (int matrix_layout, char jobz, lapack_int m, lapack_int n,
                           float *a, lapack_int lda, float *s,
                           float *u, lapack_int ldu, float *vt, lapack_int ldvt);

#endif

/* The three BLAS / LAPACKE entry points ptm_svd needs (rti-builder/ptmlib.c:657-684),
 * declared by hand because the image ships an OpenBLAS shared object but no headers.
 * On a system with liblapacke-dev / libopenblas-dev these match <lapacke.h> / <cblas.h>. */
#ifndef PTM_LAPACK_H
#define PTM_LAPACK_H

#define PTM_LAPACK_ROW_MAJOR 101
#define PTM_CBLAS_ROW_MAJOR  101
#define PTM_CBLAS_NO_TRANS   111
#define PTM_CBLAS_TRANS      112

int LAPACKE_sgesdd (int matrix_layout, char jobz, int m, int n, float *a, int lda, float *s,
                    float *u, int ldu, float *vt, int ldvt);

void cblas_sgemm (int order, int transa, int transb, int m, int n, int k, float alpha,
                  const float *a, int lda, const float *b, int ldb, float beta, float *c, int ldc);

#endif

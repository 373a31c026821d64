/* Minimal hand-written CBLAS declarations (test infrastructure only).
 *
 * The image has an OpenBLAS shared object but no cblas.h.  These are the two
 * entry points plus the enums the PTM path needs (sgemv per pixel, sgemm for
 * the pseudo-inverse).  Values are the standard CBLAS ones. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H

enum CBLAS_ORDER     { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };

void cblas_sgemv (enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans,
                  int m, int n, float alpha, const float *a, int lda,
                  const float *x, int incx, float beta, float *y, int incy);

void cblas_sgemm (enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa,
                  enum CBLAS_TRANSPOSE transb, int m, int n, int k,
                  float alpha, const float *a, int lda, const float *b, int ldb,
                  float beta, float *c, int ldc);

/* OpenBLAS extras used by the benchmark harness to report what it ran on. */
char *openblas_get_config (void);
char *openblas_get_corename (void);
void openblas_set_num_threads (int n);

#endif

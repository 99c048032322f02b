/*
 * TEST INFRASTRUCTURE ONLY (oracle/): stand-in for libmkl_intel_lp64.so.
 *
 * The reference's prebuilt KREG.so was linked against Intel MKL 2019
 * (mlatom/fortran/compile.py:7) and imports seven LAPACK symbols
 * (call sites mlatom/fortran/mathUtils.f90:65,79,94,146,158,175):
 *   dpotrf_ dpotrs_ dporfs_ dsytrf_ dsytrs_ dsyrfs_ ilaenv_
 * MKL is not in this image.  This shim forwards them to the OpenBLAS that SciPy bundles
 * (symbols carry a `scipy_` prefix).  The library path is taken from $KREG_REF_LAPACK
 * (set by oracle/ref_kreg.py) and opened lazily with dlopen.
 *
 * Consequence (stated in DESIGN.md): alpha from the "reference" is OpenBLAS', not MKL's.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

static void *lapack_handle(void) {
  static void *h = NULL;
  if (!h) {
    const char *p = getenv("KREG_REF_LAPACK");
    if (!p) {
      fprintf(stderr, "mkl_lapack_shim: KREG_REF_LAPACK is not set\n");
      abort();
    }
    h = dlopen(p, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
      fprintf(stderr, "mkl_lapack_shim: cannot open %s: %s\n", p, dlerror());
      abort();
    }
  }
  return h;
}

static void *lapack_sym(const char *name) {
  void *s = dlsym(lapack_handle(), name);
  if (!s) {
    fprintf(stderr, "mkl_lapack_shim: missing symbol %s\n", name);
    abort();
  }
  return s;
}

#define RESOLVE(type, var, name) \
  static type var = NULL;        \
  if (!var) var = (type)lapack_sym(name)

void dpotrf_(char *uplo, int *n, double *a, int *lda, int *info, size_t l1) {
  typedef void (*fn_t)(char *, int *, double *, int *, int *, size_t);
  RESOLVE(fn_t, f, "scipy_dpotrf_");
  f(uplo, n, a, lda, info, l1);
}

void dpotrs_(char *uplo, int *n, int *nrhs, double *a, int *lda, double *b, int *ldb, int *info,
             size_t l1) {
  typedef void (*fn_t)(char *, int *, int *, double *, int *, double *, int *, int *, size_t);
  RESOLVE(fn_t, f, "scipy_dpotrs_");
  f(uplo, n, nrhs, a, lda, b, ldb, info, l1);
}

void dporfs_(char *uplo, int *n, int *nrhs, double *a, int *lda, double *af, int *ldaf, double *b,
             int *ldb, double *x, int *ldx, double *ferr, double *berr, double *work, int *iwork,
             int *info, size_t l1) {
  typedef void (*fn_t)(char *, int *, int *, double *, int *, double *, int *, double *, int *,
                       double *, int *, double *, double *, double *, int *, int *, size_t);
  RESOLVE(fn_t, f, "scipy_dporfs_");
  f(uplo, n, nrhs, a, lda, af, ldaf, b, ldb, x, ldx, ferr, berr, work, iwork, info, l1);
}

void dsytrf_(char *uplo, int *n, double *a, int *lda, int *ipiv, double *work, int *lwork,
             int *info, size_t l1) {
  typedef void (*fn_t)(char *, int *, double *, int *, int *, double *, int *, int *, size_t);
  RESOLVE(fn_t, f, "scipy_dsytrf_");
  f(uplo, n, a, lda, ipiv, work, lwork, info, l1);
}

void dsytrs_(char *uplo, int *n, int *nrhs, double *a, int *lda, int *ipiv, double *b, int *ldb,
             int *info, size_t l1) {
  typedef void (*fn_t)(char *, int *, int *, double *, int *, int *, double *, int *, int *,
                       size_t);
  RESOLVE(fn_t, f, "scipy_dsytrs_");
  f(uplo, n, nrhs, a, lda, ipiv, b, ldb, info, l1);
}

void dsyrfs_(char *uplo, int *n, int *nrhs, double *a, int *lda, double *af, int *ldaf, int *ipiv,
             double *b, int *ldb, double *x, int *ldx, double *ferr, double *berr, double *work,
             int *iwork, int *info, size_t l1) {
  typedef void (*fn_t)(char *, int *, int *, double *, int *, double *, int *, int *, double *,
                       int *, double *, int *, double *, double *, double *, int *, int *, size_t);
  RESOLVE(fn_t, f, "scipy_dsyrfs_");
  f(uplo, n, nrhs, a, lda, af, ldaf, ipiv, b, ldb, x, ldx, ferr, berr, work, iwork, info, l1);
}

int ilaenv_(int *ispec, char *name, char *opts, int *n1, int *n2, int *n3, int *n4, size_t l1,
            size_t l2) {
  typedef int (*fn_t)(int *, char *, char *, int *, int *, int *, int *, size_t, size_t);
  RESOLVE(fn_t, f, "scipy_ilaenv_");
  return f(ispec, name, opts, n1, n2, n3, n4, l1, l2);
}

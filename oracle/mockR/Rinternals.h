/* Mock of the handful of R C-API names used by the kmer_spans .Call glue.
 *
 * TEST INFRASTRUCTURE ONLY.  R is not installed in the build image, so both the
 * UNMODIFIED reference source (/root/reference/src/kmer_spans.c, compiled where it
 * lies into oracle/_ref/) and this repo's replacement glue (r/src/kmer_spans_glue.c)
 * are compiled against this header for parity tests of the .Call boundary.
 *
 * Every vector is calloc'ed with MOCKR_PAD trailing zero bytes, which pins the two
 * undefined behaviours of the reference (rank buffer not zeroed,
 * T10: read past the string terminator) to "as if the memory were zero".
 */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOCKR_PAD 64

#define CHARSXP 9
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19

typedef struct mock_sexp {
  int type;
  long len;      /* number of elements */
  int nrow, ncol;
  void *data;    /* int* / double* / char* / struct mock_sexp** */
} *SEXP;

typedef void *(*DL_FUNC)(void);
typedef struct { int unused; } DllInfo;
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;

int TYPEOF(SEXP x);
int length(SEXP x);
int *INTEGER(SEXP x);
double *REAL(SEXP x);
const char *CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, long i);
SEXP VECTOR_ELT(SEXP x, long i);
SEXP SET_VECTOR_ELT(SEXP x, long i, SEXP v);
void SET_STRING_ELT(SEXP x, long i, SEXP v);
SEXP allocVector(int type, long n);
SEXP allocMatrix(int type, int nrow, int ncol);
SEXP mkChar(const char *s);
SEXP mkCharLen(const char *s, long n);
int asInteger(SEXP x);
void error(const char *fmt, ...);
void Rprintf(const char *fmt, ...);
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call,
                       const void *f, const void *e);
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))

/* harness helpers (not part of R) */
void mockR_free_all(void);
const char *mockR_last_error(void);
/* run fn(args...) under setjmp; returns NULL if error() was raised */
SEXP mockR_call(DL_FUNC fn, int nargs, SEXP *args);
const R_CallMethodDef *mockR_registered(void);

#ifdef __cplusplus
}
#endif
#endif

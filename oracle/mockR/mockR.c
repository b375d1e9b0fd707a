/* Implementation of the mock R API declared in Rinternals.h (test infrastructure). */
#include "Rinternals.h"
#include <stdarg.h>
#include <setjmp.h>

static void **g_allocs = NULL;
static size_t g_nallocs = 0, g_cap = 0;
static char g_err[1024];
static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static const R_CallMethodDef *g_registered = NULL;

static void *track(void *p) {
  if (g_nallocs == g_cap) {
    g_cap = g_cap ? g_cap * 2 : 1024;
    g_allocs = (void **)realloc(g_allocs, g_cap * sizeof(void *));
  }
  g_allocs[g_nallocs++] = p;
  return p;
}

void mockR_free_all(void) {
  for (size_t i = 0; i < g_nallocs; ++i) free(g_allocs[i]);
  g_nallocs = 0;
}

const char *mockR_last_error(void) { return g_err; }
const R_CallMethodDef *mockR_registered(void) { return g_registered; }

static size_t elt_size(int type) {
  switch (type) {
    case INTSXP: return sizeof(int);
    case REALSXP: return sizeof(double);
    case STRSXP: case VECSXP: return sizeof(SEXP);
    case CHARSXP: return 1;
    default: return 1;
  }
}

SEXP allocVector(int type, long n) {
  SEXP s = (SEXP)track(calloc(1, sizeof(struct mock_sexp)));
  s->type = type;
  s->len = n;
  s->nrow = (int)n;
  s->ncol = 1;
  s->data = track(calloc(1, (size_t)n * elt_size(type) + MOCKR_PAD));
  return s;
}

SEXP allocMatrix(int type, int nrow, int ncol) {
  SEXP s = allocVector(type, (long)nrow * (long)ncol);
  s->nrow = nrow;
  s->ncol = ncol;
  return s;
}

SEXP mkCharLen(const char *str, long n) {
  SEXP s = allocVector(CHARSXP, n);
  memcpy(s->data, str, (size_t)n);
  return s;
}
SEXP mkChar(const char *str) { return mkCharLen(str, (long)strlen(str)); }

int TYPEOF(SEXP x) { return x->type; }
int length(SEXP x) { return (int)x->len; }
int *INTEGER(SEXP x) { return (int *)x->data; }
double *REAL(SEXP x) { return (double *)x->data; }
const char *CHAR(SEXP x) { return (const char *)x->data; }
SEXP STRING_ELT(SEXP x, long i) { return ((SEXP *)x->data)[i]; }
SEXP VECTOR_ELT(SEXP x, long i) { return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, long i, SEXP v) { ((SEXP *)x->data)[i] = v; return v; }
void SET_STRING_ELT(SEXP x, long i, SEXP v) { ((SEXP *)x->data)[i] = v; }
int asInteger(SEXP x) {
  if (x->type == INTSXP) return INTEGER(x)[0];
  if (x->type == REALSXP) return (int)REAL(x)[0];
  return 0;
}

void error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  if (g_jmp_armed) longjmp(g_jmp, 1);
  fprintf(stderr, "mockR error(): %s\n", g_err);
  abort();
}

void Rprintf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
}

int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call,
                       const void *f, const void *e) {
  (void)info; (void)c; (void)f; (void)e;
  g_registered = call;
  return 1;
}

typedef SEXP (*fn1)(SEXP);
typedef SEXP (*fn2)(SEXP, SEXP);
typedef SEXP (*fn5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

SEXP mockR_call(DL_FUNC fn, int nargs, SEXP *a) {
  SEXP r = NULL;
  g_err[0] = 0;
  g_jmp_armed = 1;
  if (setjmp(g_jmp) == 0) {
    switch (nargs) {
      case 1: r = ((fn1)fn)(a[0]); break;
      case 2: r = ((fn2)fn)(a[0], a[1]); break;
      case 5: r = ((fn5)fn)(a[0], a[1], a[2], a[3], a[4]); break;
      case 8: r = ((fn8)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
      default: snprintf(g_err, sizeof g_err, "mockR_call: unsupported arity %d", nargs);
    }
  } else {
    r = NULL;
  }
  g_jmp_armed = 0;
  return r;
}

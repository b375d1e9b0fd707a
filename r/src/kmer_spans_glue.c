/* kmer_spans_glue.c -- .Call glue that binds the B200 C ABI (include/kspans.h) into R.
 *
 * Drop-in for the hot path of lmjakt/kmer_spans: builds into kmer_spans.so, exports
 * R_init_kmer_spans and registers the same six .Call names with the same arities as the
 * reference (src/kmer_spans.c:795-808), so the reference's kmer_spans.R works unchanged:
 *
 *   kmer_counts/2            -> ks_kmer_counts              (reference :453-487)
 *   kmer_regions_r/5         -> ks_kmer_regions             (reference :490-546)
 *   kmer_low_comp_regions/5  -> ks_kmer_low_comp_regions    (reference :548-621)
 *   kmer_seq_r/1             -> ks_kmer_seq (host only)     (reference :623-639)
 *   tr_lr_regions_r/5        -> ks_tr_lr_regions            (reference :649-713)
 *   windowed_kmer_count_distributions_r/5 -> ks_windowed_kmer_count_distributions (reference :715-793)
 *   kmer_mode_regions/8 (extension): counts -> scores(mode) -> scan resident on the GPU.
 *
 * Argument checks, their order and the error texts follow the reference; result lists have the
 * same element order and types (the R wrappers name them positionally, kmer_spans.R:21,50,75).
 * Several GPUs: with KSPANS_DEVICES=0,1,... (two or more entries) kmer_counts, kmer_low_comp_regions and
 * kmer_mode_regions run on a ks_mctx -- one process, the sequences of the call sharded over the listed devices,
 * count tables summed over NVLink peer memory, spans that cross a cut stitched exactly (include/kspans.h).  The
 * other entry points use the first listed device.
 * CUDA is initialised lazily inside the first call (never at dyn.load) and re-initialised in a
 * forked child (mclapply), see SURVEY.md section 8b.  error() is only raised after every ks_*
 * resource has been released: nothing longjmps across the CUDA library.
 *
 * R is not installed in the build image; this file is compiled against the mock R API in
 * oracle/mockR for the boundary tests (r/Makefile) and against the real R headers by R CMD SHLIB
 * (r/src/Makevars).
 */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "kspans.h"

#define GLUE_MAX_K 16

static ks_ctx *g_ctx = NULL;
static ks_mctx *g_mctx = NULL;
static pid_t g_pid = 0, g_mpid = 0;
static char g_msg[600];

/* device list of KSPANS_DEVICES ("0,1,2"); returns the number of entries (0: unset) */
static int glue_devices(int *dev, int cap) {
  const char *e = getenv("KSPANS_DEVICES");
  int n = 0;
  if (!e) return 0;
  while (*e && n < cap) {
    char *end = NULL;
    long v = strtol(e, &end, 10);
    if (end == e) break;
    dev[n++] = (int)v;
    e = (*end == ',') ? end + 1 : end;
    if (*end && *end != ',') break;
  }
  return n;
}

/* multi-device context of this process when KSPANS_DEVICES names two or more devices, else NULL with g_msg
 * empty; NULL with g_msg set = creation failed */
static ks_mctx *glue_mctx(void) {
  int dev[16];
  g_msg[0] = 0;
  int n = glue_devices(dev, 16);
  if (n < 2) return NULL;
  pid_t me = getpid();
  if (g_mctx && g_mpid == me) return g_mctx;
  g_mctx = NULL; /* inherited through fork(): unusable, not ours to destroy */
  if (ks_mctx_create(&g_mctx, dev, n) != KS_OK) {
    snprintf(g_msg, sizeof g_msg, "kmer_spans (CUDA): %s", ks_mctx_last_error(NULL));
    g_mctx = NULL;
    return NULL;
  }
  g_mpid = me;
  return g_mctx;
}

/* context of this process, created on first use; a forked child gets its own */
static ks_ctx *glue_ctx(void) {
  pid_t me = getpid();
  if (g_ctx && g_pid == me) return g_ctx;
  g_ctx = NULL; /* a context inherited through fork() is unusable and must not be destroyed here */
  int dev = -1;
  const char *e = getenv("KSPANS_DEVICE");
  if (e && *e) dev = atoi(e);
  else {
    int list[16];
    if (glue_devices(list, 16) >= 1) dev = list[0];
  }
  if (ks_ctx_create(&g_ctx, dev) != KS_OK) {
    snprintf(g_msg, sizeof g_msg, "kmer_spans (CUDA): %s", ks_last_error(NULL));
    g_ctx = NULL;
    return NULL;
  }
  g_pid = me;
  return g_ctx;
}

typedef struct {
  const char **ptr;
  int64_t *len;
  int n;
} seq_view;

static int view_strsxp(SEXP seq_r, seq_view *v) {
  v->n = length(seq_r);
  v->ptr = (const char **)malloc(sizeof(char *) * (size_t)v->n);
  v->len = (int64_t *)malloc(sizeof(int64_t) * (size_t)v->n);
  if (!v->ptr || !v->len) return 0;
  for (int i = 0; i < v->n; ++i) {
    SEXP s = STRING_ELT(seq_r, i);
    v->ptr[i] = CHAR(s);
    v->len[i] = length(s);
  }
  return 1;
}
static void view_free(seq_view *v) {
  free((void *)v->ptr);
  free(v->len);
}

static void fail_from_ctx(ks_ctx *ctx) {
  snprintf(g_msg, sizeof g_msg, "%s", ctx ? ks_last_error(ctx) : "kmer_spans (CUDA): no context");
}

/* ---- kmer_counts ------------------------------------------------------------------------- */
SEXP kmer_counts(SEXP seq_r, SEXP k_r) {
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) < 1)
    error("seq_r must be a character vector of length at least one");
  if (TYPEOF(k_r) != INTSXP || length(k_r) < 1)
    error("k_r must be an integer vector of length at least one");
  int k = INTEGER(k_r)[0];
  if (k < 1 || k > GLUE_MAX_K) error("k must be a positive integer less than 1+MAX_K");
  if (k == GLUE_MAX_K) error("k = 16 overflows the reference's int table index; use k <= 15");
  size_t counts_size = (size_t)1 << (2 * k);
  SEXP ret = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(ret, 0, allocVector(REALSXP, 1));
  SET_VECTOR_ELT(ret, 1, allocVector(INTSXP, counts_size));
  double *n_counts = REAL(VECTOR_ELT(ret, 0));
  int *counts = INTEGER(VECTOR_ELT(ret, 1));
  seq_view v;
  int ok = view_strsxp(seq_r, &v);
  ks_mctx *mc = ok ? glue_mctx() : NULL;
  ks_ctx *ctx = (ok && !mc && !g_msg[0]) ? glue_ctx() : NULL;
  int rc = KS_ERR_NOMEM;
  if (ok && mc) {
    rc = ks_m_kmer_counts(mc, v.ptr, v.len, v.n, k, (int32_t *)counts, n_counts);
    if (rc) snprintf(g_msg, sizeof g_msg, "%s", ks_mctx_last_error(mc));
  } else if (ok && ctx) {
    rc = ks_kmer_counts(ctx, v.ptr, v.len, v.n, k, (int32_t *)counts, n_counts);
    if (rc) fail_from_ctx(ctx);
  } else if (!ok) {
    snprintf(g_msg, sizeof g_msg, "out of memory");
  }
  view_free(&v);
  if (rc) { UNPROTECT(1); error("%s", g_msg); }
  UNPROTECT(1);
  return ret;
}

/* copy library-owned spans into the two R matrices (3 x n integer, 2 x n double) */
static void spans_to_r(SEXP ret, int at, const ks_spans *sp) {
  SET_VECTOR_ELT(ret, at, allocMatrix(INTSXP, 3, (int)sp->n));
  SET_VECTOR_ELT(ret, at + 1, allocMatrix(REALSXP, 2, (int)sp->n));
  if (sp->n) {
    memcpy(INTEGER(VECTOR_ELT(ret, at)), sp->pos, sizeof(int) * 3 * sp->n);
    memcpy(REAL(VECTOR_ELT(ret, at + 1)), sp->score, sizeof(double) * 2 * sp->n);
  }
}

/* ---- kmer_regions_r ---------------------------------------------------------------------- */
SEXP kmer_regions_r(SEXP seq_r, SEXP k_r, SEXP kmer_w_r, SEXP min_width_r, SEXP min_score_r) {
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) < 1)
    error("seq_r must be a character vector of length at least one");
  if (TYPEOF(k_r) != INTSXP || length(k_r) < 1)
    error("k_r must be an integer vector of length at least one");
  if (TYPEOF(kmer_w_r) != REALSXP) error("kmer_w_r must be a double vector of length k^4");
  if (TYPEOF(min_width_r) != INTSXP || length(min_width_r) != 1)
    error("the minimum width must be an integer vector of length 1");
  if (TYPEOF(min_score_r) != REALSXP || length(min_score_r) != 1)
    error("the minimum score must be a REAL vector of length 1");
  int k = INTEGER(k_r)[0];
  if (k >= GLUE_MAX_K) error("kmer sizes larger than or equal to %d not currently supported", GLUE_MAX_K);
  if (k < 1) error("k must be a positive integer");
  int kmer_n = length(kmer_w_r);
  if ((unsigned int)kmer_n != (1u << (2 * k)))
    error("kmer_w contains %d elements but should have %d", kmer_n, (1 << (2 * k)));
  int min_width = INTEGER(min_width_r)[0];
  double min_score = REAL(min_score_r)[0];

  SEXP ret = PROTECT(allocVector(VECSXP, 4));
  SET_VECTOR_ELT(ret, 0, allocVector(REALSXP, 1));
  SET_VECTOR_ELT(ret, 1, allocVector(INTSXP, kmer_n));
  double *nuc = REAL(VECTOR_ELT(ret, 0));
  int *k_counts = INTEGER(VECTOR_ELT(ret, 1));
  seq_view v;
  ks_spans sp = {NULL, NULL, 0};
  int ok = view_strsxp(seq_r, &v);
  ks_ctx *ctx = ok ? glue_ctx() : NULL;
  int rc = KS_ERR_NOMEM;
  if (ok && ctx) {
    rc = ks_kmer_regions(ctx, v.ptr, v.len, v.n, k, REAL(kmer_w_r), min_width, min_score, nuc,
                         (int32_t *)k_counts, &sp);
    if (rc) fail_from_ctx(ctx);
  } else if (!ok) {
    snprintf(g_msg, sizeof g_msg, "out of memory");
  }
  view_free(&v);
  if (rc) { ks_spans_free(&sp); UNPROTECT(1); error("%s", g_msg); }
  spans_to_r(ret, 2, &sp);
  ks_spans_free(&sp);
  UNPROTECT(1);
  return ret;
}

/* ---- kmer_low_comp_regions ----------------------------------------------------------------- */
SEXP kmer_low_comp_regions(SEXP seq_r, SEXP k_r, SEXP min_width_r, SEXP min_score_r, SEXP threshold_r) {
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) < 1)
    error("seq_r must be a character vector of length at least one");
  if (TYPEOF(k_r) != INTSXP || length(k_r) < 1)
    error("k_r must be an integer vector of length at least one");
  if (TYPEOF(min_width_r) != INTSXP || length(min_width_r) != 1)
    error("the minimum width must be an integer vector of length 1");
  if (TYPEOF(min_score_r) != REALSXP || length(min_score_r) != 1)
    error("the minimum score must be a REAL vector of length 1");
  if (TYPEOF(threshold_r) != REALSXP || length(threshold_r) != 1)
    error("the threshold must be a REAL vector of length 1");
  int k = INTEGER(k_r)[0];
  int min_width = INTEGER(min_width_r)[0];
  double min_score = REAL(min_score_r)[0];
  double threshold = REAL(threshold_r)[0];
  if (threshold <= 0 || threshold >= 1) error("the threshold must be between 0 and 1");
  /* the reference performs no check on k here and is undefined outside 1..15 */
  if (k < 1 || k >= GLUE_MAX_K) error("k must be between 1 and 15");

  size_t counts_size = (size_t)1 << (2 * k);
  SEXP ret = PROTECT(allocVector(VECSXP, 5));
  SET_VECTOR_ELT(ret, 0, allocVector(REALSXP, 2));
  SET_VECTOR_ELT(ret, 1, allocVector(INTSXP, counts_size));
  SET_VECTOR_ELT(ret, 2, allocVector(REALSXP, counts_size));
  double *n_counts = REAL(VECTOR_ELT(ret, 0));
  seq_view v;
  ks_spans sp = {NULL, NULL, 0};
  int ok = view_strsxp(seq_r, &v);
  ks_mctx *mc = ok ? glue_mctx() : NULL;
  ks_ctx *ctx = (ok && !mc && !g_msg[0]) ? glue_ctx() : NULL;
  int rc = KS_ERR_NOMEM;
  if (ok && mc) {
    rc = ks_m_kmer_low_comp_regions(mc, v.ptr, v.len, v.n, k, min_width, min_score, threshold, n_counts,
                                    (int32_t *)INTEGER(VECTOR_ELT(ret, 1)), REAL(VECTOR_ELT(ret, 2)), &sp);
    if (rc) snprintf(g_msg, sizeof g_msg, "%s", ks_mctx_last_error(mc));
  } else if (ok && ctx) {
    rc = ks_kmer_low_comp_regions(ctx, v.ptr, v.len, v.n, k, min_width, min_score, threshold, n_counts,
                                  (int32_t *)INTEGER(VECTOR_ELT(ret, 1)), REAL(VECTOR_ELT(ret, 2)), &sp);
    if (rc) fail_from_ctx(ctx);
  } else if (!ok) {
    snprintf(g_msg, sizeof g_msg, "out of memory");
  }
  view_free(&v);
  if (rc) { ks_spans_free(&sp); UNPROTECT(1); error("%s", g_msg); }
  spans_to_r(ret, 3, &sp);
  ks_spans_free(&sp);
  UNPROTECT(1);
  return ret;
}

/* ---- kmer_seq_r (host only) --------------------------------------------------------------- */
SEXP kmer_seq_r(SEXP k_r) {
  if (TYPEOF(k_r) != INTSXP || length(k_r) != 1) error("k_r should be an integer of length 1");
  int k = INTEGER(k_r)[0];
  if (k > GLUE_MAX_K || k < 1)
    error("k_r (%d) should be smaller than MAX_K (%d) and larger than 0", k, GLUE_MAX_K);
  if (k == GLUE_MAX_K) error("k = 16 would need 2^32 strings; use k <= 15");
  size_t n = (size_t)1 << (2 * k);
  SEXP ret = PROTECT(allocVector(STRSXP, n));
  char buf[GLUE_MAX_K + 2];
  for (size_t i = 0; i < n; ++i) {
    ks_kmer_seq(k, (uint64_t)i, buf);
    SET_STRING_ELT(ret, i, mkChar(buf));
  }
  UNPROTECT(1);
  return ret;
}

/* ---- extension: fused modes ------------------------------------------------------------------
 * .Call("kmer_mode_regions", seq, k, mode, param, thr, min.w, min.score, want.scores)
 *   mode: 0 weighted rank - thr, 1 log2(f/f_med), 2 +-1 around f_t (param, NA = median), 3 (r - r_t)/r_t
 *   returns list(n, counts, scores (or NULL), pos 3 x R, score 2 x R)                              */
SEXP kmer_mode_regions(SEXP seq_r, SEXP k_r, SEXP mode_r, SEXP param_r, SEXP thr_r, SEXP min_width_r,
                       SEXP min_score_r, SEXP want_scores_r) {
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) < 1)
    error("seq_r must be a character vector of length at least one");
  if (TYPEOF(k_r) != INTSXP || length(k_r) < 1) error("k_r must be an integer vector of length at least one");
  if (TYPEOF(mode_r) != INTSXP || length(mode_r) != 1) error("mode must be an integer vector of length 1");
  if (TYPEOF(param_r) != REALSXP || length(param_r) != 1) error("param must be a REAL vector of length 1");
  if (TYPEOF(thr_r) != REALSXP || length(thr_r) != 1) error("the threshold must be a REAL vector of length 1");
  if (TYPEOF(min_width_r) != INTSXP || length(min_width_r) != 1)
    error("the minimum width must be an integer vector of length 1");
  if (TYPEOF(min_score_r) != REALSXP || length(min_score_r) != 1)
    error("the minimum score must be a REAL vector of length 1");
  int k = INTEGER(k_r)[0];
  if (k < 1 || k >= GLUE_MAX_K) error("k must be between 1 and 15");
  int want = (TYPEOF(want_scores_r) == INTSXP && length(want_scores_r) == 1) ? INTEGER(want_scores_r)[0] : 1;
  size_t counts_size = (size_t)1 << (2 * k);
  SEXP ret = PROTECT(allocVector(VECSXP, 5));
  SET_VECTOR_ELT(ret, 0, allocVector(REALSXP, 1));
  SET_VECTOR_ELT(ret, 1, allocVector(INTSXP, counts_size));
  SET_VECTOR_ELT(ret, 2, allocVector(REALSXP, want ? counts_size : 0));
  seq_view v;
  ks_spans sp = {NULL, NULL, 0};
  int ok = view_strsxp(seq_r, &v);
  ks_mctx *mc = ok ? glue_mctx() : NULL;
  ks_ctx *ctx = (ok && !mc && !g_msg[0]) ? glue_ctx() : NULL;
  int rc = KS_ERR_NOMEM;
  if (ok && mc) {
    rc = ks_m_kmer_mode_regions(mc, v.ptr, v.len, v.n, k, INTEGER(mode_r)[0], REAL(param_r)[0], REAL(thr_r)[0],
                                INTEGER(min_width_r)[0], REAL(min_score_r)[0], REAL(VECTOR_ELT(ret, 0)),
                                (int32_t *)INTEGER(VECTOR_ELT(ret, 1)), want ? REAL(VECTOR_ELT(ret, 2)) : NULL, &sp);
    if (rc) snprintf(g_msg, sizeof g_msg, "%s", ks_mctx_last_error(mc));
  } else if (ok && ctx) {
    rc = ks_kmer_mode_regions(ctx, v.ptr, v.len, v.n, k, INTEGER(mode_r)[0], REAL(param_r)[0], REAL(thr_r)[0],
                              INTEGER(min_width_r)[0], REAL(min_score_r)[0], REAL(VECTOR_ELT(ret, 0)),
                              (int32_t *)INTEGER(VECTOR_ELT(ret, 1)), want ? REAL(VECTOR_ELT(ret, 2)) : NULL, &sp);
    if (rc) fail_from_ctx(ctx);
  } else if (!ok) {
    snprintf(g_msg, sizeof g_msg, "out of memory");
  }
  view_free(&v);
  if (rc) { ks_spans_free(&sp); UNPROTECT(1); error("%s", g_msg); }
  spans_to_r(ret, 3, &sp);
  ks_spans_free(&sp);
  UNPROTECT(1);
  return ret;
}

/* ---- tr_lr_regions_r (reference :649-713) -------------------------------------------------------- */
SEXP tr_lr_regions_r(SEXP seq_r, SEXP params_r, SEXP kmers_r, SEXP kmer_scores_r, SEXP trans_scores_r) {
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) < 1)
    error("seq_r should be a character vector of of positive length");
  if (TYPEOF(params_r) != INTSXP || length(params_r) != 2)
    error("params_r should have two integers (k, and min_length)");
  if (TYPEOF(kmers_r) != STRSXP) error("kmers_r should be a character vector");
  if (TYPEOF(kmer_scores_r) != REALSXP || TYPEOF(trans_scores_r) != REALSXP)
    error("freq_a and freq_b should be double vectors");
  int k = INTEGER(params_r)[0];
  int min_length = INTEGER(params_r)[1];
  if (k < 1 || k > GLUE_MAX_K) error("k should be a positive value less than MAX_K");
  if (k == GLUE_MAX_K) error("k = 16 overflows the reference's int table index; use k <= 15");
  if (min_length < 0) error("min_length should be a positive integer");
  size_t kmers_size = (size_t)1 << (2 * k);
  if ((size_t)length(kmers_r) != kmers_size || (size_t)length(kmer_scores_r) != kmers_size ||
      (size_t)length(trans_scores_r) != kmers_size)
    error("kmers_r, freq_a, freq_b should all be 4^k long");
  SEXP ret = PROTECT(allocVector(VECSXP, 3));
  /* the two score vectors re-ordered to the 2-bit code of their k-mer strings (:688-696) */
  SET_VECTOR_ELT(ret, 0, allocMatrix(REALSXP, (int)kmers_size, 2));
  double *init = REAL(VECTOR_ELT(ret, 0)), *trans = init + kmers_size;
  memset(init, 0, sizeof(double) * 2 * kmers_size);
  for (size_t i = 0; i < kmers_size; ++i) {
    uint32_t code = ks_kmer_code(CHAR(STRING_ELT(kmers_r, i)), k);
    init[code] = REAL(kmer_scores_r)[i];
    trans[code] = REAL(trans_scores_r)[i];
  }
  seq_view v;
  int ok = view_strsxp(seq_r, &v);
  ks_ctx *ctx = ok ? glue_ctx() : NULL;
  ks_spans sp = {0};
  int rc = KS_ERR_NOMEM;
  if (ok && ctx) {
    rc = ks_tr_lr_regions(ctx, v.ptr, v.len, v.n, k, init, trans, min_length, &sp);
    if (rc) fail_from_ctx(ctx);
  } else if (!ok) {
    snprintf(g_msg, sizeof g_msg, "out of memory");
  }
  view_free(&v);
  if (rc) { ks_spans_free(&sp); UNPROTECT(1); error("%s", g_msg); }
  spans_to_r(ret, 1, &sp);
  ks_spans_free(&sp);
  UNPROTECT(1);
  return ret;
}

/* ---- windowed_kmer_count_distributions_r (reference :715-793) -------------------------------- */
SEXP windowed_kmer_count_distributions_r(SEXP seq_r, SEXP kmers_r, SEXP k_r, SEXP window_r, SEXP ret_flag_r) {
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) < 1)
    error("seq_r should be a character vector with at least one element");
  if (TYPEOF(kmers_r) != STRSXP || length(kmers_r) < 1)
    error("kmers_r should be a character vector with at least one element");
  if (TYPEOF(k_r) != INTSXP || length(k_r) != 1)
    error("k_r should be an integer vector with one element");
  if (TYPEOF(window_r) != INTSXP || length(window_r) != 1)
    error("window_r should be an integer vector with one element");
  if (TYPEOF(ret_flag_r) != INTSXP || length(ret_flag_r) != 1)
    error("ret_flag_r should a single integer");
  int k = asInteger(k_r);
  if (k < 0 || k >= GLUE_MAX_K)
    error("kmer sizes larger than or equal to %d not currently supported", GLUE_MAX_K);
  int kmer_n = length(kmers_r);
  for (int i = 0; i < kmer_n; ++i)
    if (length(STRING_ELT(kmers_r, i)) != k) error("All kmers specified must be of the same length");
  int window = asInteger(window_r);
  if (window < 0 || window < 2 * k) error("The window size must be at least two times k");
  if (k < 1) error("k must be a positive integer");
  int want_pos = asInteger(ret_flag_r) & 1;
  int nseq = length(seq_r);

  SEXP ret = PROTECT(allocVector(VECSXP, 3));
  SET_VECTOR_ELT(ret, 0, allocMatrix(INTSXP, window + 1, kmer_n));
  SET_VECTOR_ELT(ret, 1, allocVector(INTSXP, nseq));
  uint32_t *codes = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)kmer_n);
  int32_t **pos = NULL;
  seq_view v;
  int ok = view_strsxp(seq_r, &v) && codes;
  if (ok && want_pos) {
    /* R-owned result matrices, allocated before any CUDA work (:776-783) */
    pos = (int32_t **)calloc((size_t)nseq, sizeof(int32_t *));
    ok = pos != NULL;
    if (ok) {
      SET_VECTOR_ELT(ret, 2, allocVector(VECSXP, nseq));
      for (int i = 0; i < nseq; ++i) {
        if (v.len[i] <= window) continue;
        SET_VECTOR_ELT(VECTOR_ELT(ret, 2), i, allocMatrix(INTSXP, (int)v.len[i], kmer_n));
        pos[i] = (int32_t *)INTEGER(VECTOR_ELT(VECTOR_ELT(ret, 2), i));
      }
    }
  }
  int rc = KS_ERR_NOMEM;
  ks_ctx *ctx = ok ? glue_ctx() : NULL;
  if (ok && ctx) {
    for (int i = 0; i < kmer_n; ++i) codes[i] = ks_kmer_code(CHAR(STRING_ELT(kmers_r, i)), k);
    rc = ks_windowed_kmer_count_distributions(ctx, v.ptr, v.len, v.n, k, codes, kmer_n, window,
                                              (int32_t *)INTEGER(VECTOR_ELT(ret, 0)),
                                              (int32_t *)INTEGER(VECTOR_ELT(ret, 1)), pos);
    if (rc) fail_from_ctx(ctx);
  } else if (!ok) {
    snprintf(g_msg, sizeof g_msg, "out of memory");
  }
  view_free(&v);
  free(codes);
  free(pos);
  UNPROTECT(1);
  if (rc) error("%s", g_msg);
  return ret;
}

static const R_CallMethodDef callMethods[] = {
  {"kmer_counts", (DL_FUNC)&kmer_counts, 2},
  {"kmer_regions_r", (DL_FUNC)&kmer_regions_r, 5},
  {"kmer_low_comp_regions", (DL_FUNC)&kmer_low_comp_regions, 5},
  {"kmer_seq_r", (DL_FUNC)&kmer_seq_r, 1},
  {"tr_lr_regions_r", (DL_FUNC)&tr_lr_regions_r, 5},
  {"windowed_kmer_count_distributions_r", (DL_FUNC)&windowed_kmer_count_distributions_r, 5},
  {"kmer_mode_regions", (DL_FUNC)&kmer_mode_regions, 8},
  {NULL, NULL, 0}
};

void R_init_kmer_spans(DllInfo *info) { R_registerRoutines(info, NULL, callMethods, NULL, NULL); }

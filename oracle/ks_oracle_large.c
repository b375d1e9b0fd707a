/* ks_oracle_large.c -- CPU checker for the LARGE-k path (k = 16 .. 31, BASELINE.json configs[3]: k = 21).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as ks_oracle.c).
 *
 * Parity status: UNPINNED BY THE REFERENCE, defined by extension.  The reference cannot run k >= 16
 * (src/kmer_spans.c:37 MAX_K 16, :139 `1 << (2*k)` is an int shift, :504 rejects k >= 16), so there is
 * nothing to compare with.  This file restates the same behavioural spec with 64-bit codes and a SPARSE table:
 *   counting   the rules of sequence_kmer_count (:135-155) unchanged: every k-mer of an N-free run, except that a
 *              run of exactly k bases ending at the terminator contributes nothing (:143-144)
 *   rank       rank_kmers_w (:189-202) over the k-mers that OCCUR, in the order (count ascending, code ascending)
 *              with the sequential double accumulation rank_i = rank_{i-1} + count_{i-1} / T.  This equals what
 *              the reference's formula would give on the full 4^k table: the absent k-mers sort first (count 0,
 *              code order -- glibc's stable mergesort) and each adds 0 / T = 0 to the running sum.
 *   scan       kmer_regions (:243-307) unchanged, w = rank[code] - thr.
 * What pins it: for k <= 15 the same functions must reproduce ks_oracle.c (which IS pinned to the compiled
 * reference) bit for bit -- tests/test_oracle.py::test_large_oracle_equals_pinned_oracle_at_small_k.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t *pos;
  double *score;
  size_t n, cap;
} kso_spans;

static void spans_push(kso_spans *s, int32_t id, int32_t beg, int32_t end, double score) {
  if (s->n == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 128;
    s->pos = (int32_t *)realloc(s->pos, s->cap * 3 * sizeof(int32_t));
    s->score = (double *)realloc(s->score, s->cap * 2 * sizeof(double));
  }
  s->pos[3 * s->n] = id;
  s->pos[3 * s->n + 1] = beg;
  s->pos[3 * s->n + 2] = end;
  s->score[2 * s->n] = score;
  s->score[2 * s->n + 1] = 0.0;
  s->n++;
}

static inline uint64_t base2(unsigned char c) { return (c >> 1) & 3u; }
static inline int is_n(unsigned char c) { return (c | 0x20) == 'n'; }
static uint64_t code_ending_at(const char *seq, int64_t e, int k) {
  uint64_t c = 0;
  for (int64_t j = e - k + 1; j <= e; ++j) c = (c << 2) | base2((unsigned char)seq[j]);
  return c;
}
static int next_run(const char *seq, int64_t len, int64_t from, int64_t *a, int64_t *b) {
  int64_t i = from;
  while (i < len && is_n((unsigned char)seq[i])) ++i;
  if (i >= len) return 0;
  *a = i;
  while (i < len && !is_n((unsigned char)seq[i])) ++i;
  *b = i;
  return 1;
}
static int64_t c_strlen_bounded(const char *seq, int64_t len) {
  const void *z = memchr(seq, 0, (size_t)len);
  return z ? (int64_t)((const char *)z - seq) : len;
}

typedef struct {
  uint64_t *codes;   /* distinct k-mers, ascending */
  uint32_t *counts;
  double *ranks;     /* aligned with codes */
  size_t nd;
  double total;
} kso_large;

void kso_large_free(kso_large *t) { free(t->codes); free(t->counts); free(t->ranks); memset(t, 0, sizeof *t); }

static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}

/* all counted k-mers of the call -> sorted distinct list with multiplicities */
int kso_large_count(const char *const *seqs, const int64_t *lens, int nseq, int k, kso_large *t) {
  if (k < 1 || k > 31) return -1;
  memset(t, 0, sizeof *t);
  const uint64_t mask = (((uint64_t)1) << (2 * k)) - 1;
  size_t cap = 1024, n = 0;
  uint64_t *all = (uint64_t *)malloc(cap * sizeof(uint64_t));
  for (int s = 0; s < nseq; ++s) {
    if (lens[s] < k) continue;  /* :478,595 */
    const char *seq = seqs[s];
    int64_t len = c_strlen_bounded(seq, lens[s]);
    int64_t a, b, from = 0;
    while (next_run(seq, len, from, &a, &b)) {
      from = b;
      int64_t L = b - a;
      if (L < k) continue;
      if (L == k && b == len) continue; /* :143-144 */
      uint64_t code = code_ending_at(seq, a + k - 1, k) & mask;
      for (int64_t e = a + k - 1; e < b; ++e) {
        if (e > a + k - 1) code = ((code << 2) | base2((unsigned char)seq[e])) & mask;
        if (n == cap) { cap *= 2; all = (uint64_t *)realloc(all, cap * sizeof(uint64_t)); }
        all[n++] = code;
      }
    }
  }
  t->total = (double)n;
  qsort(all, n, sizeof(uint64_t), cmp_u64);
  size_t nd = 0;
  for (size_t i = 0; i < n; ++i)
    if (i == 0 || all[i] != all[i - 1]) ++nd;
  t->codes = (uint64_t *)malloc((nd ? nd : 1) * sizeof(uint64_t));
  t->counts = (uint32_t *)malloc((nd ? nd : 1) * sizeof(uint32_t));
  t->ranks = (double *)calloc(nd ? nd : 1, sizeof(double));
  size_t j = 0;
  for (size_t i = 0; i < n; ++i) {
    if (i == 0 || all[i] != all[i - 1]) { t->codes[j] = all[i]; t->counts[j] = 0; ++j; }
    t->counts[j - 1]++;
  }
  t->nd = nd;
  free(all);
  return 0;
}

typedef struct { uint32_t count; uint32_t idx; } cnt_idx;
static int cmp_cnt_idx(const void *a, const void *b) {
  const cnt_idx *x = (const cnt_idx *)a, *y = (const cnt_idx *)b;
  if (x->count != y->count) return x->count < y->count ? -1 : 1;
  return x->idx < y->idx ? -1 : x->idx > y->idx;  /* idx ascending == code ascending */
}

/* rank by extension: order (count, code) over the k-mers that occur, sequential accumulation in double */
void kso_large_rank(kso_large *t) {
  cnt_idx *o = (cnt_idx *)malloc((t->nd ? t->nd : 1) * sizeof(cnt_idx));
  for (size_t i = 0; i < t->nd; ++i) { o[i].count = t->counts[i]; o[i].idx = (uint32_t)i; }
  qsort(o, t->nd, sizeof(cnt_idx), cmp_cnt_idx);
  double r = 0.0;
  for (size_t i = 0; i < t->nd; ++i) {
    if (i > 0) r = r + ((double)o[i - 1].count / t->total);
    t->ranks[o[i].idx] = r;
  }
  free(o);
}

static size_t find_code(const kso_large *t, uint64_t code) {
  size_t lo = 0, hi = t->nd;
  while (hi - lo > 1) {
    size_t mid = (lo + hi) / 2;
    if (t->codes[mid] <= code) lo = mid; else hi = mid;
  }
  return lo;  /* every scored k-mer was counted, so it is there */
}

/* kmer_regions (:243-307) with the weight of a k-mer looked up in the sparse table: mode 0 rank - thr,
 * mode 2 (count / T >= f_t ? +1 : -1) - thr */
static void regions_seq_large(const char *seq, int64_t len, int seq_id, int k, const kso_large *t, int mode,
                              double param, double thr, uint64_t min_width, double min_score, kso_spans *out) {
  len = c_strlen_bounded(seq, len);
  const uint64_t mask = (((uint64_t)1) << (2 * k)) - 1;
  int64_t a, b, from = 0;
  while (next_run(seq, len, from, &a, &b)) {
    from = b;
    if (b - a < (int64_t)k + 1) continue;
    int64_t i = a + k;
    uint64_t code = code_ending_at(seq, i - 1, k) & mask;
    double S = 0.0, M = 0.0;
    int64_t beg = 0, pk = 0;
    for (;;) {
      if (i >= b) {
        if (S > 0 && (uint64_t)(pk - beg) >= min_width && M >= min_score) {
          spans_push(out, seq_id, (int32_t)beg, (int32_t)pk, M);
          i = pk + 1;
          code = code_ending_at(seq, pk, k) & mask;
          S = 0.0; M = 0.0;
          continue;
        }
        break;
      }
      size_t at = find_code(t, code);
      double W = mode == 0 ? t->ranks[at] : (((double)t->counts[at] / t->total) >= param ? 1.0 : -1.0);
      double w = W - thr;
      double Sn = S + w;
      Sn = Sn > 0 ? Sn : 0;
      if (S == 0 && Sn > 0) { beg = i; pk = i; M = Sn; }
      if (Sn == 0 && S > 0) {
        if ((uint64_t)(pk - beg) >= min_width && M >= min_score) {
          spans_push(out, seq_id, (int32_t)beg, (int32_t)pk, M);
          i = pk + 1;
          code = code_ending_at(seq, pk, k) & mask;
          S = 0.0; M = 0.0;
          continue;
        }
        M = 0.0; pk = i;
      }
      if (Sn > M) { M = Sn; pk = i; }
      S = Sn;
      code = ((code << 2) | base2((unsigned char)seq[i])) & mask;
      ++i;
    }
  }
}

/* count -> score (mode 0: rank by extension; mode 2: +-1 around the frequency `param`) -> scan */
int kso_large_regions(const char *const *seqs, const int64_t *lens, int nseq, int k, int mode, double param,
                      double thr, int min_width, double min_score, kso_large *t, kso_spans *out) {
  if (mode != 0 && mode != 2) return -3;
  int rc = kso_large_count(seqs, lens, nseq, k, t);
  if (rc) return rc;
  kso_large_rank(t);
  for (int i = 0; i < nseq; ++i) {
    if (lens[i] < k) continue;
    regions_seq_large(seqs[i], lens[i], i, k, t, mode, param, thr, (uint64_t)(int64_t)min_width, min_score, out);
  }
  return 0;
}

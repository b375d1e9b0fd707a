/* ks_oracle.c -- CPU restatement of the kmer_spans hot path (count -> score -> scan -> spans).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (kmer_spans_b200/, include/, r/)
 * may link, import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and only as the checker.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
 *   - the reference's own known answers KA1-KA3 (test.R:365-375, test.R:66-77, kmer_spans.R:81-83)
 *   - the UNMODIFIED reference C file compiled where it lies into oracle/_ref/ (see Makefile),
 *     on seeded random inputs (counts, ranks bit-exact, spans bit-exact incl. doubles).
 * The log2 / +-1 score modes exist in the reference only as README formulas
 * (README.md:27-42); for those two modes parity is UNPINNED and defined by kso_scores() here.
 *
 * The code is a restatement from the behavioural spec (runs, scored indices, restart at peak),
 * not a transcription: it decomposes every sequence into N-free runs first.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define KSO_MAX_K 15

typedef struct {
  int32_t *pos;   /* 3 x n, column-major: seq_id, start, end (reference seq_regions int_data) */
  double *score;  /* 2 x n, column-major: score, 0           (reference seq_regions double_data) */
  size_t n, cap;
} kso_spans;

void kso_spans_init(kso_spans *s) { memset(s, 0, sizeof *s); }
void kso_spans_free(kso_spans *s) { free(s->pos); free(s->score); memset(s, 0, sizeof *s); }

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
  s->score[2 * s->n + 1] = 0.0; /* "entropy" column is always 0: src/kmer_spans.c:280,302 */
  s->n++;
}

/* 2-bit code of one byte: A0 C1 T2 G3 by construction, any byte accepted (src/kmer_spans.c:34) */
static inline uint32_t base2(unsigned char c) { return (c >> 1) & 3u; }
/* only N / n break a run (src/kmer_spans.c:35,112,123) */
static inline int is_n(unsigned char c) { return (c | 0x20) == 'n'; }

/* code of the k-mer whose LAST base is at index e (first base most significant) */
static uint32_t code_ending_at(const char *seq, int64_t e, int k) {
  uint32_t c = 0;
  for (int64_t j = e - k + 1; j <= e; ++j) c = (c << 2) | base2((unsigned char)seq[j]);
  return c; /* k <= 15 -> 30 bits */
}

/* next run [a,b) at or after `from`; returns 0 when none.  len = bytes before the terminator. */
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

/* SURVEY A.2 / src/kmer_spans.c:135-155.  Returns the number of words counted. */
uint64_t kso_count_seq(const char *seq, int64_t len, int k, int32_t *counts) {
  len = c_strlen_bounded(seq, len);
  const uint32_t mask = (k == 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
  uint64_t words = 0;
  int64_t a, b, from = 0;
  while (next_run(seq, len, from, &a, &b)) {
    from = b;
    int64_t L = b - a;
    if (L < k) continue;
    /* a run of exactly k bases that ends at the terminator is dropped (src/kmer_spans.c:143-144) */
    if (L == k && b == len) continue;
    uint32_t code = code_ending_at(seq, a + k - 1, k) & mask;
    counts[code]++;
    ++words;
    for (int64_t e = a + k; e < b; ++e) {
      code = ((code << 2) | base2((unsigned char)seq[e])) & mask;
      counts[code]++;
      ++words;
    }
  }
  return words;
}

/* stable LSD radix sort of indices 0..n-1 by count ascending: order (count, index), which is
 * what glibc 2.39 qsort_r (mergesort) yields for src/kmer_spans.c:177-197. */
static uint32_t *stable_order_by_count(const int32_t *counts, size_t n) {
  uint32_t *idx = (uint32_t *)malloc(n * sizeof(uint32_t));
  uint32_t *tmp = (uint32_t *)malloc(n * sizeof(uint32_t));
  for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
  uint32_t maxc = 0;
  for (size_t i = 0; i < n; ++i) if ((uint32_t)counts[i] > maxc) maxc = (uint32_t)counts[i];
  for (int shift = 0; shift < 32 && (maxc >> shift) != 0; shift += 8) {
    size_t hist[257];
    memset(hist, 0, sizeof hist);
    for (size_t i = 0; i < n; ++i) hist[(((uint32_t)counts[idx[i]]) >> shift & 255u) + 1]++;
    for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
    for (size_t i = 0; i < n; ++i) tmp[hist[((uint32_t)counts[idx[i]]) >> shift & 255u]++] = idx[i];
    uint32_t *t = idx; idx = tmp; tmp = t;
  }
  free(tmp);
  return idx;
}

/* SURVEY A.3 / src/kmer_spans.c:189-202, with the zero-filled-buffer reading of T5:
 * rank[pi_0] = 0; rank[pi_i] = rank[pi_{i-1}] + count[pi_{i-1}] / T, sequentially in double. */
void kso_rank(const int32_t *counts, int k, double total, double *ranks) {
  size_t n = (size_t)1 << (2 * k);
  uint32_t *order = stable_order_by_count(counts, n);
  double r = 0.0;
  ranks[order[0]] = 0.0;
  for (size_t i = 1; i < n; ++i) {
    r = r + ((double)counts[order[i - 1]] / total);
    ranks[order[i]] = r;
  }
  free(order);
}

/* the permutation itself, for order-parity tests */
void kso_rank_order(const int32_t *counts, int k, uint32_t *order_out) {
  size_t n = (size_t)1 << (2 * k);
  uint32_t *order = stable_order_by_count(counts, n);
  memcpy(order_out, order, n * sizeof(uint32_t));
  free(order);
}

/* Score modes.  KSO_RANK is the only one coded in the reference (src/kmer_spans.c:268);
 * LOG2 and SIGN follow README.md:27-42 with the conventions fixed in DESIGN.md:
 *   f_i = count_i / T; f_med = R median() over all 4^k entries (mean of the two middle
 *   order statistics); LOG2: log2(f_i / f_med); SIGN: f_i >= f_t ? +1 : -1 with f_t = f_med
 *   unless `param` is finite, in which case f_t = param.
 *   RANK_REL is the README variant (r_i - r_t) / r_t with r_t = param (src/kmer_spans.c:268 comment). */
enum { KSO_RANK = 0, KSO_LOG2 = 1, KSO_SIGN = 2, KSO_RANK_REL = 3 };

int kso_scores(const int32_t *counts, int k, double total, int mode, double param, double *W) {
  size_t n = (size_t)1 << (2 * k);
  if (mode == KSO_RANK) { kso_rank(counts, k, total, W); return 0; }
  if (mode == KSO_RANK_REL) {
    kso_rank(counts, k, total, W);
    for (size_t i = 0; i < n; ++i) W[i] = (W[i] - param) / param;
    return 0;
  }
  uint32_t *order = stable_order_by_count(counts, n);
  double f_lo = (double)counts[order[n / 2 - 1]] / total;
  double f_hi = (double)counts[order[n / 2]] / total;
  double f_med = (f_lo + f_hi) / 2.0;
  free(order);
  if (mode == KSO_LOG2) {
    for (size_t i = 0; i < n; ++i) W[i] = log2(((double)counts[i] / total) / f_med);
    return 0;
  }
  if (mode == KSO_SIGN) {
    double f_t = isfinite(param) ? param : f_med;
    for (size_t i = 0; i < n; ++i) W[i] = ((double)counts[i] / total) >= f_t ? 1.0 : -1.0;
    return 0;
  }
  return -1;
}

/* SURVEY A.4 / src/kmer_spans.c:243-307: clamped scan with restart at the peak. */
void kso_regions_seq(const char *seq, int64_t len, int seq_id, int k, const double *W, double thr,
                     uint64_t min_width, double min_score, kso_spans *out, int32_t *inscan_counts) {
  len = c_strlen_bounded(seq, len);
  const uint32_t mask = (1u << (2 * k)) - 1u;
  int64_t a, b, from = 0;
  while (next_run(seq, len, from, &a, &b)) {
    from = b;
    if (b - a < (int64_t)k + 1) continue; /* scored indices are a+k .. b-1 */
    int64_t i = a + k;
    uint32_t code = code_ending_at(seq, i - 1, k) & mask;
    double S = 0.0, M = 0.0;
    int64_t beg = 0, pk = 0;
    for (;;) {
      if (i >= b) {
        /* run end: an open excursion is tested, and a hit restarts behind its peak (:298-305) */
        if (S > 0 && (uint64_t)(pk - beg) >= min_width && M >= min_score) {
          spans_push(out, seq_id, (int32_t)beg, (int32_t)pk, M);
          i = pk + 1;
          code = code_ending_at(seq, pk, k) & mask;
          S = 0.0; M = 0.0;
          continue;
        }
        break;
      }
      if (inscan_counts) inscan_counts[code]++;
      double w = W[code] - thr;
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

/* ---- entry-point equivalents (plain C mirrors of the three .Call functions) ---- */

/* kmer_counts, src/kmer_spans.c:453-487 */
int kso_kmer_counts(const char *const *seqs, const int64_t *lens, int nseq, int k, int32_t *counts,
                    double *n_words) {
  if (k < 1 || k > KSO_MAX_K) return -1;
  memset(counts, 0, sizeof(int32_t) << (2 * k));
  *n_words = 0;
  for (int i = 0; i < nseq; ++i) {
    if (lens[i] < k) continue;
    *n_words += (double)kso_count_seq(seqs[i], lens[i], k, counts);
  }
  return 0;
}

/* kmer_regions_r, src/kmer_spans.c:490-546 (threshold fixed to 0, in-scan counts, T8) */
int kso_kmer_regions(const char *const *seqs, const int64_t *lens, int nseq, int k, const double *W,
                     int min_width, double min_score, double *nuc, int32_t *inscan_counts,
                     kso_spans *out) {
  if (k < 1 || k > KSO_MAX_K) return -1;
  memset(inscan_counts, 0, sizeof(int32_t) << (2 * k));
  *nuc = 0;
  for (int i = 0; i < nseq; ++i) {
    if (lens[i] < k) continue;
    *nuc += (double)lens[i];
    kso_regions_seq(seqs[i], lens[i], i, k, W, 0.0, (uint64_t)(int64_t)min_width, min_score, out,
                    inscan_counts);
  }
  return 0;
}

/* kmer_low_comp_regions, src/kmer_spans.c:548-621 */
int kso_low_comp(const char *const *seqs, const int64_t *lens, int nseq, int k, int min_width,
                 double min_score, double thr, double *n_out /*2*/, int32_t *counts, double *ranks,
                 kso_spans *out) {
  if (k < 1 || k > KSO_MAX_K) return -1;
  if (!(thr > 0 && thr < 1)) return -2;
  kso_kmer_counts(seqs, lens, nseq, k, counts, &n_out[0]);
  n_out[1] = 0;
  kso_rank(counts, k, n_out[0], ranks);
  for (int i = 0; i < nseq; ++i) {
    if (lens[i] < k) continue;
    kso_regions_seq(seqs[i], lens[i], i, k, ranks, thr, (uint64_t)(int64_t)min_width, min_score,
                    out, NULL);
  }
  return 0;
}

/* fused "mode" pipeline used as checker for the extension entry point ks_mode_regions():
 * counts (clean pass) -> W = scores(mode) -> scan with threshold thr. */
int kso_mode_regions(const char *const *seqs, const int64_t *lens, int nseq, int k, int mode,
                     double param, double thr, int min_width, double min_score, double *n_words,
                     int32_t *counts, double *W, kso_spans *out) {
  if (k < 1 || k > KSO_MAX_K) return -1;
  kso_kmer_counts(seqs, lens, nseq, k, counts, n_words);
  if (kso_scores(counts, k, *n_words, mode, param, W)) return -3;
  for (int i = 0; i < nseq; ++i) {
    if (lens[i] < k) continue;
    kso_regions_seq(seqs[i], lens[i], i, k, W, thr, (uint64_t)(int64_t)min_width, min_score, out,
                    NULL);
  }
  return 0;
}

/* kmer_seq, src/kmer_spans.c:161-171: index -> string, alphabet order A,C,T,G */
int kso_kmer_seq(int k, uint64_t code, char *out) {
  static const char nuc[4] = {'A', 'C', 'T', 'G'};
  if (k < 1 || k > 16) return -1;
  out[k] = 0;
  for (int j = k - 1; j >= 0; --j) { out[j] = nuc[code & 3]; code >>= 2; }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * "next" rows of SURVEY 8(f): the two remaining .Call entries of the reference.
 * Both are checked against the compiled reference in tests/test_oracle.py. */

/* What init_kmer (src/kmer_spans.c:119-132) leaves in *offset for a k-mer STRING, as the callers at
 * :691 and :747 use it: the first min(k, len) bases of the last N-free piece it looks at; it stops
 * looking at the first piece that holds k bases.  For a clean string of k bases this is its code. */
uint32_t kso_kmer_code(const char *s, int k) {
  uint32_t code = 0;
  size_t i = 0;
  while (s[i]) {
    code = 0;
    int got = 0;
    while (got < k && s[i] && !is_n((unsigned char)s[i])) { code = (code << 2) | base2((unsigned char)s[i]); ++i; ++got; }
    if (got == k || !s[i]) break;
    while (s[i] && is_n((unsigned char)s[i])) ++i;
  }
  return code;
}

/* windowed_kmer_count_distributions(_r), src/kmer_spans.c:398-449,715-793.
 * A window is `window` consecutive bases inside one run; it starts at every s in [a, b - window] of a
 * run [a,b).  Its value for a selected k-mer x is the number of occurrences of x lying entirely
 * inside it, i.e. k-mers ending at e in [s+k-1, s+window-1] (:425-444).  dist[i*(window+1) + c]
 * counts the windows with value c for selected k-mer i; pos[seq][i*len + s] (optional) holds the
 * value of the window starting at s (:440-441).  Sequences with len <= window are left out (:775). */
int kso_window_dist(const char *const *seqs, const int64_t *lens, int nseq, int k, const uint32_t *codes,
                    int kmer_n, int window, int32_t *dist, int32_t *included, int32_t *const *pos) {
  if (k < 1 || k > KSO_MAX_K || window < 2 * k || kmer_n < 1) return -1;
  const uint32_t mask = (1u << (2 * k)) - 1u;
  memset(dist, 0, sizeof(int32_t) * (size_t)(window + 1) * (size_t)kmer_n);
  for (int q = 0; q < nseq; ++q) {
    included[q] = lens[q] > window;
    if (!included[q]) continue;
    const char *seq = seqs[q];
    int64_t len = c_strlen_bounded(seq, lens[q]);
    if (pos && pos[q]) memset(pos[q], 0, sizeof(int32_t) * (size_t)lens[q] * (size_t)kmer_n);
    int64_t a, b, from = 0;
    while (next_run(seq, len, from, &a, &b)) {
      from = b;
      if (b - a < window) continue;
      uint8_t *hit = (uint8_t *)malloc((size_t)(b - a));
      for (int i = 0; i < kmer_n; ++i) {
        /* hit[e-a] = the k-mer ending at e is the selected one */
        memset(hit, 0, (size_t)(b - a));
        uint32_t code = code_ending_at(seq, a + k - 1, k) & mask;
        hit[k - 1] = code == codes[i];
        for (int64_t e = a + k; e < b; ++e) {
          code = ((code << 2) | base2((unsigned char)seq[e])) & mask;
          hit[e - a] = code == codes[i];
        }
        int32_t c = 0;
        for (int64_t e = a + k - 1; e < a + window - 1; ++e) c += hit[e - a];
        for (int64_t s = a; s + window <= b; ++s) {
          c += hit[s + window - 1 - a];
          dist[(size_t)i * (size_t)(window + 1) + (size_t)c]++;
          if (pos && pos[q]) pos[q][(size_t)i * (size_t)lens[q] + (size_t)s] = c;
          c -= hit[s + k - 1 - a];
        }
      }
      free(hit);
    }
  }
  return 0;
}

/* find_kmer_tr_lr_regions / tr_lr_regions_r, src/kmer_spans.c:329-395,649-713.  `init` and `trans`
 * are in 2-bit code order (the .Call wrapper's reordering by k-mer strings, :688-696, is the
 * caller's job: see kso_kmer_code).  A run [a,b) starts with S = max(init[first k-mer], 0) "at"
 * index a+k, then S_i = max(S_{i-1} + trans[code(i)], 0) for i = a+k .. b-1 (the k-mer ENDING at
 * i, :361-362).  Every excursion that returns to 0 is tested on width only (:377) and the scan ALWAYS
 * resumes behind its peak (:382-388); an excursion still open at the run end is tested but not
 * re-scanned (:392-393).  Coordinates and seq_id are 1-based (:379,681).  A run is not looked at when
 * the string ends within one base of its first k-mer (:340-341). */
void kso_tr_lr_seq(const char *seq, int64_t len, int seq_id1, int k, const double *init, const double *trans,
                   int min_len, kso_spans *out) {
  len = c_strlen_bounded(seq, len);
  const uint32_t mask = (1u << (2 * k)) - 1u;
  int64_t a, b, from = 0;
  while (next_run(seq, len, from, &a, &b)) {
    from = b;
    if (b - a < k) continue;
    int64_t i = a + k;
    if (i >= len || i + 1 >= len) break; /* terminator at i or i+1 ends the whole sequence (:340) */
    uint32_t code = code_ending_at(seq, i - 1, k) & mask;
    double S = init[code] < 0 ? 0 : init[code];
    double M = 0;
    int64_t pk = 0, beg = 0;
    if (S > 0) { M = S; pk = i; beg = i; }
    while (i < b) {
      code = ((code << 2) | base2((unsigned char)seq[i])) & mask;
      double Sn = S + trans[code];
      if (Sn > M) { M = Sn; pk = i; }
      Sn = Sn < 0 ? 0 : Sn;
      if (S == 0 && Sn > 0) { M = Sn; pk = i; beg = i; }
      if (Sn == 0 && S > 0) {
        if ((int)pk - (int)beg >= min_len) spans_push(out, seq_id1, (int32_t)(1 + beg), (int32_t)(1 + pk), M);
        i = pk; /* resume behind the peak with a clean state */
        beg = pk;
        pk = 0;
        M = 0;
        Sn = 0;
        code = code_ending_at(seq, i, k) & mask;
      }
      S = Sn;
      ++i;
    }
    if (M > 0 && (int)pk - (int)beg >= min_len)
      spans_push(out, seq_id1, (int32_t)(1 + beg), (int32_t)(1 + pk), M);
  }
}

int kso_tr_lr_regions(const char *const *seqs, const int64_t *lens, int nseq, int k, const double *init,
                      const double *trans, int min_len, kso_spans *out) {
  if (k < 1 || k > KSO_MAX_K || min_len < 0) return -1;
  for (int i = 0; i < nseq; ++i) kso_tr_lr_seq(seqs[i], lens[i], i + 1, k, init, trans, min_len, out);
  return 0;
}

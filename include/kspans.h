/* kspans.h -- C ABI of the B200-native kmer_spans hot path (count -> score -> scan -> spans).
 *
 * Plain C: pointers and sizes only, no R, torch or C++ types.  The shared library that exports
 * these symbols (kmer_spans_b200/csrc/libkspans_cuda.so) contains hand-written sm_100a kernels
 * and NO CPU fallback: every compute entry point returns KS_ERR_CUDA when no device is usable.
 *
 * Each entry point names the reference interface it replaces
 * (file:line in lmjakt/kmer_spans, src/kmer_spans.c unless stated otherwise).
 * INTEGRATION.md shows the .Call glue (r/src/kmer_spans_glue.c) that binds them into R.
 *
 * Conventions shared by all calls
 *  - sequences are (pointer, length) pairs; a NUL byte inside [ptr, ptr+len) is treated as the
 *    string terminator the reference stops at (:121,140,261) for that run (R strings never hold one)
 *  - only 'N'/'n' break a run; every other byte maps through (c>>1)&3: A0 C1 T2 G3   (:34-35)
 *  - sequences shorter than k are skipped but keep their 0-based seq_id             (:478,533,595)
 *  - k in 1..15; count tables are int32[4^k], score tables double[4^k], index = 2-bit code
 *  - span coordinates follow the reference exactly: int32, 0-based seq_id,
 *    start/end = 1-based position of the last base of the first-positive / peak k-mer
 */
#ifndef KSPANS_H
#define KSPANS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ks_ctx ks_ctx;       /* one device, one stream, cached scratch memory; use from one host thread at a time */
typedef struct ks_seqset ks_seqset; /* sequences resident in HBM   */

enum {
  KS_OK = 0,
  KS_ERR_ARG = 1,      /* argument rejected; message mirrors the reference's error() text where one exists */
  KS_ERR_CUDA = 2,     /* no device / CUDA runtime failure: there is no CPU path */
  KS_ERR_RANGE = 3,    /* a weight is +inf or >= 2^40, or nonzero but more than 2^57 times smaller than the largest
                        * (exact fixed-point scan range) */
  KS_ERR_NOMEM = 4
};

/* score modes of README.md:27-49.  KS_MODE_RANK is the only one the reference codes (:268). */
enum { KS_MODE_RANK = 0, KS_MODE_LOG2 = 1, KS_MODE_SIGN = 2, KS_MODE_RANK_REL = 3 };

/* Spans in the reference's own layout (struct seq_regions :46-58, as copied out at :541-542):
 * pos   = int32[3*n], column-major 3 x n : seq_id, start, end
 * score = double[2*n], column-major 2 x n : peak score, 0 ("entropy", always 0 :280)
 * Order = reference discovery order = ascending (seq_id, start).  Library-owned. */
typedef struct {
  int32_t *pos;
  double *score;
  size_t n;
} ks_spans;
void ks_spans_free(ks_spans *s);

/* -------- context ------------------------------------------------------------------------- */
/* CUDA is initialised lazily here, never at library load (R may fork). device < 0
 * = current device. */
int ks_ctx_create(ks_ctx **out, int device);
void ks_ctx_destroy(ks_ctx *ctx);
const char *ks_last_error(const ks_ctx *ctx); /* ctx may be NULL: error of the last failed ks_ctx_create */
void *ks_ctx_stream(ks_ctx *ctx);             /* cudaStream_t all kernels of this ctx are launched on */
int ks_ctx_sync(ks_ctx *ctx);
/* on: ks_dev_scores (log2 / +-1 modes) writes the 4^k score TABLE -- an output the scan does not read -- on the
 * context's copy stream, next to the scan that follows; the table is complete when the next ks_dev_scan_counts*
 * call or ks_ctx_sync returns.  Off (default): complete in stream order.  ks_dev_pipeline always does this and
 * joins before it returns. */
void ks_ctx_side_table(ks_ctx *ctx, int on);
/* kernels launched through this ctx since creation (or since the last reset) */
uint64_t ks_ctx_launches(const ks_ctx *ctx);
void ks_ctx_reset_launches(ks_ctx *ctx);

/* -------- host-buffer entry points: one per reference .Call on the path -------------------- */

/* kmer_counts (:453-487): counts_out[4^k] (overwritten), *n_words = number of words counted. */
int ks_kmer_counts(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                   int32_t *counts_out, double *n_words);

/* kmer_regions_r (:490-546): user weights W[4^k], threshold 0, in-scan counts that miss the last
 * k-mer of every run and double-count rescanned positions.
 * *nuc = sum of the lengths of the sequences with length >= k (:535). min_width is the R integer;
 * negative values wrap to huge exactly as the reference's size_t conversion (:243,279). */
int ks_kmer_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                    const double *W, int min_width, double min_score, double *nuc,
                    int32_t *inscan_counts_out, ks_spans *out);

/* kmer_low_comp_regions (:548-621): counts -> weighted ranks (:189-202, stable (count,index) order,
 * rank of the first k-mer in that order = 0) -> scan with score = rank - thr, 0 < thr < 1.
 * n_out[0] = words counted, n_out[1] = 0 (:613).  counts_out / ranks_out may be NULL (not copied back). */
int ks_kmer_low_comp_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq,
                             int k, int min_width, double min_score, double thr, double n_out[2],
                             int32_t *counts_out, double *ranks_out, ks_spans *out);

/* Extension (no reference entry point; the reference reaches these modes by computing a weight
 * vector in R and calling kmer_regions_r, kmer_spans.R:41-52): counts -> scores(mode) -> scan with
 * threshold thr, everything device-resident.  scores_out may be NULL. */
int ks_kmer_mode_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                         int mode, double param, double thr, int min_width, double min_score,
                         double *n_words, int32_t *counts_out, double *scores_out, ks_spans *out);

/* rank_kmers_w (:189-202) and the README modes as a table-to-table operator on host buffers. */
int ks_kmer_scores(ks_ctx *ctx, int k, const int32_t *counts, double total, int mode, double param,
                   double *scores_out);

/* kmer_seq (:161-171): index -> k-mer string (A,C,T,G order); host only, no device needed. */
int ks_kmer_seq(int k, uint64_t code, char *out /* k+1 bytes */);
/* What init_kmer (:119-132) leaves in its offset for a k-mer STRING, as the callers at :691 and :747
 * use it: the 2-bit code of a clean string of k bases (first base most significant).  Host only. */
uint32_t ks_kmer_code(const char *kmer, int k);

/* tr_lr_regions_r (:649-713, core find_kmer_tr_lr_regions :329-395), SURVEY 8(f) row 3.
 * init_scores / trans_scores: double[4^k] in 2-bit code order (the reordering by k-mer strings of
 * :688-696 is the caller's job: table[ks_kmer_code(kmers[i], k)] = value[i]).  Every run starts with
 * S = max(init[first k-mer], 0), then S = max(S + trans[k-mer ending at i], 0); every excursion that
 * returns to 0 is tested on peak - start >= min_length and the scan resumes behind its peak; an
 * excursion open at the run end is tested, not re-scanned.  Spans: seq_id, start, end all 1-based,
 * score = peak, second score column 0.  NaN scores are rejected (KS_ERR_RANGE), -Inf is allowed. */
int ks_tr_lr_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                     const double *init_scores, const double *trans_scores, int min_length, ks_spans *out);

/* windowed_kmer_count_distributions_r (:715-793, core :398-449), SURVEY 8(f) row 4.
 * codes[kmer_n] = 2-bit codes of the selected k-mers (ks_kmer_code), 1 <= k <= 15, window >= 2k.
 * dist_out[i * (window+1) + c] = number of windows (window consecutive bases inside one run of one
 * sequence longer than `window`) that hold c occurrences of selected k-mer i: the R matrix
 * (window+1) x kmer_n, column-major.  included_out[nseq] = 1 for sequences longer than window (:775).
 * pos_out: NULL, or nseq pointers; pos_out[q] (NULL for sequences left out) receives the R matrix
 * lens[q] x kmer_n: the value of the window STARTING at each position, 0 where none starts (:440-441). */
int ks_windowed_kmer_count_distributions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq,
                                         int k, const uint32_t *codes, int kmer_n, int window,
                                         int32_t *dist_out, int32_t *included_out, int32_t *const *pos_out);

/* -------- device-resident sequence sets (upload once, scan many times) --------------------- */
int ks_seqset_upload(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, ks_seqset **out);
/* Upload new sequences into an existing (library-owned) set, re-using its device buffers.  count_k > 0:
 * the pack+count pass runs behind the copies, d_counts[4^k] is overwritten and the word count is left at
 * d_nwords (device uint64, may be NULL).  Asynchronous: pinned host buffers must stay valid until the
 * next synchronising call on this ctx. */
int ks_seqset_reupload(ks_ctx *ctx, ks_seqset *s, const char *const *seqs, const int64_t *lens, int nseq,
                       int count_k, int32_t *d_counts, uint64_t *d_nwords);
/* wrap a buffer ALREADY in device memory, laid out as csrc/ks_layout.h describes
 * (d_buf must stay valid; starts = nseq+1 host offsets as returned by ks_layout) */
int ks_seqset_wrap(ks_ctx *ctx, const void *d_buf, int64_t total_bytes, const int64_t *lens, int nseq,
                   ks_seqset **out);
void ks_seqset_free(ks_seqset *s);
int64_t ks_seqset_bases(const ks_seqset *s);      /* sum of lengths */
int64_t ks_seqset_buffer_bytes(const ks_seqset *s);

/* Stage-level device API (what bench.py times, and what the multi-GPU host composes around its
 * collectives).  d_* arguments are DEVICE pointers owned by the caller. */
int ks_dev_count(ks_ctx *ctx, const ks_seqset *s, int k, int32_t *d_counts /*4^k, zeroed here*/,
                 double *n_words);
int ks_dev_scores(ks_ctx *ctx, int k, const int32_t *d_counts, double total, int mode, double param,
                  double *d_scores /*4^k*/);
int ks_dev_scan(ks_ctx *ctx, const ks_seqset *s, int k, const double *d_W, double thr, int min_width,
                double min_score, int32_t *d_inscan_or_null, ks_spans *host_out_or_null,
                uint64_t *n_spans);
/* The same two stages without a host round trip, for compositions that sum the count table and the word
 * count with a collective on the ctx stream (kmer_spans_b200/dist.py): ks_dev_count_async leaves the word
 * count in device memory (*d_nwords, uint64), ks_dev_scores_devtotal takes the total from device memory
 * and reads it back together with the first table the score stage needs (*total_out, may be NULL). */
int ks_dev_count_async(ks_ctx *ctx, const ks_seqset *s, int k, int32_t *d_counts, uint64_t *d_nwords);
int ks_dev_scores_devtotal(ks_ctx *ctx, int k, const int32_t *d_counts, const uint64_t *d_total, int mode,
                           double param, double *d_scores, double *total_out);
/* Sum of the count tables of all ranks over NVLink / NVSwitch peer memory in one kernel (csrc/ks_xgpu.cuh):
 * tables[nranks] = every rank's buffer of n_u64 64-bit words (pairs of int32 counters, then the u64 word
 * count) as mapped into this process (symmetric memory), mc_table = its NVSwitch multicast address or NULL
 * (then plain peer loads / stores are used).  Rank `rank` reduces its slice and writes the sums into every
 * rank's buffer.  The caller places a cross-GPU barrier on the ctx stream before and after the call. */
int ks_dev_xsum(ks_ctx *ctx, void *const *tables, int nranks, int rank, void *mc_table, uint64_t n_u64);
/* Scan with score = f(count) for the count-derived modes: f is the function the last
 * ks_dev_scores(mode LOG2 | SIGN) on this ctx derived (d_scores may be NULL there).  The kernel
 * gathers the 4-byte count (table L2 resident up to k = 12) and maps it through a dense LUT. */
int ks_dev_scan_counts(ks_ctx *ctx, const ks_seqset *s, int k, const int32_t *d_counts, double thr,
                       int min_width, double min_score, ks_spans *host_out_or_null, uint64_t *n_spans);
/* Scan in rank mode, score = rank - thr (the one mode the reference codes, src/kmer_spans.c:268,602-612), with
 * the rank order the last ks_dev_scores(mode KS_MODE_RANK) on this ctx derived.  The kernel gathers the 4-byte
 * position of the k-mer in the stable (count, index) order (4^k x 4 B, L2 resident up to k = 12) and evaluates
 * the rank from the linear pieces of the closed form; bit-identical to ks_dev_scan on the rank table. */
int ks_dev_scan_ranks(ks_ctx *ctx, const ks_seqset *s, int k, double thr, int min_width, double min_score,
                      ks_spans *host_out_or_null, uint64_t *n_spans);
/* transition-score scan on a resident set; d_init / d_trans = device double[4^k] in code order */
int ks_dev_tr_lr_regions(ks_ctx *ctx, const ks_seqset *s, int k, const double *d_init, const double *d_trans,
                         int min_length, ks_spans *host_out_or_null, uint64_t *n_spans);
/* windowed occurrence histograms on a resident set: d_dist = device int32[kmer_n * (window+1)]
 * (overwritten); d_pos = NULL or device int32[kmer_n * ks_seqset_positions(s)] (overwritten), the value
 * of the window starting at every buffer position (position of base i of sequence q:
 * ks_seqset_start(s, q) + i).  codes are host memory. */
int ks_dev_window_dist(ks_ctx *ctx, const ks_seqset *s, int k, const uint32_t *codes, int kmer_n, int window,
                       int32_t *d_dist, int32_t *d_pos);
int64_t ks_seqset_positions(const ks_seqset *s);
int64_t ks_seqset_start(const ks_seqset *s, int seq);
/* count -> scores(mode) -> scan, all resident; spans stay on the device unless host_out != NULL */
int ks_dev_pipeline(ks_ctx *ctx, const ks_seqset *s, int k, int mode, double param, double thr,
                    int min_width, double min_score, int32_t *d_counts, double *d_scores,
                    double *n_words, ks_spans *host_out_or_null, uint64_t *n_spans);

/* -------- one sequence set sharded across GPUs (exact stitching of the scan) -----------------------
 * Every rank uploads (and packs) the whole set, counts and scans only its dense chunk range
 * [chunk0, chunk0 + nchunks) of the buffer (16 positions per chunk, ks_seqset_chunks() in total) and
 * exchanges two 48-byte carries: the shard's aggregate max-plus transform and its open-excursion
 * aggregate.  `fn` is called on the host inside ks_dev_scan*_shard: what = 0 / 1, `mine` = this shard's
 * 48-byte aggregate; it must fill `carry_in` with what ks_fold_carry() derives from the aggregates of
 * ALL shards (e.g. after an all-gather).  Spans come back with global coordinates; a span is reported
 * by the shard in which it closes. */
typedef int (*ks_exchange_fn)(void *user, int what, const void *mine48, void *carry_in48);
/* Rank-mode score stage sliced over the devices of a multi-GPU run: the summed count table is identical on every
 * device, so device `slice` of `nslices` sorts and ranks only its slice [slice * 4^k / nslices, (slice + 1) * 4^k /
 * nslices) of the k-mer index space; the slices' run-length tables (count, multiplicity: a few thousand pairs) are
 * exchanged through `fn`, called on the host: it must fill `all` with the `bytes` of every slice in slice order
 * (an all-gather).  On return d_scores and the ctx's rank-order position table (ks_ctx_rank_positions, uint32[4^k])
 * hold the caller's slice; the caller gathers the other slices into both.  Replaces the redundant full sort of
 * rank_kmers_w (src/kmer_spans.c:189-202) on every device; results are bit-identical to ks_dev_scores. */
typedef int (*ks_gather_fn)(void *user, const void *mine, size_t bytes, void *all);
int ks_dev_scores_rank_sliced(ks_ctx *ctx, int k, const int32_t *d_counts, double total, int slice, int nslices,
                              ks_gather_fn fn, void *user, double *d_scores);
void *ks_ctx_rank_positions(ks_ctx *ctx);
/* Shard planner + sharded upload: the layout of ALL sequences is cut into nranks contiguous chunk ranges; shard
 * `rank` owns chunks [chunk0, chunk0 + nchunks) and keeps only bytes [win_lo, win_hi) of the layout in HBM (its
 * range plus the head of the sequence the range starts in, where the re-scans of a span closing in the range may
 * reach).  A window set is counted with ks_dev_count_range and scanned with ks_dev_scan*_shard; coordinates and
 * sequence ids stay global.  Replaces the per-sequence loops of src/kmer_spans.c:592-612 across GPUs. */
int ks_plan_shard(const int64_t *lens, int nseq, int nranks, int rank, int64_t *chunk0, int64_t *nchunks,
                  int64_t *win_lo, int64_t *win_hi);
int ks_seqset_upload_window(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int64_t win_lo,
                            int64_t win_hi, ks_seqset **out);
int64_t ks_seqset_chunks(const ks_seqset *s);
int ks_dev_count_range(ks_ctx *ctx, const ks_seqset *s, int k, int64_t chunk0, int64_t nchunks,
                       int32_t *d_counts, double *n_words);
/* the same without a host round trip: the word count is left in device memory (*d_nwords, uint64) */
int ks_dev_count_range_async(ks_ctx *ctx, const ks_seqset *s, int k, int64_t chunk0, int64_t nchunks,
                             int32_t *d_counts, uint64_t *d_nwords);
int ks_dev_scan_shard(ks_ctx *ctx, const ks_seqset *s, int k, const double *d_W, double thr, int min_width,
                      double min_score, int64_t chunk0, int64_t nchunks, ks_exchange_fn fn, void *user,
                      ks_spans *host_out_or_null, uint64_t *n_spans);
int ks_dev_scan_counts_shard(ks_ctx *ctx, const ks_seqset *s, int k, const int32_t *d_counts, double thr,
                             int min_width, double min_score, int64_t chunk0, int64_t nchunks,
                             ks_exchange_fn fn, void *user, ks_spans *host_out_or_null, uint64_t *n_spans);
int ks_dev_scan_ranks_shard(ks_ctx *ctx, const ks_seqset *s, int k, double thr, int min_width, double min_score,
                            int64_t chunk0, int64_t nchunks, ks_exchange_fn fn, void *user,
                            ks_spans *host_out_or_null, uint64_t *n_spans);
/* host helper: carry entering shard `rank` from the 48-byte aggregates of shards 0..nranks-1 */
int ks_fold_carry(int what, const void *all48, int nranks, int rank, void *carry_in48);

/* -------- large k (BASELINE.json configs[3]: k = 21) -------------------------------------------------------
 * Outside the reference's domain (src/kmer_spans.c:37 MAX_K 16, :139 int shift, :504 k >= 16 rejected), defined by
 * extension: 64-bit codes, the count table is an open-addressing hash table in HBM (it holds the
 * k-mers that occur, not 4^k entries), the score is the weighted rank over the k-mers that occur in (count, code)
 * order -- what rank_kmers_w (:189-202) would give on the full table, where absent k-mers add 0 -- or, mode
 * KS_MODE_SIGN, +-1 around the frequency `param`; the scan is kmer_regions (:243-307) unchanged.  k may be any
 * value in 1..31 (k <= 15 lets tests compare this path with the direct-table path).  One GPU. */
int ks_dev_large_regions(ks_ctx *ctx, const ks_seqset *s, int k, int mode, double param, double thr, int min_width,
                         double min_score, double *n_words, uint64_t *n_distinct, ks_spans *host_out_or_null,
                         uint64_t *n_spans);
int ks_kmer_large_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k, int mode,
                          double param, double thr, int min_width, double min_score, double *n_words,
                          uint64_t *n_distinct, ks_spans *out);
/* the sparse table the last large-k call left on this ctx, in (count, code) order: n_distinct entries per array
 * (host memory, any may be NULL) */
int ks_large_table(ks_ctx *ctx, uint64_t *codes_out, uint32_t *counts_out, double *ranks_out);

/* -------- several GPUs behind one call (one process, N devices; csrc/ks_multi.inc) -----------------
 * The reference runs its loops over all sequences of a call in one process (src/kmer_spans.c:592-601 count,
 * :604-612 scan); ks_mctx keeps that shape on N devices: the layout of all sequences is cut into N contiguous
 * chunk ranges (ks_plan_shard), device r uploads only its window, counts its range, the tables are summed by
 * one kernel per device over NVLink peer memory (ordered by CUDA events), scores are derived redundantly, every
 * device scans its range and the two 48-byte carries of all shards are folded exactly, so a span that crosses a
 * cut comes out as from one GPU.  `devices` may repeat an index (several shards on one GPU; how the single-GPU
 * test tier exercises this).  The R glue creates one from KSPANS_DEVICES=0,1,.... */
typedef struct ks_mctx ks_mctx;
int ks_mctx_create(ks_mctx **out, const int *devices, int ndev);
void ks_mctx_destroy(ks_mctx *m);
const char *ks_mctx_last_error(const ks_mctx *m); /* m may be NULL after a failed create */
int ks_mctx_ndev(const ks_mctx *m);
ks_ctx *ks_mctx_ctx(ks_mctx *m, int i); /* the per-device context (profiling, launch counts) */
/* same arguments, results and errors as ks_kmer_counts / ks_kmer_mode_regions / ks_kmer_low_comp_regions */
int ks_m_kmer_counts(ks_mctx *m, const char *const *seqs, const int64_t *lens, int nseq, int k,
                     int32_t *counts_out, double *n_words);
int ks_m_kmer_mode_regions(ks_mctx *m, const char *const *seqs, const int64_t *lens, int nseq, int k, int mode,
                           double param, double thr, int min_width, double min_score, double *n_words,
                           int32_t *counts_out_or_null, double *scores_out_or_null, ks_spans *out);
int ks_m_kmer_low_comp_regions(ks_mctx *m, const char *const *seqs, const int64_t *lens, int nseq, int k,
                               int min_width, double min_score, double thr, double n_out[2],
                               int32_t *counts_out_or_null, double *ranks_out_or_null, ks_spans *out);
/* resident shards: ks_m_load plans the cut and uploads every device's window; ks_m_pipeline runs count -> sum ->
 * scores(mode) -> sharded scan on them (count_only != 0: stop after the summed count table); ks_m_tables returns
 * the device pointers of the tables device i holds afterwards: the count table is identical on every device; so is
 * the score table, except after a weighted-rank pipeline on several devices, where device i holds slice i of the
 * rank table (the scan gathers rank-order positions, not ranks; ks_m_kmer_* collect the slices when asked). */
int ks_m_load(ks_mctx *m, const char *const *seqs, const int64_t *lens, int nseq);
int ks_m_pipeline(ks_mctx *m, int k, int mode, double param, double thr, int min_width, double min_score,
                  double *n_words, ks_spans *out_or_null, uint64_t *n_spans, int count_only);
int ks_m_tables(ks_mctx *m, int i, int32_t **d_counts, double **d_scores);

/* -------- timing on the launching stream (CUDA events; what bench.py reports) ---------------- */
int ks_ctx_timer_start(ks_ctx *ctx);
int ks_ctx_timer_stop(ks_ctx *ctx, float *ms); /* records, synchronises, returns elapsed ms */
/* per-kernel-class device time, measured live with event pairs around the launches.
 * which: 0 pack_count_kernel, 1 scan kernels of level 0, 2 scan kernels of the deeper levels,
 *        3 score-table stage (sort + rank/lut kernels, includes its host control steps), 4 wmax+wfx */
enum { KS_PROF_COUNT = 0, KS_PROF_SCAN0 = 1, KS_PROF_SCANN = 2, KS_PROF_SCORES = 3, KS_PROF_WFX = 4, KS_PROF_N = 5 };
void ks_ctx_set_profile(ks_ctx *ctx, int on);
int ks_ctx_profile_get(ks_ctx *ctx, int which, double *ms_total, uint64_t *launches);
void ks_ctx_profile_reset(ks_ctx *ctx);

/* diagnostics of the last scan on this ctx: restart levels run and positions visited beyond level 0 */
void ks_ctx_scan_stats(const ks_ctx *ctx, int *levels, uint64_t *revisited_chunks);

#ifdef __cplusplus
}
#endif
#endif

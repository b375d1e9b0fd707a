// ks_api.cu -- C ABI (include/kspans.h) and host orchestration of the sm_100a kernels.
//
// Host work here is control plane only: buffer management, launch sequencing, the O(#distinct
// counts) piece table of the exact rank closed form, and result marshaling.  Every per-base and
// per-k-mer computation runs in the kernels of ks_kernels.cuh / ks_sort.cuh.  There is no CPU
// compute path: without a usable device every entry point fails with KS_ERR_CUDA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <exception>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kspans.h"
#include "ks_kernels.cuh"
#include "ks_count.cuh"
#include "ks_large.cuh"
#include "ks_layout.h"
#include "ks_rankseg.h"
#include "ks_sort.cuh"
#include "ks_window.cuh"
#include "ks_xgpu.cuh"

using namespace ks;

namespace {

std::string g_create_error;

struct DBuf {  // grow-only device buffer
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes, bool keep = false, cudaStream_t st = 0) {
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = bytes + bytes / 4 + 256;
    void *np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { cudaFree(np); return e; }
    }
    if (p) cudaFree(p);
    p = np;
    cap = ncap;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Host worker threads that live as long as the context: the staging memcpy runs once per 16 MiB window, and
// creating threads for every window costs more than the copy.  Used by one caller thread at a time.
struct WorkPool {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv_start, cv_done;
  std::function<void(unsigned)> job;
  unsigned long gen = 0;
  unsigned pending = 0;
  bool stop = false;
  unsigned size() const { return (unsigned)th.size() + 1; }  // workers + the calling thread
  void ensure(unsigned nthreads) {
    while (th.size() + 1 < nthreads) {
      const unsigned id = (unsigned)th.size() + 1;
      unsigned long born;
      {  // a worker created after jobs have run must not pick up the finished job of an earlier generation
        std::lock_guard<std::mutex> lk(m);
        born = gen;
      }
      th.emplace_back([this, id, born]() {
        unsigned long seen = born;
        for (;;) {
          std::function<void(unsigned)> f;
          {
            std::unique_lock<std::mutex> lk(m);
            cv_start.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            f = job;
          }
          f(id);
          {
            std::lock_guard<std::mutex> lk(m);
            if (--pending == 0) cv_done.notify_one();
          }
        }
      });
    }
  }
  // f(t) for t = 0 .. size()-1, t = 0 on the calling thread; returns when all are done
  void run(const std::function<void(unsigned)> &f) {
    if (th.empty()) { f(0); return; }
    {
      std::lock_guard<std::mutex> lk(m);
      job = f;
      pending = (unsigned)th.size();
      ++gen;
    }
    cv_start.notify_all();
    f(0);
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return pending == 0; });
    job = nullptr;  // the callable may reference the caller's stack
  }
  ~WorkPool() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv_start.notify_all();
    for (auto &t : th) t.join();
  }
};

}  // namespace

struct ks_seqset {
  ks_ctx *ctx = nullptr;
  uint8_t *d_buf = nullptr;  // layout of ks_layout.h
  bool owned = false;
  // A WINDOW set keeps only bytes [win_lo, win_hi) of the layout in HBM (one shard of a multi-GPU run: its own
  // range plus the head of the sequence the range starts in).  d_buf / d_pk / d_brk are then VIRTUAL bases
  // (allocation - offset), so every kernel keeps addressing by global position; *_alloc are the allocations.
  int64_t win_lo = 0, win_hi = 0;
  bool window = false;
  uint8_t *buf_alloc = nullptr;
  uint32_t *pk_alloc = nullptr;
  uint16_t *brk_alloc = nullptr;
  int64_t total = 0;         // bytes of the layout (multiple of 16)
  int64_t bases = 0;
  int nseq = 0;
  std::vector<int64_t> lens, starts;  // starts: nseq + 1
  int64_t *d_starts = nullptr;
  // K1 output: 2-bit packed codes + break masks, one entry per 16 positions (written by the pack+count
  // pass, read by every scan pass)
  uint32_t *d_pk = nullptr;
  uint16_t *d_brk = nullptr;
  mutable bool packed = false;
  // capacities (a set cached in the context is re-used by the host-buffer entry points)
  size_t cap_buf = 0, cap_chunks = 0, cap_starts = 0;
};

struct ks_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  int last_levels = 0;
  uint64_t last_revisit_chunks = 0;
  // scan scratch
  DBuf wfx, prm;
  DBuf rec_beg, rec_pk, rec_c, rec_mhi, rec_mlo, rec_count;
  size_t rec_cap = 0;
  DBuf seg_start, seg_len, seg_chunks, seg_chunk0, scan_tmp;
  DBuf sort_keys_a, sort_keys_b, sort_vals_a, sort_vals_b, sort_hist, sort_scan;
  DBuf out_pos, out_score;
  // score scratch
  DBuf sc_keys_a, sc_keys_b, sc_vals_a, sc_vals_b, sc_small, sc_gcount, sc_gstart, sc_segfirst, sc_segj0,
      sc_segx0, sc_seginc, sc_lut, sc_dense;
  // count -> score function of the last ks_dev_scores(LOG2 | SIGN): one value per distinct count
  DBuf lut_fx, lut_spc, lut_spv, core_lut, core_cc;
  std::vector<uint32_t> core_groups;  // group (distinct count) behind every class byte of the core records
  bool lut_valid = false;
  int lut_k = 0;
  std::vector<uint32_t> lut_gcount;
  std::vector<double> lut_gval;
  // staging / misc
  cudaStream_t copy_stream = nullptr;     // H2D of the sequence / D2H of tables, overlapped with kernels
  cudaEvent_t ev_copy = nullptr, ev_compute = nullptr, ev_table = nullptr;
  bool side_table_opt = false; // ks_ctx_side_table: the same for the stage calls (joined by the next scan / sync)
  bool defer_table = false;    // ks_dev_pipeline: the score TABLE (an output the scan does not read) is written on the
  bool table_pending = false;  // copy stream, next to the scan; joined before the call returns
  ks_seqset *host_set = nullptr;          // device buffers re-used by the host-buffer entry points
  void *pinned = nullptr;
  size_t pinned_cap = 0;
  DBuf tmp_counts, tmp_scores, tmp_inscan, nwords, pending, foc_hist, foc_big;
  // large k (ks_large.cuh): hash table, composite sort keys, slot indices, ranks in (count, code) order
  DBuf lg_slots, lg_stats, lg_comp_a, lg_comp_b, lg_slot_a, lg_slot_b, lg_ranks;
  uint64_t lg_nd = 0, lg_mask = 0;
  int lg_k = 0;
  uint64_t *lg_sorted = nullptr;   // sorted keys (one of lg_comp_a / lg_comp_b): composite, or codes in code order
  uint32_t *lg_idx = nullptr, *lg_cnt = nullptr;  // two-pass order: code-order index and count per position
  DBuf lg_cnt_a, lg_cnt_b, lg_idx_a, lg_idx_b;
  DBuf bk_table_a;         // bucketed counting: counts of the a.c k-mers per bucket, folded into the table at the end
  DBuf bk_buf, bk_cursor;  // bucketed counting (ks_count.cuh): sub-keys per bucket, fill of every bucket
  bool smem_attr_set = false, core_attr_set = false;
  DBuf win_match, win_cnt, win_pre, win_scratch, win_codes, win_fix, win_hist, win_pos;
  DBuf st_aux, child_pk, child_c, child_count, tr_tables;
  DBuf st_mn, st_mx, st_bm, detail, detail_count;
  // table copies into pageable host memory (what R hands over): a helper thread stages them through the
  // pinned windows and spreads the final memcpy over several threads while the scan runs
  WorkPool pool;  // host threads of the staging copies
  struct OutJob { void *dst; const void *src; size_t bytes; };
  std::vector<OutJob> out_jobs;
  std::thread out_thread;
  int out_rc = 0;
  // pinned host scratch for the small control-plane tables (histogram readback, count -> class / score
  // tables, scan parameters): asynchronous copies without a synchronisation per table
  char *hpin = nullptr;
  static constexpr size_t HPIN_HIST = 0, HPIN_SMALL = 256u << 10, HPIN_CLS = 320u << 10, HPIN_GCOUNT = 512u << 10,
                          HPIN_DENSE = 768u << 10, HPIN_LUT = 1280u << 10, HPIN_PRM = 1792u << 10,
                          HPIN_FX = 2048u << 10,  // the scan's per-class fixed-point table: its own region, because the
                                                  // score stage may still be copying out of HPIN_LUT when the scan starts
                          HPIN_COREFX = 2560u << 10,  // scores of the classes the core records name (255 x 8 B)
                          HPIN_CC = 2564u << 10,      // group -> core class bytes (<= 64 KiB)
                          HPIN_BYTES = 2628u << 10;
  // class table of the last ks_dev_scores(LOG2 | SIGN): index of every k-mer's count among the distinct counts
  DBuf cls, cls_dense, core;
  // rank order of the last ks_dev_scores(RANK): position of every k-mer, piece starts, bucket table (the pieces'
  // x0 / inc stay in sc_segx0 / sc_seginc); scan_ranks_impl gathers 4-byte positions instead of 8-byte scores
  DBuf rk_pos, rk_p0, rk_blob, rk_tail, rk_core;
  bool rk_valid = false;
  int rk_k = 0, rk_shift = 0;
  uint32_t rk_npieces = 0, rk_win_lo = 0, rk_win_len = 0, rk_n = 0;
  double rk_max = 0;
  bool core_valid = false;           // ctx->core holds the core records of ctx->cls (scan_gather_kernel, core mode)
  const void *cls_counts = nullptr;  // the count table it was derived from
  size_t cls_n = 0;
  size_t child_cap = 0;
  DBuf st_c, st_s, st_ea, st_eb, st_flags, st_p0, tile_xf, tile_ex, group_xf, group_S, group_ex, pending_list, pending_count, launch_rec;

  // timing / profiling
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool profile = false;
  struct ProfEv { int which; cudaEvent_t a, b; };
  std::vector<ProfEv> prof_pending;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[KS_PROF_N] = {0, 0, 0, 0, 0};
  uint64_t prof_n[KS_PROF_N] = {0, 0, 0, 0, 0};
  cudaEvent_t get_event() {
    cudaEvent_t e = nullptr;
    if (!ev_pool.empty()) { e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
  }
  cudaEvent_t prof_begin() {
    if (!profile) return nullptr;
    cudaEvent_t a = get_event();
    cudaEventRecord(a, stream);
    return a;
  }
  void prof_end(int which, cudaEvent_t a) {
    if (!a) return;
    cudaEvent_t b = get_event();
    cudaEventRecord(b, stream);
    prof_pending.push_back({which, a, b});
  }
  void prof_resolve() {
    if (prof_pending.empty()) return;
    cudaStreamSynchronize(stream);
    if (getenv("KS_PROF_GAPS")) {  // debugging aid: idle time between consecutive profiled stages
      for (size_t i = 0; i + 1 < prof_pending.size(); ++i) {
        float g = 0;
        if (cudaEventElapsedTime(&g, prof_pending[i].b, prof_pending[i + 1].a) == cudaSuccess)
          fprintf(stderr, "[ks gaps] after stage %d before stage %d: %.3f ms\n", prof_pending[i].which,
                  prof_pending[i + 1].which, g);
      }
    }
    for (auto &p : prof_pending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { prof_ms[p.which] += ms; prof_n[p.which] += 1; }
      ev_pool.push_back(p.a);
      ev_pool.push_back(p.b);
    }
    prof_pending.clear();
  }

  int fail(int code, const char *fmt, ...) {
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    err = b;
    return code;
  }
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return ctx->fail(KS_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, \
                       __LINE__, cudaGetErrorString(e_));                                     \
  } while (0)
#define LAUNCHED(n) (ctx->launches += (uint64_t)(n))
// no C++ exception may cross the C ABI (the caller is R / ctypes): host allocations that fail become an error code
#define KS_TRY try {
#define KS_CATCH(c)                                                                                     \
  }                                                                                                     \
  catch (const std::bad_alloc &) {                                                                      \
    ks_ctx *c_ = (c);                                                                                   \
    return c_ ? c_->fail(KS_ERR_NOMEM, "out of host memory") : KS_ERR_NOMEM;                            \
  }                                                                                                     \
  catch (const std::exception &e) {                                                                     \
    ks_ctx *c_ = (c);                                                                                   \
    return c_ ? c_->fail(KS_ERR_CUDA, "internal error: %s", e.what()) : KS_ERR_CUDA;                    \
  }

static inline unsigned grid_for(size_t n, int threads, unsigned cap = 148u * 16u) {
  size_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}
static inline unsigned blocks_exact(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// passes of the counting kernel: slices of at most 64 MiB of the table per pass (16 at most)
static void count_parts(int k, int *nparts, int *part_shift) {
  size_t bytes = ((size_t)4) << (2 * k);
  int lg = 0;
  while ((bytes >> lg) > ((size_t)64 << 20) && lg < 4) ++lg;
  if (lg & 1) ++lg;  // whole bases
  if (lg > 4) lg = 4;
  *nparts = 1 << lg;
  *part_shift = lg ? 2 * k - lg : 32;
}

static int check_k(ks_ctx *ctx, int k) {
  if (k < 1 || k > 15) return ctx->fail(KS_ERR_ARG, "k must be a positive integer less than 16 (got %d)", k);
  return KS_OK;
}

// ================================================================================================
extern "C" {

void ks_spans_free(ks_spans *s) {
  if (!s) return;
  free(s->pos);
  free(s->score);
  s->pos = nullptr;
  s->score = nullptr;
  s->n = 0;
}

int ks_ctx_create(ks_ctx **out, int device) {
  KS_TRY
  if (!out) return KS_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                     "); kspans has no CPU path";
    cudaGetLastError();
    return KS_ERR_CUDA;
  }
  if (device < 0) {
    e = cudaGetDevice(&device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return KS_ERR_CUDA; }
  }
  if (device >= ndev) { g_create_error = "device index out of range"; return KS_ERR_ARG; }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return KS_ERR_CUDA; }
  ks_ctx *ctx = new ks_ctx();
  ctx->device = device;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return KS_ERR_CUDA; }
  *out = ctx;
  return KS_OK;
  KS_CATCH(((ks_ctx *)nullptr))
}

static int join_table(ks_ctx *ctx);
void ks_ctx_destroy(ks_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->out_thread.joinable()) ctx->out_thread.join();
  join_table(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  DBuf *all[] = {&ctx->wfx, &ctx->prm, &ctx->rec_beg, &ctx->rec_pk, &ctx->rec_c,
                 &ctx->rec_mhi, &ctx->rec_mlo, &ctx->rec_count, &ctx->seg_start, &ctx->seg_len,
                 &ctx->seg_chunks, &ctx->seg_chunk0, &ctx->scan_tmp, &ctx->sort_keys_a, &ctx->sort_keys_b,
                 &ctx->sort_vals_a, &ctx->sort_vals_b, &ctx->sort_hist, &ctx->sort_scan, &ctx->out_pos,
                 &ctx->out_score, &ctx->sc_keys_a, &ctx->sc_keys_b, &ctx->sc_vals_a, &ctx->sc_vals_b,
                 &ctx->sc_small, &ctx->sc_gcount, &ctx->sc_gstart, &ctx->sc_segfirst, &ctx->sc_segj0,
                 &ctx->sc_segx0, &ctx->sc_seginc, &ctx->sc_lut, &ctx->sc_dense, &ctx->tmp_counts, &ctx->tmp_scores,
                 &ctx->tmp_inscan, &ctx->nwords, &ctx->pending, &ctx->st_c, &ctx->st_s, &ctx->st_ea, &ctx->st_eb, &ctx->st_flags, &ctx->st_p0, &ctx->tile_xf, &ctx->tile_ex, &ctx->group_xf, &ctx->group_S, &ctx->group_ex, &ctx->launch_rec, &ctx->pending_list, &ctx->pending_count, &ctx->foc_hist, &ctx->foc_big, &ctx->lut_fx, &ctx->lut_spc, &ctx->lut_spv, &ctx->core_lut, &ctx->core_cc,
                 &ctx->win_match, &ctx->win_cnt, &ctx->win_pre, &ctx->win_scratch, &ctx->win_codes, &ctx->win_fix,
                 &ctx->win_hist, &ctx->win_pos, &ctx->st_aux, &ctx->child_pk, &ctx->child_c, &ctx->child_count,
                 &ctx->tr_tables, &ctx->st_mn, &ctx->st_mx, &ctx->st_bm, &ctx->detail, &ctx->detail_count,
                 &ctx->cls, &ctx->cls_dense, &ctx->core, &ctx->rk_pos, &ctx->rk_p0, &ctx->rk_blob, &ctx->rk_tail, &ctx->rk_core,
                 &ctx->bk_buf, &ctx->bk_cursor, &ctx->bk_table_a, &ctx->lg_slots, &ctx->lg_stats, &ctx->lg_comp_a, &ctx->lg_comp_b,
                 &ctx->lg_slot_a, &ctx->lg_slot_b, &ctx->lg_ranks, &ctx->lg_cnt_a, &ctx->lg_cnt_b, &ctx->lg_idx_a,
                 &ctx->lg_idx_b};
  for (DBuf *b : all) b->release();
  ctx->prof_resolve();
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->t0) cudaEventDestroy(ctx->t0);
  if (ctx->t1) cudaEventDestroy(ctx->t1);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->hpin) cudaFreeHost(ctx->hpin);
  if (ctx->host_set) ks_seqset_free(ctx->host_set);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->ev_compute) cudaEventDestroy(ctx->ev_compute);
  if (ctx->ev_table) cudaEventDestroy(ctx->ev_table);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *ks_last_error(const ks_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void *ks_ctx_stream(ks_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
// the score table written on the copy stream (ks_dev_pipeline, ks_ctx_side_table): make it part of the compute
// stream's order and wait for it
static int join_table(ks_ctx *ctx) {
  if (!ctx->table_pending) return KS_OK;
  ctx->table_pending = false;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_table, 0));
  CK(cudaEventSynchronize(ctx->ev_table));
  return KS_OK;
}
int ks_ctx_sync(ks_ctx *ctx) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  int rc = join_table(ctx);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return KS_OK;
  KS_CATCH(ctx)
}
void ks_ctx_side_table(ks_ctx *ctx, int on) { if (ctx) ctx->side_table_opt = on != 0; }
uint64_t ks_ctx_launches(const ks_ctx *ctx) { return ctx ? ctx->launches : 0; }
void ks_ctx_reset_launches(ks_ctx *ctx) { if (ctx) ctx->launches = 0; }
void ks_ctx_scan_stats(const ks_ctx *ctx, int *levels, uint64_t *revisited_chunks) {
  if (levels) *levels = ctx ? ctx->last_levels : 0;
  if (revisited_chunks) *revisited_chunks = ctx ? ctx->last_revisit_chunks : 0;
}

int ks_ctx_timer_start(ks_ctx *ctx) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->t0) { CK(cudaEventCreate(&ctx->t0)); CK(cudaEventCreate(&ctx->t1)); }
  CK(cudaEventRecord(ctx->t0, ctx->stream));
  return KS_OK;
  KS_CATCH(ctx)
}
int ks_ctx_timer_stop(ks_ctx *ctx, float *ms) {
  KS_TRY
  if (!ctx || !ctx->t0) return KS_ERR_ARG;
  CK(cudaEventRecord(ctx->t1, ctx->stream));
  CK(cudaEventSynchronize(ctx->t1));
  float v = 0;
  CK(cudaEventElapsedTime(&v, ctx->t0, ctx->t1));
  if (ms) *ms = v;
  return KS_OK;
  KS_CATCH(ctx)
}
void ks_ctx_set_profile(ks_ctx *ctx, int on) { if (ctx) ctx->profile = on != 0; }
int ks_ctx_profile_get(ks_ctx *ctx, int which, double *ms_total, uint64_t *launches) {
  KS_TRY
  if (!ctx || which < 0 || which >= KS_PROF_N) return KS_ERR_ARG;
  ctx->prof_resolve();
  if (ms_total) *ms_total = ctx->prof_ms[which];
  if (launches) *launches = ctx->prof_n[which];
  return KS_OK;
  KS_CATCH(ctx)
}
void ks_ctx_profile_reset(ks_ctx *ctx) {
  if (!ctx) return;
  ctx->prof_resolve();
  for (int i = 0; i < KS_PROF_N; ++i) { ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
}

int ks_kmer_seq(int k, uint64_t code, char *out) {
  KS_TRY
  static const char nuc[4] = {'A', 'C', 'T', 'G'};
  if (k < 1 || k > 16 || !out) return KS_ERR_ARG;
  out[k] = 0;
  for (int j = k - 1; j >= 0; --j) { out[j] = nuc[code & 3]; code >>= 2; }
  return KS_OK;
  KS_CATCH(((ks_ctx *)nullptr))
}

// ------------------------------------------------------------------------------------------------
// stage: count.  Three ways to count (ks_count.cuh): shared-memory table (k <= 7), buckets through shared memory
// with one sub-key per two k-mers (8 <= k <= 13), direct global reductions (k >= 14 in slices of the table, and
// inputs too small to pay for the extra launches).  KS_COUNT_PATH = direct | smem | bucket forces one (tests, measurements).
namespace {
enum { COUNT_DIRECT = 0, COUNT_SMEM = 1, COUNT_BUCKET = 2 };
struct CountRun {
  int path = COUNT_DIRECT;
  int k = 0;
  int32_t *d_counts = nullptr;
  uint32_t kmask = 0, gcap = 0;
  int nparts = 1, part_shift = 32;
};
}  // namespace

// the bucket kernels are compiled per k (shift counts and masks as immediates)
#define KS_BUCKET_K(k, stmt)                    \
  switch (k) {                                  \
    case 8: { constexpr int KK = 8; stmt; } break;   \
    case 9: { constexpr int KK = 9; stmt; } break;   \
    case 10: { constexpr int KK = 10; stmt; } break; \
    case 11: { constexpr int KK = 11; stmt; } break; \
    case 12: { constexpr int KK = 12; stmt; } break; \
    default: { constexpr int KK = 13; stmt; } break; \
  }
extern "C++" {
template <int K>
static cudaError_t bucket_smem_attrs_k() {
  cudaError_t e = cudaFuncSetAttribute(bucket_scatter_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)bk_scatter_smem(K));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(bucket_count_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
}
}  // extern "C++"
static cudaError_t bucket_smem_attrs() {
  cudaError_t e;
  if ((e = bucket_smem_attrs_k<8>()) != cudaSuccess) return e;
  if ((e = bucket_smem_attrs_k<9>()) != cudaSuccess) return e;
  if ((e = bucket_smem_attrs_k<10>()) != cudaSuccess) return e;
  if ((e = bucket_smem_attrs_k<11>()) != cudaSuccess) return e;
  if ((e = bucket_smem_attrs_k<12>()) != cudaSuccess) return e;
  return bucket_smem_attrs_k<13>();
}

// zeroes the table, the word count and (bucket path) the bucket fills; est_chunks sizes the bucket regions
static int count_begin(ks_ctx *ctx, int k, int32_t *d_counts, int64_t est_chunks, CountRun *run) {
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)1 << (2 * k);
  run->k = k;
  run->d_counts = d_counts;
  run->kmask = (uint32_t)(n - 1);
  count_parts(k, &run->nparts, &run->part_shift);
  int path = COUNT_DIRECT;
  if (k <= 7 && est_chunks >= (1 << 14)) path = COUNT_SMEM;
  else if (k >= 8 && k <= 13 && est_chunks >= (1 << 18)) path = COUNT_BUCKET;
  if (const char *e = getenv("KS_COUNT_PATH")) {
    if (!strcmp(e, "direct")) path = COUNT_DIRECT;
    else if (!strcmp(e, "smem") && k <= 7) path = COUNT_SMEM;
    else if (!strcmp(e, "bucket") && k >= 8 && k <= 13) path = COUNT_BUCKET;
  }
  run->path = path;
  if (!ctx->smem_attr_set) {
    CK(cudaFuncSetAttribute(pack_count_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(bucket_smem_attrs());
    ctx->smem_attr_set = true;
  }
  CK(ctx->nwords.ensure(sizeof(unsigned long long)));
  CK(cudaMemsetAsync(d_counts, 0, n * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ctx->nwords.p, 0, sizeof(unsigned long long), st));
  if (path == COUNT_BUCKET) {
    const size_t nbuckets = (size_t)1 << bk_log(k);
    uint64_t per = (uint64_t)(est_chunks > 0 ? est_chunks : 1) * 8 / nbuckets;  // pairs per bucket
    uint64_t gcap = per + per / 2 + 8192;  // spectrum skew; what does not fit overflows to the direct reduction
    if (const char *e = getenv("KS_BUCKET_CAP")) gcap = (uint64_t)atoll(e);  // tests: force the overflow path
    gcap = (gcap + 7) & ~7ull;
    if (gcap < 8) gcap = 8;
    if (gcap > 0xfffffff0ull) return ctx->fail(KS_ERR_ARG, "input too large for the bucketed count");
    run->gcap = (uint32_t)gcap;
    CK(ctx->bk_buf.ensure((size_t)gcap * nbuckets * 2));
    CK(ctx->bk_cursor.ensure(BK_MAX_BUCKETS * 4));
    CK(ctx->bk_table_a.ensure(n * 4));
    CK(cudaMemsetAsync(ctx->bk_cursor.p, 0, BK_MAX_BUCKETS * 4, st));
  }
  return KS_OK;
}

// chunks [first, first + nchunks) of the set: pack them and count (or scatter) their k-mers
static int count_chunks(ks_ctx *ctx, const ks_seqset *s, const CountRun &run, int64_t first, int64_t nchunks) {
  if (nchunks <= 0) return KS_OK;
  cudaStream_t st = ctx->stream;
  unsigned long long *nw = ctx->nwords.as<unsigned long long>();
  if (run.path == COUNT_SMEM) {
    const size_t smem = (size_t)4 << (2 * run.k);
    const unsigned per_sm = smem > 32768 ? 3u : 6u;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)148 * per_sm, (size_t)((nchunks + 1023) / 1024));
    pack_count_smem_kernel<<<grid, 256, smem, st>>>(s->d_buf, first, nchunks, run.k, run.kmask, s->d_pk, s->d_brk,
                                                   run.d_counts, nw);
  } else if (run.path == COUNT_BUCKET) {
    const int64_t tile_chunks = bk_tile_chunks(run.k);
    const int64_t ntiles = (nchunks + tile_chunks - 1) / tile_chunks;
    const unsigned grid = (unsigned)std::min<int64_t>(run.k >= 13 ? 148 : 148 * 4, ntiles);
    KS_BUCKET_K(run.k, (bucket_scatter_kernel<KK><<<grid, bk_threads(KK), bk_scatter_smem(KK), st>>>(
        s->d_buf, first, nchunks, s->d_pk, s->d_brk, run.d_counts, nw, ctx->bk_buf.as<uint16_t>(),
        ctx->bk_cursor.as<uint32_t>(), run.gcap)));
  } else {
    pack_count_kernel<true><<<grid_for((size_t)nchunks, 256, 148u * 8u), 256, 0, st>>>(
        s->d_buf, first, nchunks, run.k, run.kmask, s->d_pk, s->d_brk, run.d_counts, nw, run.part_shift, 0u);
  }
  LAUNCHED(1);
  CK(cudaGetLastError());
  return KS_OK;
}

// after the last count_chunks: phase 2 of the bucket path, or the remaining table slices of the direct path over
// all chunks [first, first + nchunks) that were counted
static int count_end(ks_ctx *ctx, const ks_seqset *s, const CountRun &run, int64_t first, int64_t nchunks) {
  cudaStream_t st = ctx->stream;
  if (run.path == COUNT_BUCKET) {
    const size_t nk = (size_t)1 << (2 * run.k);
    KS_BUCKET_K(run.k, (bucket_count_kernel<KK><<<1 << bk_log(KK), BK_COUNT_THREADS, 2 * (nk >> bk_log(KK)) * 4, st>>>(
        ctx->bk_buf.as<uint16_t>(), ctx->bk_cursor.as<uint32_t>(), run.gcap, run.d_counts,
        ctx->bk_table_a.as<uint32_t>())));
    KS_BUCKET_K(run.k, (bucket_fold_kernel<KK><<<grid_for(nk / 4, 256, 148u * 8u), 256, 0, st>>>(
        run.d_counts, ctx->bk_table_a.as<uint32_t>())));
    LAUNCHED(2);
  } else if (run.path == COUNT_DIRECT && nchunks > 0) {
    for (int part = 1; part < run.nparts; ++part) {  // tables beyond L2: one pass over the sequence per slice
      pack_count_kernel<true><<<grid_for((size_t)nchunks, 256, 148u * 8u), 256, 0, st>>>(
          s->d_buf, first, nchunks, run.k, run.kmask, s->d_pk, s->d_brk, run.d_counts,
          ctx->nwords.as<unsigned long long>(), run.part_shift, (uint32_t)part);
      LAUNCHED(1);
    }
  }
  CK(cudaGetLastError());
  return KS_OK;
}

// ------------------------------------------------------------------------------------------------
// sequence sets
// (re)initialise a set for these lengths; device buffers only grow
static int seqset_prepare(ks_ctx *ctx, const int64_t *lens, int nseq, ks_seqset *s, bool own_buffer,
                          int64_t win_lo = 0, int64_t win_hi = -1) {
  s->ctx = ctx;
  s->nseq = nseq;
  s->packed = false;
  s->lens.assign(lens, lens + nseq);
  s->starts.resize((size_t)nseq + 1);
  s->total = ks_layout_total(lens, nseq, s->starts.data());
  s->bases = 0;
  for (int i = 0; i < nseq; ++i) {
    if (lens[i] < 0 || lens[i] > 2147483646LL) return ctx->fail(KS_ERR_ARG, "sequence %d: length out of range", i);
    s->bases += lens[i];
  }
  if ((size_t)nseq + 1 > s->cap_starts) {
    if (s->d_starts) cudaFree(s->d_starts);
    s->d_starts = nullptr;
    s->cap_starts = (size_t)nseq + 1 + (size_t)nseq / 4;
    CK(cudaMalloc(&s->d_starts, sizeof(int64_t) * s->cap_starts));
  }
  CK(cudaMemcpyAsync(s->d_starts, s->starts.data(), sizeof(int64_t) * ((size_t)nseq + 1),
                     cudaMemcpyHostToDevice, ctx->stream));
  if (win_hi < 0 || win_hi > s->total) win_hi = s->total;
  if (win_lo < 0) win_lo = 0;
  win_lo &= ~15ll;
  win_hi = (win_hi + 15) & ~15ll;
  if (win_hi > s->total) win_hi = s->total;
  if (win_lo >= win_hi) { win_lo = 0; win_hi = 16; }  // an empty shard keeps the front pad only
  s->win_lo = win_lo;
  s->win_hi = win_hi;
  s->window = !(win_lo == 0 && win_hi == s->total);
  if (s->window && !own_buffer) return ctx->fail(KS_ERR_ARG, "a wrapped buffer cannot be a window set");
  size_t nch = (size_t)((win_hi - win_lo) / 16) + 8;
  if (nch > s->cap_chunks) {
    if (s->pk_alloc) cudaFree(s->pk_alloc);
    if (s->brk_alloc) cudaFree(s->brk_alloc);
    s->pk_alloc = nullptr; s->brk_alloc = nullptr;
    s->cap_chunks = nch + nch / 8;
    CK(cudaMalloc(&s->pk_alloc, s->cap_chunks * sizeof(uint32_t)));
    CK(cudaMalloc(&s->brk_alloc, s->cap_chunks * sizeof(uint16_t)));
  }
  s->d_pk = s->pk_alloc - win_lo / 16;
  s->d_brk = s->brk_alloc - win_lo / 16;
  // beyond the data every position is a break; the pack pass overwrites the resident chunks
  CK(cudaMemsetAsync(s->brk_alloc + (nch - 8), 0xff, 8 * sizeof(uint16_t), ctx->stream));
  CK(cudaMemsetAsync(s->pk_alloc + (nch - 8), 0, 8 * sizeof(uint32_t), ctx->stream));
  if (own_buffer) {
    size_t need = (size_t)(win_hi - win_lo) + KS_SLACK;
    if (need > s->cap_buf) {
      if (s->buf_alloc && s->owned) cudaFree(s->buf_alloc);
      s->buf_alloc = nullptr;
      s->cap_buf = need + need / 8;
      if (cudaMalloc(&s->buf_alloc, s->cap_buf) != cudaSuccess) {
        cudaGetLastError();
        s->cap_buf = 0;
        return ctx->fail(KS_ERR_NOMEM, "cudaMalloc(%lld) failed", (long long)need);
      }
    }
    s->d_buf = s->buf_alloc - win_lo;
    s->owned = true;
  }
  return KS_OK;
}

static int ensure_copy_stream(ks_ctx *ctx) {
  if (ctx->copy_stream) return KS_OK;
  CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
  return KS_OK;
}

static bool host_ptr_is_pinned(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// H2D of the sequences into the layout of ks_layout.h on the copy stream.  With count_k > 0 the
// pack+count kernel is launched on the compute stream slab by slab behind the copies (it is far
// faster than PCIe, so counting hides completely behind the upload).
static int upload_impl(ks_ctx *ctx, ks_seqset *s, const char *const *seqs, const int64_t *lens, int nseq,
                       int count_k, int32_t *d_counts) {
  int rc = ensure_copy_stream(ctx);
  if (rc) return rc;
  cudaStream_t cs = ctx->copy_stream, st = ctx->stream;
  CK(cudaEventRecord(ctx->ev_compute, st));       // the buffers may still be read by earlier work
  CK(cudaStreamWaitEvent(cs, ctx->ev_compute, 0));
  CK(cudaMemsetAsync(s->d_buf + s->win_lo, 0, (size_t)(s->win_hi - s->win_lo) + KS_SLACK, cs));
  if (s->window && count_k) return ctx->fail(KS_ERR_ARG, "a window set is counted by range (ks_dev_count_range)");
  const int64_t nchunks = (s->total - 16) / 16;
  CountRun run;
  if (count_k) {
    rc = count_begin(ctx, count_k, d_counts, nchunks, &run);
    if (rc) return rc;
  }
  int64_t done = 0;  // chunks [0, done) of the count grid are launched
  const int64_t SLAB = (24ll << 20) / 16;
  auto progress = [&](int64_t covered, bool final) -> cudaError_t {
    if (!count_k) return cudaSuccess;
    // chunk ci reads bytes [16 ci, 16 ci + 33): available once covered >= 16 ci + 33
    int64_t avail = final ? nchunks : (covered - 33) / 16 + 1;
    if (avail > nchunks) avail = nchunks;
    if (avail <= done || (!final && avail - done < SLAB)) return cudaSuccess;
    cudaError_t e2 = cudaEventRecord(ctx->ev_copy, cs);
    if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(st, ctx->ev_copy, 0);
    if (e2 != cudaSuccess) return e2;
    cudaEvent_t pe = ctx->prof_begin();
    const int crc = count_chunks(ctx, s, run, done, avail - done);
    ctx->prof_end(KS_PROF_COUNT, pe);
    done = avail;
    return crc ? cudaErrorUnknown : cudaSuccess;
  };
  // large sequences go straight from the caller's memory; small ones are packed into pinned
  // staging windows (two halves, alternating) so that 100k contigs do not cost 100k copies
  const int64_t DIRECT = 4ll << 20;
  const size_t HALF = 16u << 20;
  if (!ctx->pinned) {
    CK(cudaMallocHost(&ctx->pinned, 2 * HALF));
    ctx->pinned_cap = 2 * HALF;
  }
  cudaEvent_t ev[2];
  cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
  bool ev_used[2] = {false, false};
  int half = 0;
  // plan: maximal groups of consecutive small sequences that fit one staging half
  struct Item { const char *src; size_t len; size_t off; bool sep; };  // sep: zero the byte in front (separator)
  std::vector<Item> items;
  size_t fill = 0;
  int64_t win_start = -1;  // global offset the current window maps to
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads > 12) nthreads = 12;
  if (const char *e = getenv("KS_STAGE_THREADS")) nthreads = (unsigned)atoi(e);
  if (nthreads < 1) nthreads = 1;
  auto flush = [&]() -> cudaError_t {
    if (fill == 0) return cudaSuccess;
    char *win = (char *)ctx->pinned + (size_t)half * HALF;
    cudaError_t ee = cudaSuccess;
    if (ev_used[half]) ee = cudaEventSynchronize(ev[half]);  // the copy that last used this half is done
    if (ee != cudaSuccess) return ee;
    // fill the window with several host threads (R hands over pageable, separately allocated strings)
    auto work = [&](unsigned t) {
      if (t >= nthreads) return;  // the pool may hold more threads than this call wants
      for (size_t i = t; i < items.size(); i += nthreads) {
        const Item &it = items[i];
        memcpy(win + it.off, it.src, it.len);
        if (it.sep) win[it.off - 1] = 0;  // separator in front of every sequence but the first of the window
      }
    };
    if (nthreads > 1 && fill > (1u << 20)) {
      ctx->pool.ensure(nthreads);
      ctx->pool.run(work);
    } else {
      for (unsigned t = 0; t < nthreads; ++t) work(t);
    }
    ee = cudaMemcpyAsync(s->d_buf + win_start, win, fill, cudaMemcpyHostToDevice, cs);
    if (ee != cudaSuccess) return ee;
    ee = progress(win_start + (int64_t)fill, false);
    if (ee != cudaSuccess) return ee;
    cudaEventRecord(ev[half], cs);
    ev_used[half] = true;
    half ^= 1;
    fill = 0;
    win_start = -1;
    items.clear();
    return ee;
  };
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < nseq && e == cudaSuccess; ++i) {
    int64_t ln = lens[i];
    if (ln == 0) continue;
    if (s->window) {
      // only the part of the sequence inside the window travels; plain copies (a shard holds few sequences)
      const int64_t a = std::max<int64_t>(s->starts[i], s->win_lo), b = std::min<int64_t>(s->starts[i] + ln, s->win_hi);
      if (a >= b) continue;
      if (!host_ptr_is_pinned(seqs[i])) {
        e = flush();
        const size_t SUB = 1u << 20;
        for (int64_t off = a - s->starts[i]; off < b - s->starts[i] && e == cudaSuccess;) {
          if (fill == HALF) { e = flush(); if (e != cudaSuccess) break; }
          if (fill == 0) win_start = s->starts[i] + off;
          size_t room = HALF - fill;
          size_t take = (size_t)(b - s->starts[i] - off) < room ? (size_t)(b - s->starts[i] - off) : room;
          for (size_t q = 0; q < take; q += SUB)
            items.push_back({seqs[i] + off + (int64_t)q, take - q < SUB ? take - q : SUB, fill + q, false});
          fill += take;
          off += (int64_t)take;
        }
        if (e == cudaSuccess) e = flush();
      } else {
        e = flush();
        if (e == cudaSuccess)
          e = cudaMemcpyAsync(s->d_buf + a, seqs[i] + (a - s->starts[i]), (size_t)(b - a), cudaMemcpyHostToDevice, cs);
      }
      continue;
    }
    if (ln >= DIRECT && !host_ptr_is_pinned(seqs[i])) {
      // a long sequence in pageable memory (what R hands over): the driver would stage it at a few GB/s;
      // packed into the pinned windows by the host threads it moves near the PCIe rate
      e = flush();
      const size_t SUB = 1u << 20;  // granule of the per-thread memcpy
      for (int64_t off = 0; off < ln && e == cudaSuccess;) {
        if (fill == HALF) { e = flush(); if (e != cudaSuccess) break; }
        if (fill == 0) win_start = s->starts[i] + off;
        size_t room = HALF - fill;
        size_t take = (size_t)(ln - off) < room ? (size_t)(ln - off) : room;
        for (size_t a = 0; a < take; a += SUB)
          items.push_back({seqs[i] + off + (int64_t)a, take - a < SUB ? take - a : SUB, fill + a, false});
        fill += take;
        off += (int64_t)take;
      }
      if (e == cudaSuccess) e = flush();
      continue;
    }
    if (ln >= DIRECT) {
      e = flush();
      // in pieces, so that counting can start while the rest is still on the bus
      const int64_t PIECE = 32ll << 20;
      for (int64_t off = 0; off < ln && e == cudaSuccess; off += PIECE) {
        int64_t n = ln - off < PIECE ? ln - off : PIECE;
        e = cudaMemcpyAsync(s->d_buf + s->starts[i] + off, seqs[i] + off, (size_t)n, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = progress(s->starts[i] + off + n, false);
      }
      continue;
    }
    // contiguous with the window?  (separator bytes between sequences are copied as zeros)
    if (fill && (s->starts[i] != win_start + (int64_t)fill + 1 || fill + 1 + (size_t)ln > HALF)) e = flush();
    if (e != cudaSuccess) break;
    bool sep = false;
    if (fill == 0) win_start = s->starts[i];
    else { fill += 1; sep = true; }
    items.push_back({seqs[i], (size_t)ln, fill, sep});
    fill += (size_t)ln;
  }
  if (e == cudaSuccess) e = flush();
  if (e == cudaSuccess) e = progress(s->total, true);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copy, cs);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ctx->ev_copy, 0);  // later kernels see the whole buffer
  if (e == cudaSuccess && count_k) {  // phase 2 of the bucket path / remaining table slices beyond L2
    cudaEvent_t pe = ctx->prof_begin();
    if (count_end(ctx, s, run, 0, nchunks)) e = cudaErrorUnknown;
    ctx->prof_end(KS_PROF_COUNT, pe);
  }
  if (e == cudaSuccess && ev_used[0]) e = cudaEventSynchronize(ev[0]);  // staging halves are re-used next call
  if (e == cudaSuccess && ev_used[1]) e = cudaEventSynchronize(ev[1]);
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  CK(e);
  if (count_k) s->packed = true;
  return KS_OK;
}

int ks_seqset_upload(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, ks_seqset **out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!out || !seqs || !lens || nseq < 1)
    return ctx->fail(KS_ERR_ARG, "seq_r must be a character vector of length at least one");
  CK(cudaSetDevice(ctx->device));
  ks_seqset *s = new ks_seqset();
  int rc = seqset_prepare(ctx, lens, nseq, s, true);
  if (!rc) rc = upload_impl(ctx, s, seqs, lens, nseq, 0, nullptr);
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = ctx->fail(KS_ERR_CUDA, "upload: %s", cudaGetErrorString(e));
  }
  if (rc) { ks_seqset_free(s); return rc; }
  *out = s;
  return KS_OK;
  KS_CATCH(ctx)
}

// One shard of a multi-GPU run: the layout of ALL sequences cut into nranks contiguous ranges of
// chunks; shard `rank` counts and scans chunks [chunk0, chunk0 + nchunks) and keeps bytes [win_lo, win_hi) of the
// layout resident: its own range plus the head of the sequence the range starts in (a span that closes in the
// range may have started there, and its re-scans run on this shard) and one chunk of slack behind it.
int ks_plan_shard(const int64_t *lens, int nseq, int nranks, int rank, int64_t *chunk0, int64_t *nchunks,
                  int64_t *win_lo, int64_t *win_hi) {
  KS_TRY
  if (!lens || nseq < 1 || nranks < 1 || rank < 0 || rank >= nranks) return KS_ERR_ARG;
  std::vector<int64_t> starts((size_t)nseq + 1);
  const int64_t total = ks_layout_total(lens, nseq, starts.data());
  const int64_t chunks = (total - 16) / 16;
  const int64_t per = (chunks + nranks - 1) / nranks;
  const int64_t c0 = std::min<int64_t>((int64_t)rank * per, chunks);
  const int64_t cn = std::min<int64_t>(per, chunks - c0);
  if (chunk0) *chunk0 = c0;
  if (nchunks) *nchunks = cn;
  int64_t lo = 0, hi = 16;
  if (cn > 0) {
    const int64_t first_pos = 16 * c0 + 16;  // first position of the range
    size_t i = (size_t)(std::upper_bound(starts.begin(), starts.begin() + nseq, first_pos) - starts.begin());
    i = i ? i - 1 : 0;                       // the sequence that holds it (or the one before a separator)
    lo = std::min<int64_t>(16 * c0, (starts[i] & ~15ll) - 16);
    if (lo < 0) lo = 0;
    hi = std::min<int64_t>(total, 16 * (c0 + cn) + 48);
  }
  if (win_lo) *win_lo = lo;
  if (win_hi) *win_hi = hi;
  return KS_OK;
  KS_CATCH(((ks_ctx *)nullptr))
}

// Upload only bytes [win_lo, win_hi) of the layout of these sequences (a window set; see ks_plan_shard).  The set
// is counted with ks_dev_count_range and scanned with ks_dev_scan*_shard on chunk ranges inside the window; span
// coordinates and sequence ids stay global.
int ks_seqset_upload_window(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int64_t win_lo,
                            int64_t win_hi, ks_seqset **out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!out || !seqs || !lens || nseq < 1)
    return ctx->fail(KS_ERR_ARG, "seq_r must be a character vector of length at least one");
  CK(cudaSetDevice(ctx->device));
  ks_seqset *s = new ks_seqset();
  int rc = seqset_prepare(ctx, lens, nseq, s, true, win_lo, win_hi);
  if (!rc) rc = upload_impl(ctx, s, seqs, lens, nseq, 0, nullptr);
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = ctx->fail(KS_ERR_CUDA, "upload: %s", cudaGetErrorString(e));
  }
  if (rc) { ks_seqset_free(s); return rc; }
  *out = s;
  return KS_OK;
  KS_CATCH(ctx)
}

// Upload new sequences into an existing set (device buffers are re-used, they only grow).  With count_k > 0
// the pack+count pass runs behind the copies (d_counts is overwritten) and the word count is left at
// d_nwords (device uint64, may be NULL).  Nothing is synchronised: pinned host buffers must stay valid until
// the next synchronising call on this ctx.
int ks_seqset_reupload(ks_ctx *ctx, ks_seqset *s, const char *const *seqs, const int64_t *lens, int nseq,
                       int count_k, int32_t *d_counts, uint64_t *d_nwords) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!s || !seqs || !lens || nseq < 1)
    return ctx->fail(KS_ERR_ARG, "seq_r must be a character vector of length at least one");
  if (!s->owned && s->d_buf) return ctx->fail(KS_ERR_ARG, "ks_seqset_reupload: the set wraps a caller-owned buffer");
  if (s->window) return ctx->fail(KS_ERR_ARG, "ks_seqset_reupload: not for window sets");
  if (count_k) {
    int rc = check_k(ctx, count_k);
    if (rc) return rc;
    if (!d_counts) return ctx->fail(KS_ERR_ARG, "ks_seqset_reupload: null count table");
  }
  CK(cudaSetDevice(ctx->device));
  int rc = seqset_prepare(ctx, lens, nseq, s, true);
  if (!rc) rc = upload_impl(ctx, s, seqs, lens, nseq, count_k, d_counts);
  if (!rc && count_k && d_nwords)
    CK(cudaMemcpyAsync(d_nwords, ctx->nwords.p, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return rc;
  KS_CATCH(ctx)
}

// the cached set of the host-buffer entry points
static int host_set_acquire(ks_ctx *ctx, const int64_t *lens, int nseq, ks_seqset **out) {
  if (!ctx->host_set) ctx->host_set = new ks_seqset();
  int rc = seqset_prepare(ctx, lens, nseq, ctx->host_set, true);
  if (rc) return rc;
  *out = ctx->host_set;
  return KS_OK;
}

int ks_seqset_wrap(ks_ctx *ctx, const void *d_buf, int64_t total_bytes, const int64_t *lens, int nseq,
                   ks_seqset **out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!out || !d_buf || !lens || nseq < 1) return ctx->fail(KS_ERR_ARG, "ks_seqset_wrap: bad arguments");
  CK(cudaSetDevice(ctx->device));
  ks_seqset *s = new ks_seqset();
  int rc = seqset_prepare(ctx, lens, nseq, s, false);
  if (rc) { ks_seqset_free(s); return rc; }
  if (total_bytes < s->total + KS_SLACK) {
    ks_seqset_free(s);
    return ctx->fail(KS_ERR_ARG, "ks_seqset_wrap: buffer holds %lld bytes, layout needs %lld", (long long)total_bytes,
                     (long long)(s->total + KS_SLACK));
  }
  s->d_buf = (uint8_t *)d_buf;
  s->buf_alloc = nullptr;
  s->owned = false;
  CK(cudaStreamSynchronize(ctx->stream));
  *out = s;
  return KS_OK;
  KS_CATCH(ctx)
}

void ks_seqset_free(ks_seqset *s) {
  if (!s) return;
  if (s->ctx) cudaSetDevice(s->ctx->device);
  if (s->owned && s->buf_alloc) cudaFree(s->buf_alloc);
  if (s->d_starts) cudaFree(s->d_starts);
  if (s->pk_alloc) cudaFree(s->pk_alloc);
  if (s->brk_alloc) cudaFree(s->brk_alloc);
  delete s;
}
int64_t ks_seqset_bases(const ks_seqset *s) { return s ? s->bases : 0; }
int64_t ks_seqset_buffer_bytes(const ks_seqset *s) { return s ? s->total + KS_SLACK : 0; }

int64_t ks_seqset_positions(const ks_seqset *s) { return s ? s->total : 0; }
int64_t ks_seqset_start(const ks_seqset *s, int seq) { return (s && seq >= 0 && seq < s->nseq) ? s->starts[seq] : -1; }

// pack only (no counting pass ran on this set)
static int ensure_packed(ks_ctx *ctx, const ks_seqset *s) {
  if (s->packed) return KS_OK;
  const int64_t c0 = s->win_lo / 16;
  int64_t nch = (s->win_hi - 16) / 16 - c0;
  if (nch < 0) nch = 0;
  cudaEvent_t pe = ctx->prof_begin();
  if (nch)
    pack_count_kernel<false><<<grid_for((size_t)nch, 256, 148u * 8u), 256, 0, ctx->stream>>>(
        s->d_buf, c0, nch, 1, 3u, s->d_pk, s->d_brk, nullptr, nullptr);
  ctx->prof_end(KS_PROF_COUNT, pe);
  LAUNCHED(1);
  CK(cudaGetLastError());
  s->packed = true;
  return KS_OK;
}

// ------------------------------------------------------------------------------------------------
// stage: count
static int dev_count_impl(ks_ctx *ctx, const ks_seqset *s, int k, int32_t *d_counts, double *n_words, bool sync);
int ks_dev_count(ks_ctx *ctx, const ks_seqset *s, int k, int32_t *d_counts, double *n_words) {
  KS_TRY
  return dev_count_impl(ctx, s, k, d_counts, n_words, true);
  KS_CATCH(ctx)
}
// sync = false: the number of words stays in ctx->nwords for the stage that follows on the stream
static int dev_count_impl(ks_ctx *ctx, const ks_seqset *s, int k, int32_t *d_counts, double *n_words, bool sync) {
  if (!ctx) return KS_ERR_ARG;
  if (!s || !d_counts) return ctx->fail(KS_ERR_ARG, "ks_dev_count: null argument");
  int rc = check_k(ctx, k);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (s->window) return ctx->fail(KS_ERR_ARG, "a window set is counted by range (ks_dev_count_range)");
  int64_t nchunks = (s->total - 16) / 16;
  CountRun run;
  rc = count_begin(ctx, k, d_counts, nchunks, &run);
  if (rc) return rc;
  cudaEvent_t pe = ctx->prof_begin();
  rc = count_chunks(ctx, s, run, 0, nchunks);
  if (!rc) rc = count_end(ctx, s, run, 0, nchunks);
  ctx->prof_end(KS_PROF_COUNT, pe);
  if (rc) return rc;
  s->packed = true;
  if (!sync) return KS_OK;
  unsigned long long nw = 0;
  CK(cudaMemcpyAsync(&nw, ctx->nwords.p, sizeof nw, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (n_words) *n_words = (double)nw;
  return KS_OK;
}

// multi-GPU composition without host round trips: the word count stays on the device (it is summed by a
// collective on the same stream) and is read back together with the first table the score stage needs
int ks_dev_count_async(ks_ctx *ctx, const ks_seqset *s, int k, int32_t *d_counts, uint64_t *d_nwords) {
  KS_TRY
  int rc = dev_count_impl(ctx, s, k, d_counts, nullptr, false);
  if (rc) return rc;
  if (d_nwords)
    CK(cudaMemcpyAsync(d_nwords, ctx->nwords.p, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return KS_OK;
  KS_CATCH(ctx)
}

// Sum of the count tables of all ranks over peer memory (ks_xgpu.cuh).  tables[nranks] = the address of every
// rank's buffer of n_u64 64-bit words (pairs of int32 counters, the word count last) as mapped into THIS
// process; mc_table = multicast address of the same buffer or NULL.  Rank `rank` sums its slice.  The caller
// puts a cross-GPU barrier on the ctx stream before and after.
int ks_dev_xsum(ks_ctx *ctx, void *const *tables, int nranks, int rank, void *mc_table, uint64_t n_u64) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!tables || nranks < 1 || nranks > XSUM_MAX_RANKS || rank < 0 || rank >= nranks)
    return ctx->fail(KS_ERR_ARG, "ks_dev_xsum: bad arguments");
  CK(cudaSetDevice(ctx->device));
  // slices in units of 2 words (16-byte vectors); the last rank takes the remainder
  const uint64_t pairs = n_u64 / 2, per = (pairs + nranks - 1) / nranks;
  uint64_t first = std::min<uint64_t>((uint64_t)rank * per, pairs) * 2;
  uint64_t last = std::min<uint64_t>((uint64_t)(rank + 1) * per, pairs) * 2;
  if (rank == nranks - 1) last = n_u64;
  const uint64_t count = last - first;
  if (count == 0) return KS_OK;
  if (mc_table) {
    xsum_multicast_kernel<<<grid_for((size_t)(count / 4 + 1), 256, 148u * 8u), 256, 0, ctx->stream>>>(
        reinterpret_cast<uint64_t *>(mc_table), (size_t)first, (size_t)count);
  } else {
    XsumPeers P;
    for (int r = 0; r < nranks; ++r) P.table[r] = reinterpret_cast<uint64_t *>(tables[r]);
    xsum_peer_kernel<<<grid_for((size_t)(count / 2 + 1), 256, 148u * 4u), 256, 0, ctx->stream>>>(P, nranks, (size_t)first,
                                                                                         (size_t)count);
  }
  LAUNCHED(1);
  CK(cudaGetLastError());
  return KS_OK;
  KS_CATCH(ctx)
}

int64_t ks_seqset_chunks(const ks_seqset *s) { return s ? (s->total - 16) / 16 : 0; }

static int count_range_impl(ks_ctx *ctx, const ks_seqset *s, int k, int64_t chunk0, int64_t nchunks,
                            int32_t *d_counts, double *n_words, uint64_t *d_nwords, bool sync) {
  if (!ctx) return KS_ERR_ARG;
  if (!s || !d_counts) return ctx->fail(KS_ERR_ARG, "ks_dev_count_range: null argument");
  int rc = check_k(ctx, k);
  if (rc) return rc;
  const int64_t total = (s->total - 16) / 16;
  if (chunk0 < 0 || nchunks < 0 || chunk0 + nchunks > total) return ctx->fail(KS_ERR_ARG, "chunk range outside the buffer");
  if (s->window && nchunks && (16 * chunk0 < s->win_lo || 16 * (chunk0 + nchunks) + 16 > s->win_hi))
    return ctx->fail(KS_ERR_ARG, "chunk range outside the resident window of the set");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  size_t n = (size_t)1 << (2 * k);
  CountRun run;
  rc = count_begin(ctx, k, d_counts, nchunks, &run);
  if (rc) return rc;
  cudaEvent_t pe = ctx->prof_begin();
  rc = count_chunks(ctx, s, run, chunk0, nchunks);
  if (!rc) rc = count_end(ctx, s, run, chunk0, nchunks);
  if (rc) return rc;
  if (!s->packed) {  // the scan of every shard reads packed data beyond its own range: pack the rest of what is resident
    const int64_t r0 = s->win_lo / 16, r1 = std::max<int64_t>(r0, (s->win_hi - 16) / 16);
    const int64_t a1 = std::min<int64_t>(chunk0, r1), b0 = std::max<int64_t>(chunk0 + nchunks, r0);
    if (a1 > r0) {
      pack_count_kernel<false><<<grid_for((size_t)(a1 - r0), 256, 148u * 8u), 256, 0, st>>>(
          s->d_buf, r0, a1 - r0, k, (uint32_t)(n - 1), s->d_pk, s->d_brk, nullptr, nullptr);
      LAUNCHED(1);
    }
    if (r1 > b0) {
      pack_count_kernel<false><<<grid_for((size_t)(r1 - b0), 256, 148u * 8u), 256, 0, st>>>(
          s->d_buf, b0, r1 - b0, k, (uint32_t)(n - 1), s->d_pk, s->d_brk, nullptr, nullptr);
      LAUNCHED(1);
    }
  }
  ctx->prof_end(KS_PROF_COUNT, pe);
  CK(cudaGetLastError());
  s->packed = true;
  if (d_nwords) CK(cudaMemcpyAsync(d_nwords, ctx->nwords.p, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  if (!sync) return KS_OK;
  unsigned long long nw = 0;
  CK(cudaMemcpyAsync(&nw, ctx->nwords.p, sizeof nw, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (n_words) *n_words = (double)nw;
  return KS_OK;
}
int ks_dev_count_range(ks_ctx *ctx, const ks_seqset *s, int k, int64_t chunk0, int64_t nchunks,
                       int32_t *d_counts, double *n_words) {
  KS_TRY
  return count_range_impl(ctx, s, k, chunk0, nchunks, d_counts, n_words, nullptr, true);
  KS_CATCH(ctx)
}
int ks_dev_count_range_async(ks_ctx *ctx, const ks_seqset *s, int k, int64_t chunk0, int64_t nchunks,
                             int32_t *d_counts, uint64_t *d_nwords) {
  KS_TRY
  return count_range_impl(ctx, s, k, chunk0, nchunks, d_counts, nullptr, d_nwords, false);
  KS_CATCH(ctx)
}

static int ensure_hpin(ks_ctx *ctx) {
  if (ctx->hpin) return KS_OK;
  CK(cudaMallocHost((void **)&ctx->hpin, ks_ctx::HPIN_BYTES));
  return KS_OK;
}

// ------------------------------------------------------------------------------------------------
// stage: score tables
// The pieces of the rank order addressed by absolute position, for the scan's 4-byte gather (rank_value,
// ks_kernels.cuh): piece starts, the shared-memory image (bucket table + the window of pieces that covers the most
// positions), and the scalars the kernels need.  The vectors in `rp` must stay alive until the stream is synchronised.
namespace {
struct RankPosHost {
  std::vector<uint32_t> p0;
  std::vector<RankSmem> blob;
};
}  // namespace
static int rank_positions_setup(ks_ctx *ctx, int k, size_t n, size_t ng, const std::vector<uint64_t> &gstart,
                                const std::vector<uint32_t> &seg_first, const std::vector<unsigned long long> &j0,
                                const std::vector<double> &x0, const std::vector<double> &inc, RankPosHost &rp) {
  cudaStream_t st = ctx->stream;
  const size_t nsg = j0.size();
  std::vector<uint32_t> &p0 = rp.p0;
  rp.blob.resize(1);
  p0.resize(nsg + 1);
  for (size_t g = 0; g < ng; ++g)
    for (uint32_t i = seg_first[g]; i < seg_first[g + 1]; ++i) p0[i] = (uint32_t)(gstart[g] + j0[i]);
  p0[nsg] = (uint32_t)n;
  // window of at most RK_SMEM_PIECES consecutive pieces that covers the most positions of the rank order
  size_t w0 = 0, w1 = std::min<size_t>(nsg, RK_SMEM_PIECES);
  {
    uint64_t best = (uint64_t)p0[w1] - p0[0];
    for (size_t i = 1; i + RK_SMEM_PIECES <= nsg; ++i) {
      const uint64_t cover = (uint64_t)p0[i + RK_SMEM_PIECES] - p0[i];
      if (cover > best) { best = cover; w0 = i; w1 = i + RK_SMEM_PIECES; }
    }
  }
  const uint32_t win_lo = p0[w0], win_len = p0[w1] - p0[w0];
  int shift = 0;
  while (((uint64_t)RK_BUCKETS << shift) < win_len) ++shift;
  RankSmem &im = rp.blob[0];
  memset(&im, 0, sizeof im);
  {
    size_t a = w0;
    for (size_t bkt = 0; bkt <= (size_t)RK_BUCKETS; ++bkt) {
      const uint64_t pos = (uint64_t)win_lo + ((uint64_t)bkt << shift);
      while (a + 1 < w1 && p0[a + 1] <= pos) ++a;
      im.bucket[bkt] = (uint32_t)(a - w0);
    }
  }
  for (size_t i = 0; i <= (size_t)RK_SMEM_PIECES; ++i) im.p0[i] = w0 + i <= w1 ? p0[w0 + i] : 0xffffffffu;
  for (size_t i = 0; i < (size_t)RK_SMEM_PIECES && w0 + i < w1; ++i) { im.x0[i] = x0[w0 + i]; im.inc[i] = inc[w0 + i]; }
  CK(ctx->rk_pos.ensure(n * 4));
  CK(ctx->rk_p0.ensure((nsg + 1) * 4));
  CK(ctx->rk_blob.ensure(sizeof(RankSmem)));
  CK(cudaMemcpyAsync(ctx->rk_p0.p, p0.data(), (nsg + 1) * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->rk_blob.p, &im, sizeof im, cudaMemcpyHostToDevice, st));
  ctx->rk_shift = shift;
  ctx->rk_k = k;
  ctx->rk_npieces = (uint32_t)nsg;
  ctx->rk_win_lo = win_lo;
  ctx->rk_win_len = win_len;
  ctx->rk_n = (uint32_t)n;
  ctx->rk_max = fma((double)(n - 1 - p0[nsg - 1]), inc[nsg - 1], x0[nsg - 1]);  // rank of the last k-mer in order
  return KS_OK;
}

static int dev_scores_impl(ks_ctx *ctx, int k, const int32_t *d_counts, double total, bool total_on_device,
                           int mode, double param, double *d_scores, double *total_out);
int ks_dev_scores(ks_ctx *ctx, int k, const int32_t *d_counts, double total, int mode, double param,
                  double *d_scores) {
  KS_TRY
  return dev_scores_impl(ctx, k, d_counts, total, false, mode, param, d_scores, nullptr);
  KS_CATCH(ctx)
}
int ks_dev_scores_devtotal(ks_ctx *ctx, int k, const int32_t *d_counts, const uint64_t *d_total, int mode,
                           double param, double *d_scores, double *total_out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!d_total) return ctx->fail(KS_ERR_ARG, "ks_dev_scores_devtotal: null argument");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->nwords.ensure(sizeof(unsigned long long)));
  CK(cudaMemcpyAsync(ctx->nwords.p, d_total, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return dev_scores_impl(ctx, k, d_counts, 0.0, true, mode, param, d_scores, total_out);
  KS_CATCH(ctx)
}
// total_on_device: the number of words is still in ctx->nwords (the count pass was not synchronised);
// it is read back together with the histogram
static int dev_scores_impl(ks_ctx *ctx, int k, const int32_t *d_counts, double total, bool total_on_device,
                           int mode, double param, double *d_scores, double *total_out) {
  if (!ctx) return KS_ERR_ARG;
  {  // a table still being written next to an earlier scan reads the staging areas this call is about to refill
    const int rcj = join_table(ctx);
    if (rcj) return rcj;
  }
  if (total_on_device && mode != KS_MODE_LOG2 && mode != KS_MODE_SIGN) {
    unsigned long long nw = 0;
    CK(cudaMemcpyAsync(&nw, ctx->nwords.p, sizeof nw, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    total = (double)nw;
    total_on_device = false;
  }
  if (total_out && !total_on_device) *total_out = total;
  const bool count_fn_mode = (mode == KS_MODE_LOG2 || mode == KS_MODE_SIGN);
  if (!d_counts || (!d_scores && !count_fn_mode)) return ctx->fail(KS_ERR_ARG, "ks_dev_scores: null argument");
  int rc = check_k(ctx, k);
  if (rc) return rc;
  if (mode < KS_MODE_RANK || mode > KS_MODE_RANK_REL) return ctx->fail(KS_ERR_ARG, "unknown score mode %d", mode);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)1 << (2 * k);
  const bool rank_mode = (mode == KS_MODE_RANK || mode == KS_MODE_RANK_REL);
  struct ProfScope {
    ks_ctx *c; cudaEvent_t a;
    ~ProfScope() { c->prof_end(KS_PROF_SCORES, a); }
  } prof_scope{ctx, ctx->prof_begin()};
  if (!rank_mode) {
    // Count-function modes need no per-k-mer order: frequency-of-counts histogram + prefix sum.
    const uint32_t DENSE = 1u << 16;
    if (((uintptr_t)d_counts & 15u) != 0)
      return ctx->fail(KS_ERR_ARG, "ks_dev_scores: the count table must be 16-byte aligned");
    CK(ctx->sc_small.ensure(64));
    CK(ctx->foc_hist.ensure((size_t)DENSE * 4 + 8192 * 8));  // bins + the compact (count, multiplicity) list
    uint32_t big_cap = 1u << 16;
    uint32_t npairs = 0;
    std::vector<uint32_t> full_hist;
    rc = ensure_hpin(ctx);
    if (rc) return rc;
    uint32_t *hist = reinterpret_cast<uint32_t *>(ctx->hpin + ks_ctx::HPIN_HIST);
    uint32_t *hsmall = reinterpret_cast<uint32_t *>(ctx->hpin + ks_ctx::HPIN_SMALL);
    std::vector<uint32_t> big;
    for (int attempt = 0; attempt < 2; ++attempt) {
      CK(ctx->foc_big.ensure((size_t)big_cap * 4));
      CK(cudaMemsetAsync(ctx->sc_small.p, 0, 64, st));
      CK(cudaMemsetAsync(ctx->foc_hist.p, 0, (size_t)DENSE * 4, st));
      foc_hist_kernel<<<grid_for(n, 256, 148u * 4u), 256, 0, st>>>(
          reinterpret_cast<const uint32_t *>(d_counts), n, ctx->foc_hist.as<uint32_t>(), DENSE,
          ctx->foc_big.as<uint32_t>(), ctx->sc_small.as<uint32_t>(), big_cap);
      LAUNCHED(1);
      CK(cudaGetLastError());
      // the non-empty bins as (count, multiplicity) pairs: a few hundred for a genome, against 65 536 bins
      const uint32_t PAIR_CAP = 8192;
      uint2 *d_pairs = reinterpret_cast<uint2 *>(ctx->foc_hist.as<uint32_t>() + DENSE);
      foc_compact_kernel<<<64, 256, 0, st>>>(ctx->foc_hist.as<uint32_t>(), DENSE, d_pairs,
                                            ctx->sc_small.as<uint32_t>() + 4, PAIR_CAP);
      LAUNCHED(1);
      CK(cudaMemcpyAsync(hsmall, ctx->sc_small.p, 32, cudaMemcpyDeviceToHost, st));
      if (total_on_device) CK(cudaMemcpyAsync(hsmall + 8, ctx->nwords.p, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(hist, d_pairs, (size_t)PAIR_CAP * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const uint32_t nbig = hsmall[0];
      npairs = hsmall[4];
      if (total_on_device) {
        unsigned long long nw;
        memcpy(&nw, hsmall + 8, 8);
        total = (double)nw;
        total_on_device = false;
        if (total_out) *total_out = total;
      }
      if (nbig > big_cap) { big_cap = nbig; continue; }
      if (npairs > PAIR_CAP) {  // more distinct counts than the compact list holds: read the bins themselves
        full_hist.resize(DENSE);
        CK(cudaMemcpy(full_hist.data(), ctx->foc_hist.p, (size_t)DENSE * 4, cudaMemcpyDeviceToHost));
      }
      big.resize(nbig);
      if (nbig) CK(cudaMemcpy(big.data(), ctx->foc_big.p, (size_t)nbig * 4, cudaMemcpyDeviceToHost));
      break;
    }
    std::sort(big.begin(), big.end());
    // distinct counts ascending with multiplicities
    std::vector<uint32_t> gcount;
    std::vector<uint64_t> gmult;
    if (!full_hist.empty()) {
      for (uint32_t c = 0; c < DENSE; ++c)
        if (full_hist[c]) { gcount.push_back(c); gmult.push_back(full_hist[c]); }
    } else {
      std::vector<std::pair<uint32_t, uint32_t>> pr(npairs);
      for (uint32_t i = 0; i < npairs; ++i) pr[i] = {hist[2 * i], hist[2 * i + 1]};
      std::sort(pr.begin(), pr.end());
      for (auto &q : pr) { gcount.push_back(q.first); gmult.push_back(q.second); }
    }
    for (size_t i = 0; i < big.size();) {
      size_t j = i;
      while (j < big.size() && big[j] == big[i]) ++j;
      gcount.push_back(big[i]);
      gmult.push_back(j - i);
      i = j;
    }
    const size_t ng = gcount.size();
    auto count_at = [&](uint64_t pos) -> double {  // count at position pos of the ascending order
      uint64_t acc = 0;
      for (size_t g = 0; g < ng; ++g) { acc += gmult[g]; if (pos < acc) return (double)(int32_t)gcount[g]; }
      return ng ? (double)(int32_t)gcount[ng - 1] : 0.0;
    };
    double f_lo = count_at(n / 2 - 1) / total, f_hi = count_at(n / 2) / total;
    double f_med = (f_lo + f_hi) / 2.0;
    std::vector<double> lut(ng);
    for (size_t g = 0; g < ng; ++g) {
      double f = (double)(int32_t)gcount[g] / total;
      if (mode == KS_MODE_LOG2) lut[g] = log2(f / f_med);
      else {
        double f_t = isfinite(param) ? param : f_med;
        lut[g] = f >= f_t ? 1.0 : -1.0;
      }
    }
    ctx->lut_gcount = gcount;
    ctx->lut_gval = lut;
    ctx->lut_valid = true;
    ctx->lut_k = k;
    // class table for the scan: 2 bytes per k-mer instead of the 4-byte count (stays L2 resident at k = 12)
    ctx->cls_counts = nullptr;
    ctx->core_valid = false;
    bool gcount_on_device = false;
    if (ng && ng <= 65535 && getenv("KS_NO_CLASS_TABLE") == nullptr) {
      uint32_t ndense = std::min<uint32_t>(gcount[ng - 1] + 1, DENSE);
      // staged in pinned memory: every earlier copy out of it completed before the synchronisation above
      uint16_t *dense = reinterpret_cast<uint16_t *>(ctx->hpin + ks_ctx::HPIN_CLS);
      uint32_t *gc_pin = reinterpret_cast<uint32_t *>(ctx->hpin + ks_ctx::HPIN_GCOUNT);
      memset(dense, 0, (size_t)ndense * 2);
      for (size_t g = 0; g < ng && gcount[g] < ndense; ++g) dense[gcount[g]] = (uint16_t)g;
      memcpy(gc_pin, gcount.data(), ng * 4);
      CK(ctx->cls.ensure(n * 2));
      CK(ctx->cls_dense.ensure((size_t)ndense * 2));
      CK(ctx->sc_gcount.ensure((ng + 1) * 4));
      CK(cudaMemcpyAsync(ctx->cls_dense.p, dense, (size_t)ndense * 2, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->sc_gcount.p, gc_pin, ng * 4, cudaMemcpyHostToDevice, st));
      gcount_on_device = true;
      class_apply_kernel<<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(d_counts), n,
                                                          ctx->cls_dense.as<uint16_t>(), ndense,
                                                          ctx->sc_gcount.as<uint32_t>(), (uint32_t)ng,
                                                          ctx->cls.as<uint16_t>());
      LAUNCHED(1);
      CK(cudaGetLastError());
      ctx->cls_counts = d_counts;
      ctx->cls_n = n;
      // core records: one 8-byte gather serves two consecutive positions of the scan
      ctx->core_valid = false;
      if (getenv("KS_NO_CORE_TABLE") == nullptr) {
        // a record names a class in one byte: the 255 groups that cover the most POSITIONS (count x number of
        // k-mers with that count) get a byte of their own, whatever their counts; the rest escape to cls[]
        std::vector<uint32_t> order(ng);
        for (size_t g = 0; g < ng; ++g) order[g] = (uint32_t)g;
        const size_t ncc = std::min<size_t>(ng, CORE_ESCAPE);
        if (getenv("KS_CORE_BY_VALUE") == nullptr)
          std::partial_sort(order.begin(), order.begin() + ncc, order.end(), [&](uint32_t a, uint32_t b) {
            const unsigned __int128 wa = (unsigned __int128)gmult[a] * gcount[a], wb = (unsigned __int128)gmult[b] * gcount[b];
            return wa != wb ? wa > wb : a < b;
          });
        ctx->core_groups.assign(order.begin(), order.begin() + ncc);
        uint8_t *cc = reinterpret_cast<uint8_t *>(ctx->hpin + ks_ctx::HPIN_CC);
        memset(cc, (int)CORE_ESCAPE, ng);
        for (size_t i = 0; i < ncc; ++i) cc[ctx->core_groups[i]] = (uint8_t)i;
        CK(ctx->core_cc.ensure(ng + 16));
        CK(cudaMemcpyAsync(ctx->core_cc.p, cc, ng, cudaMemcpyHostToDevice, st));
        const size_t ncore = n / 4;
        CK(ctx->core.ensure(ncore * sizeof(uint2)));
        core_apply_kernel<<<grid_for(ncore, 256), 256, 0, st>>>(ctx->cls.as<uint16_t>(), ncore, ctx->core.as<uint2>(),
                                                               ctx->core_cc.as<uint8_t>());
        LAUNCHED(1);
        CK(cudaGetLastError());
        ctx->core_valid = true;
      }
    }
    if (d_scores) {  // the per-k-mer table itself (an output; the scan gathers counts + LUT instead)
      uint32_t ndense = ng ? std::min<uint32_t>(gcount[ng - 1] + 1, DENSE) : 0;
      const bool staged = ng <= 65536;  // pinned staging: no synchronisation needed behind the copies
      std::vector<double> dense_v;
      double *dense = reinterpret_cast<double *>(ctx->hpin + ks_ctx::HPIN_DENSE);
      double *lut_src = reinterpret_cast<double *>(ctx->hpin + ks_ctx::HPIN_LUT);
      uint32_t *gc_src = reinterpret_cast<uint32_t *>(ctx->hpin + ks_ctx::HPIN_GCOUNT);
      if (!staged) {
        dense_v.assign(ndense ? ndense : 1, 0.0);
        dense = dense_v.data();
        lut_src = lut.data();
        gc_src = gcount.data();
      } else {
        memcpy(lut_src, lut.data(), ng * 8);
        memcpy(gc_src, gcount.data(), ng * 4);
      }
      const size_t dense_n = ndense ? ndense : 1;
      memset(dense, 0, dense_n * 8);
      for (size_t g = 0; g < ng && gcount[g] < ndense; ++g) dense[gcount[g]] = lut[g];
      CK(ctx->sc_gcount.ensure((ng + 1) * 4));
      CK(ctx->sc_lut.ensure(ng * 8 + 8));
      CK(ctx->sc_dense.ensure(dense_n * 8));
      // inside ks_dev_pipeline the table is written next to the scan, which reads counts and classes, not the table
      const bool side = (ctx->defer_table || ctx->side_table_opt) && staged && gcount_on_device;
      cudaStream_t ts = st;
      if (side) {
        int rc2 = ensure_copy_stream(ctx);
        if (rc2) return rc2;
        ts = ctx->copy_stream;
        CK(cudaEventRecord(ctx->ev_compute, st));  // counts final, sc_gcount uploaded
        CK(cudaStreamWaitEvent(ts, ctx->ev_compute, 0));
      }
      if (!side) CK(cudaMemcpyAsync(ctx->sc_gcount.p, gc_src, ng * 4, cudaMemcpyHostToDevice, ts));
      CK(cudaMemcpyAsync(ctx->sc_lut.p, lut_src, ng * 8, cudaMemcpyHostToDevice, ts));
      CK(cudaMemcpyAsync(ctx->sc_dense.p, dense, dense_n * 8, cudaMemcpyHostToDevice, ts));
      lut_apply_kernel<<<blocks_exact(n, 256), 256, 0, ts>>>(reinterpret_cast<const uint32_t *>(d_counts), n,
                                                             ctx->sc_dense.as<double>(), ndense,
                                                             ctx->sc_gcount.as<uint32_t>(), (uint32_t)ng,
                                                             ctx->sc_lut.as<double>(), d_scores);
      LAUNCHED(1);
      CK(cudaGetLastError());
      if (side) {
        if (!ctx->ev_table) CK(cudaEventCreateWithFlags(&ctx->ev_table, cudaEventDisableTiming));
        CK(cudaEventRecord(ctx->ev_table, ts));
        ctx->table_pending = true;
      }
      if (!staged) CK(cudaStreamSynchronize(st));
    }
    return KS_OK;
  }
  ctx->rk_valid = false;
  if (rank_mode && total == 0) {
    // 0/0 addends: every rank but the first in sort order (k-mer 0) is NaN (:200)
    fill_nan_kernel<<<blocks_exact(n, 256), 256, 0, st>>>(d_scores, n);
    LAUNCHED(1);
    CK(cudaGetLastError());
  } else {
    // 1. stable (count, index) order
    CK(ctx->sc_keys_a.ensure(n * 4));
    CK(ctx->sc_keys_b.ensure(n * 4));
    CK(ctx->sc_vals_a.ensure(n * 4));
    CK(ctx->sc_vals_b.ensure(n * 4));
    size_t nb = radix_nblocks(n);
    CK(ctx->sort_hist.ensure((256 * nb + 2) * 4));
    CK(ctx->sort_scan.ensure(exclusive_scan_scratch_elems(256 * nb) * 4));
    CK(ctx->sc_small.ensure(64));
    CK(cudaMemsetAsync(ctx->sc_small.p, 0, 64, st));
    uint32_t *d_max = ctx->sc_small.as<uint32_t>();
    uint32_t *d_ngroups = d_max + 1;
    CK(cudaMemcpyAsync(ctx->sc_keys_a.p, d_counts, n * 4, cudaMemcpyDeviceToDevice, st));
    max_u32_kernel<<<grid_for(n, 256), 256, 0, st>>>(ctx->sc_keys_a.as<uint32_t>(), n, d_max);
    LAUNCHED(1);
    uint32_t maxc = 0;
    CK(cudaMemcpyAsync(&maxc, d_max, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int nbits = 0;
    while (nbits < 32 && (maxc >> nbits) != 0) ++nbits;
    RadixScratch rs{ctx->sort_hist.as<uint32_t>(), ctx->sort_scan.as<uint32_t>()};
    uint32_t *skeys = nullptr, *svals = nullptr;
    LAUNCHED(radix_sort_pairs<uint32_t>(ctx->sc_keys_a.as<uint32_t>(), ctx->sc_vals_a.as<uint32_t>(),
                                        ctx->sc_keys_b.as<uint32_t>(), ctx->sc_vals_b.as<uint32_t>(), n, nbits,
                                        true, rs, st, &skeys, &svals));
    CK(cudaGetLastError());
    // 2. run-length table of the sorted counts (#distinct counts <= sqrt(2 * total) + 1)
    size_t gcap = (size_t)(sqrt(2.0 * (total > 0 ? total : 1.0)) + 16.0);
    if (gcap > n) gcap = n;
    if (gcap < 16) gcap = 16;
    std::vector<uint32_t> gcount, gstart32;
    for (int attempt = 0; attempt < 2; ++attempt) {
      CK(ctx->sc_gcount.ensure((gcap + 1) * 4));
      CK(ctx->sc_gstart.ensure((gcap + 1) * 4));
      CK(cudaMemsetAsync(d_ngroups, 0, 4, st));
      rle_heads_kernel<<<blocks_exact(n, 256), 256, 0, st>>>(skeys, n, ctx->sc_gcount.as<uint32_t>(),
                                                             ctx->sc_gstart.as<uint32_t>(), d_ngroups,
                                                             (uint32_t)gcap);
      LAUNCHED(1);
      uint32_t ng = 0;
      CK(cudaMemcpyAsync(&ng, d_ngroups, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (ng > gcap) { gcap = ng; continue; }
      gcount.resize(ng);
      gstart32.resize(ng);
      CK(cudaMemcpyAsync(gcount.data(), ctx->sc_gcount.p, (size_t)ng * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(gstart32.data(), ctx->sc_gstart.p, (size_t)ng * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      break;
    }
    const size_t ng = gcount.size();
    {  // appended in arbitrary order: order the (few) groups by start position
      std::vector<uint32_t> ord(ng);
      for (size_t i = 0; i < ng; ++i) ord[i] = (uint32_t)i;
      std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return gstart32[a] < gstart32[b]; });
      std::vector<uint32_t> c2(ng), s2(ng);
      for (size_t i = 0; i < ng; ++i) { c2[i] = gcount[ord[i]]; s2[i] = gstart32[ord[i]]; }
      gcount.swap(c2);
      gstart32.swap(s2);
    }
    std::vector<uint64_t> gstart(ng + 1);
    for (size_t i = 0; i < ng; ++i) gstart[i] = gstart32[i];
    gstart[ng] = n;
    {
      // 3a. linear pieces of the sequential accumulation (ks_rankseg.h), evaluated on the device
      std::vector<uint32_t> seg_first;
      std::vector<RankSeg> segs;
      build_rank_segments(gcount.data(), gstart.data(), ng, total, seg_first, segs);
      size_t nsg = segs.size();
      std::vector<unsigned long long> j0(nsg);
      std::vector<double> x0(nsg), inc(nsg);
      for (size_t i = 0; i < nsg; ++i) { j0[i] = segs[i].j0; x0[i] = segs[i].x0; inc[i] = segs[i].inc; }
      gstart32.push_back((uint32_t)n);
      CK(ctx->sc_gstart.ensure((ng + 1) * 4));
      CK(ctx->sc_segfirst.ensure((ng + 1) * 4));
      CK(ctx->sc_segj0.ensure(nsg * 8 + 8));
      CK(ctx->sc_segx0.ensure(nsg * 8 + 8));
      CK(ctx->sc_seginc.ensure(nsg * 8 + 8));
      CK(cudaMemcpyAsync(ctx->sc_gstart.p, gstart32.data(), (ng + 1) * 4, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->sc_segfirst.p, seg_first.data(), (ng + 1) * 4, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->sc_segj0.p, j0.data(), nsg * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->sc_segx0.p, x0.data(), nsg * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->sc_seginc.p, inc.data(), nsg * 8, cudaMemcpyHostToDevice, st));
      // the same pieces addressed by absolute position in the rank order, for the scan's 4-byte gather
      const bool want_pos = mode == KS_MODE_RANK && nsg > 0 && n < 0xffffffffull && getenv("KS_NO_RANK_POS") == nullptr;
      RankPosHost rp;  // host images: they must outlive the copies (synchronised below)
      if (want_pos) {
        rc = rank_positions_setup(ctx, k, n, ng, gstart, seg_first, j0, x0, inc, rp);
        if (rc) return rc;
      }
      if (want_pos) {
        // positions first (the one random scatter, 4 bytes), then the rank table in index order: coalesced
        rank_pos_scatter_kernel<<<blocks_exact(n, 256), 256, 0, st>>>(svals, n, ctx->rk_pos.as<uint32_t>());
        RankPieces R;
        R.p0 = ctx->rk_p0.as<uint32_t>(); R.x0 = ctx->sc_segx0.as<double>(); R.inc = ctx->sc_seginc.as<double>();
        R.npieces = ctx->rk_npieces; R.win_lo = ctx->rk_win_lo; R.win_len = ctx->rk_win_len; R.shift = ctx->rk_shift;
        rank_from_pos_kernel<<<grid_for(n, 256, 148u * 8u), 256, 0, st>>>(ctx->rk_pos.as<uint32_t>(), n, ctx->rk_blob.p, R,
                                                                         d_scores);
        LAUNCHED(1);
      } else {
        rank_eval_kernel<<<blocks_exact(n, 256), 256, 0, st>>>(
            svals, n, ctx->sc_gstart.as<uint32_t>(), (uint32_t)ng, ctx->sc_segfirst.as<uint32_t>(),
            ctx->sc_segj0.as<unsigned long long>(), ctx->sc_segx0.as<double>(), ctx->sc_seginc.as<double>(),
            d_scores, nullptr);
      }
      LAUNCHED(1);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(st));  // host vectors above must outlive the copies
      ctx->rk_valid = want_pos;
    }
  }
  if (mode == KS_MODE_RANK_REL) {
    affine_kernel<<<blocks_exact(n, 256), 256, 0, st>>>(d_scores, n, param, param);
    LAUNCHED(1);
    CK(cudaGetLastError());
  }
  return KS_OK;
}

// ------------------------------------------------------------------------------------------------
// Rank-mode score stage SLICED over several devices (multi-GPU): the summed count table is the
// same on every device, so device `slice` of `nslices` derives the rank order of its own slice of the k-mer index
// space only -- stable sort of the slice by count, run-length table of the slice -- and the slices' run-length tables
// (a few thousand entries) are exchanged through `fn`.  With them every device knows, for each distinct count, how many
// k-mers hold it in all slices before its own: position in the global (count, index) order = first position of the
// count's group + that number + the ordinal inside the slice.  The linear pieces of ks_rankseg.h are derived on every
// device from the merged table (identical), ranks and rank-order positions are written for the slice
// [slice * n / nslices, (slice + 1) * n / nslices) only; the caller all-gathers both tables.
static int rank_scores_sliced(ks_ctx *ctx, int k, const int32_t *d_counts, double total, int slice, int nslices,
                              ks_gather_fn fn, void *user, double *d_scores) {
  if (!ctx) return KS_ERR_ARG;
  if (!d_counts || !d_scores || !fn || nslices < 1 || slice < 0 || slice >= nslices)
    return ctx->fail(KS_ERR_ARG, "ks_dev_scores_rank_sliced: bad arguments");
  int rc = check_k(ctx, k);
  if (rc) return rc;
  if (!(total > 0)) return ctx->fail(KS_ERR_ARG, "ks_dev_scores_rank_sliced: no k-mers counted");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)1 << (2 * k);
  const size_t lo = n * (size_t)slice / (size_t)nslices, hi = n * (size_t)(slice + 1) / (size_t)nslices;
  const size_t m = hi - lo;
  ctx->rk_valid = false;
  struct ProfScope {
    ks_ctx *c; cudaEvent_t a;
    ~ProfScope() { c->prof_end(KS_PROF_SCORES, a); }
  } prof_scope{ctx, ctx->prof_begin()};
  // 1. stable (count, index) order of the slice
  std::vector<uint32_t> lcount, lstart;  // local run-length table: count, first local sorted position
  uint32_t *skeys = nullptr, *svals = nullptr;
  if (m) {
    CK(ctx->sc_keys_a.ensure(m * 4));
    CK(ctx->sc_keys_b.ensure(m * 4));
    CK(ctx->sc_vals_a.ensure(m * 4));
    CK(ctx->sc_vals_b.ensure(m * 4));
    size_t nb = radix_nblocks(m);
    CK(ctx->sort_hist.ensure((256 * nb + 2) * 4));
    CK(ctx->sort_scan.ensure(exclusive_scan_scratch_elems(256 * nb) * 4));
    CK(ctx->sc_small.ensure(64));
    CK(cudaMemsetAsync(ctx->sc_small.p, 0, 64, st));
    uint32_t *d_max = ctx->sc_small.as<uint32_t>();
    uint32_t *d_ngroups = d_max + 1;
    CK(cudaMemcpyAsync(ctx->sc_keys_a.p, d_counts + lo, m * 4, cudaMemcpyDeviceToDevice, st));
    max_u32_kernel<<<grid_for(m, 256), 256, 0, st>>>(ctx->sc_keys_a.as<uint32_t>(), m, d_max);
    LAUNCHED(1);
    uint32_t maxc = 0;
    CK(cudaMemcpyAsync(&maxc, d_max, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int nbits = 0;
    while (nbits < 32 && (maxc >> nbits) != 0) ++nbits;
    RadixScratch rs{ctx->sort_hist.as<uint32_t>(), ctx->sort_scan.as<uint32_t>()};
    LAUNCHED(radix_sort_pairs<uint32_t>(ctx->sc_keys_a.as<uint32_t>(), ctx->sc_vals_a.as<uint32_t>(),
                                        ctx->sc_keys_b.as<uint32_t>(), ctx->sc_vals_b.as<uint32_t>(), m, nbits, true,
                                        rs, st, &skeys, &svals));
    CK(cudaGetLastError());
    size_t gcap = (size_t)(sqrt(2.0 * total) + 16.0);
    if (gcap > m) gcap = m;
    if (gcap < 16) gcap = 16;
    for (int attempt = 0; attempt < 2; ++attempt) {
      CK(ctx->sc_gcount.ensure((gcap + 1) * 4));
      CK(ctx->sc_gstart.ensure((gcap + 1) * 4));
      CK(cudaMemsetAsync(d_ngroups, 0, 4, st));
      rle_heads_kernel<<<blocks_exact(m, 256), 256, 0, st>>>(skeys, m, ctx->sc_gcount.as<uint32_t>(),
                                                             ctx->sc_gstart.as<uint32_t>(), d_ngroups, (uint32_t)gcap);
      LAUNCHED(1);
      uint32_t ngl = 0;
      CK(cudaMemcpyAsync(&ngl, d_ngroups, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (ngl > gcap) { gcap = ngl; continue; }
      lcount.resize(ngl);
      lstart.resize(ngl);
      CK(cudaMemcpyAsync(lcount.data(), ctx->sc_gcount.p, (size_t)ngl * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(lstart.data(), ctx->sc_gstart.p, (size_t)ngl * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      break;
    }
    std::vector<uint32_t> ord(lcount.size());
    for (size_t i = 0; i < ord.size(); ++i) ord[i] = (uint32_t)i;
    std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return lstart[a] < lstart[b]; });
    std::vector<uint32_t> c2(ord.size()), s2(ord.size());
    for (size_t i = 0; i < ord.size(); ++i) { c2[i] = lcount[ord[i]]; s2[i] = lstart[ord[i]]; }
    lcount.swap(c2);
    lstart.swap(s2);
  }
  const size_t ngl = lcount.size();
  // 2. exchange the run-length tables: sizes first, then (count, multiplicity) pairs padded to the longest
  std::vector<uint64_t> sizes((size_t)nslices, 0);
  {
    uint64_t mine = ngl;
    if (fn(user, &mine, sizeof mine, sizes.data())) return ctx->fail(KS_ERR_ARG, "rank exchange (sizes) failed");
  }
  size_t maxg = 0;
  for (uint64_t v : sizes) maxg = std::max<size_t>(maxg, (size_t)v);
  if (maxg == 0) return ctx->fail(KS_ERR_ARG, "ks_dev_scores_rank_sliced: empty table");
  std::vector<uint32_t> mine_pairs(2 * maxg, 0), all_pairs(2 * maxg * (size_t)nslices, 0);
  for (size_t g = 0; g < ngl; ++g) {
    mine_pairs[2 * g] = lcount[g];
    mine_pairs[2 * g + 1] = (uint32_t)((g + 1 < ngl ? lstart[g + 1] : (uint32_t)m) - lstart[g]);
  }
  if (fn(user, mine_pairs.data(), mine_pairs.size() * 4, all_pairs.data()))
    return ctx->fail(KS_ERR_ARG, "rank exchange (tables) failed");
  // 3. merged table: distinct counts ascending, total multiplicity, multiplicity in the slices before this one
  std::vector<uint32_t> gcount;
  {
    std::vector<uint32_t> allc;
    for (int sidx = 0; sidx < nslices; ++sidx)
      for (size_t g = 0; g < (size_t)sizes[sidx]; ++g) allc.push_back(all_pairs[2 * (maxg * sidx + g)]);
    std::sort(allc.begin(), allc.end());
    allc.erase(std::unique(allc.begin(), allc.end()), allc.end());
    gcount.swap(allc);
  }
  const size_t ng = gcount.size();
  std::vector<uint64_t> gmult(ng, 0), before(ng, 0);
  for (int sidx = 0; sidx < nslices; ++sidx)
    for (size_t g = 0; g < (size_t)sizes[sidx]; ++g) {
      const uint32_t c = all_pairs[2 * (maxg * sidx + g)], mu = all_pairs[2 * (maxg * sidx + g) + 1];
      const size_t G = (size_t)(std::lower_bound(gcount.begin(), gcount.end(), c) - gcount.begin());
      gmult[G] += mu;
      if (sidx < slice) before[G] += mu;
    }
  std::vector<uint64_t> gstart(ng + 1, 0);
  for (size_t G = 0; G < ng; ++G) gstart[G + 1] = gstart[G] + gmult[G];
  if (gstart[ng] != n) return ctx->fail(KS_ERR_ARG, "rank exchange: the slices do not add up to 4^k entries");
  std::vector<uint32_t> seg_first;
  std::vector<RankSeg> segs;
  build_rank_segments(gcount.data(), gstart.data(), ng, total, seg_first, segs);
  const size_t nsg = segs.size();
  std::vector<unsigned long long> j0(nsg);
  std::vector<double> x0(nsg), inc(nsg);
  for (size_t i = 0; i < nsg; ++i) { j0[i] = segs[i].j0; x0[i] = segs[i].x0; inc[i] = segs[i].inc; }
  // 4. per local group: where it sits in the global order
  std::vector<uint32_t> lg_first(ngl + 1, 0), lg_seg0(ngl, 0), lg_seg1(ngl, 0);
  std::vector<unsigned long long> lg_j(ngl, 0), lg_p(ngl, 0);
  for (size_t g = 0; g < ngl; ++g) {
    const size_t G = (size_t)(std::lower_bound(gcount.begin(), gcount.end(), lcount[g]) - gcount.begin());
    lg_first[g] = lstart[g];
    lg_j[g] = before[G];                 // ordinal of the slice's first member inside the global group
    lg_p[g] = gstart[G] + before[G];     // its position in the global order
    lg_seg0[g] = seg_first[G];
    lg_seg1[g] = seg_first[G + 1];
  }
  lg_first[ngl] = (uint32_t)m;
  RankPosHost rp;
  const bool want_pos = n < 0xffffffffull;
  if (want_pos) {
    rc = rank_positions_setup(ctx, k, n, ng, gstart, seg_first, j0, x0, inc, rp);
    if (rc) return rc;
  }
  if (m) {
    CK(ctx->sc_gstart.ensure((ngl + 1) * 4));
    CK(ctx->sc_segfirst.ensure((2 * ngl + 2) * 4));
    CK(ctx->sc_gcount.ensure((4 * ngl + 4) * 8));
    CK(ctx->sc_segj0.ensure(nsg * 8 + 8));
    CK(ctx->sc_segx0.ensure(nsg * 8 + 8));
    CK(ctx->sc_seginc.ensure(nsg * 8 + 8));
    unsigned long long *d_lgj = ctx->sc_gcount.as<unsigned long long>();
    unsigned long long *d_lgp = d_lgj + ngl + 1;
    uint32_t *d_seg0 = ctx->sc_segfirst.as<uint32_t>(), *d_seg1 = d_seg0 + ngl + 1;
    CK(cudaMemcpyAsync(ctx->sc_gstart.p, lg_first.data(), (ngl + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_lgj, lg_j.data(), ngl * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_lgp, lg_p.data(), ngl * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_seg0, lg_seg0.data(), ngl * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_seg1, lg_seg1.data(), ngl * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->sc_segj0.p, j0.data(), nsg * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->sc_segx0.p, x0.data(), nsg * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->sc_seginc.p, inc.data(), nsg * 8, cudaMemcpyHostToDevice, st));
    rank_eval_slice_kernel<<<blocks_exact(m, 256), 256, 0, st>>>(
        svals, m, (uint32_t)lo, ctx->sc_gstart.as<uint32_t>(), (uint32_t)ngl, d_lgj, d_lgp, d_seg0, d_seg1,
        ctx->sc_segj0.as<unsigned long long>(), ctx->sc_segx0.as<double>(), ctx->sc_seginc.as<double>(), d_scores,
        want_pos ? ctx->rk_pos.as<uint32_t>() : nullptr);
    LAUNCHED(1);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(st));  // host vectors above must outlive the copies
  ctx->rk_valid = want_pos;
  return KS_OK;
}

int ks_dev_scores_rank_sliced(ks_ctx *ctx, int k, const int32_t *d_counts, double total, int slice, int nslices,
                              ks_gather_fn fn, void *user, double *d_scores) {
  KS_TRY
  return rank_scores_sliced(ctx, k, d_counts, total, slice, nslices, fn, user, d_scores);
  KS_CATCH(ctx)
}

// the rank-order position table (uint32[4^k]) the last rank-mode score stage left on this ctx: after the sliced
// stage only the caller's slice is filled, and the caller gathers the other slices into it
void *ks_ctx_rank_positions(ks_ctx *ctx) { return ctx ? ctx->rk_pos.p : nullptr; }

// ------------------------------------------------------------------------------------------------
// stage: scan + spans
// per-level scratch: stash (indexed by work chunk) and per-tile arrays
static int ensure_tiles(ks_ctx *ctx, size_t tiles, bool lut_mode, bool need_p0, bool need_stash) {
  const size_t Q = tiles * TILE_THREADS;
  if (need_stash) {  // per-position stash: only the position-by-position walk reads it
    if (lut_mode) CK(ctx->st_c.ensure(Q * 16 * 4)); else CK(ctx->st_s.ensure(Q * 16 * 8));
  }
  CK(ctx->st_ea.ensure(Q * 16));
  CK(ctx->st_eb.ensure(Q * 16));
  CK(ctx->st_flags.ensure(Q * 4));
  if (need_p0) CK(ctx->st_p0.ensure(Q * 8));
  const size_t xf_tiles = tiles * TILE_WARPS;  // room for warp tiles (scan_gather_core_kernel)
  CK(ctx->tile_xf.ensure(xf_tiles * sizeof(XfRec)));
  CK(ctx->group_xf.ensure((xf_tiles / 32 + 2) * sizeof(XfRec)));
  CK(ctx->group_S.ensure((xf_tiles / 32 + 2) * 16));
  CK(ctx->group_ex.ensure((tiles / 32 + 2) * sizeof(ExRec)));
  CK(ctx->pending_list.ensure(tiles * 4 + 64));
  CK(ctx->pending_count.ensure(64));
  CK(ctx->tile_ex.ensure(tiles * sizeof(ExRec)));
  CK(ctx->pending.ensure(tiles * sizeof(ExPending)));
  return KS_OK;
}

static int ensure_children(ks_ctx *ctx, size_t cap) {
  if (cap <= ctx->child_cap) return KS_OK;
  CK(ctx->child_pk.ensure(cap * 8));
  CK(ctx->child_c.ensure(cap * 8));
  ctx->child_cap = cap;
  return KS_OK;
}

static int ensure_recs(ks_ctx *ctx, size_t cap) {
  if (cap <= ctx->rec_cap) return KS_OK;
  cudaStream_t st = ctx->stream;
  CK(ctx->rec_beg.ensure(cap * 8, true, st));
  CK(ctx->rec_pk.ensure(cap * 8, true, st));
  CK(ctx->rec_c.ensure(cap * 8, true, st));
  CK(ctx->rec_mhi.ensure(cap * 8, true, st));
  CK(ctx->rec_mlo.ensure(cap * 8, true, st));
  ctx->rec_cap = cap;
  return KS_OK;
}

namespace {
// Level 0 restricted to the dense chunks [chunk0, chunk0 + nchunks) of the buffer (one shard of a
// multi-GPU run).  The carries are exchanged through `fn` on the host between the kernels: what = 0
// hands over the shard's aggregate transform and asks for the state entering the shard, what = 1 the
// shard's open-excursion aggregate and asks for the one entering it (ks_fold_carry computes both from
// the gathered aggregates of all shards).
struct ShardCtl {
  int64_t chunk0 = 0, nchunks = 0;
  ks_exchange_fn fn = nullptr;
  void *user = nullptr;
};
struct ScanTable {  // what scan_gather_kernel gathers from
  bool use_lut = false;
  const uint32_t *counts = nullptr;
  uint32_t lut_size = 0, sp_n = 0;
  bool use_cls = false;  // class mode: ctx->cls + per-class table in ctx->lut_fx
  bool use_core = false; // ... gathered two positions at a time through ctx->core
  bool use_hash = false; // large k: 64-bit codes, scores in the slots of ctx->lg_slots
  bool use_rank = false; // rank mode: ctx->rk_pos + the linear pieces of the rank order
  bool use_rank_core = false;  // ... gathered through the 32-byte records of ctx->rk_core (tables beyond L2)
  double rk_thr = 0;
  bool tr = false;  // transition-score scan: ctx->wfx = [trans | init], every close is re-scanned
};
}  // namespace

// Level loop + ordering + marshaling.  The fixed-point table (ctx->wfx, or ctx->lut_fx + sparse list)
// and the DevScanParams block (ctx->prm) have been prepared on the stream by the caller.
static int scan_core(ks_ctx *ctx, const ks_seqset *s, int k, const ScanTable &tab, uint64_t mw,
                     int32_t *d_inscan, ks_spans *host_out, uint64_t *n_spans, const ShardCtl *sh = nullptr) {
  int rc = KS_OK;
  cudaStream_t st = ctx->stream;
  const size_t nk = (size_t)1 << (2 * k);
  DevScanParams *d_prm = ctx->prm.as<DevScanParams>();
  CK(ctx->rec_count.ensure(64));
  rc = ensure_packed(ctx, s);  // no counting pass ran on this set (user-supplied weights): pack only
  if (rc) return rc;
  unsigned long long *d_rec_count = ctx->rec_count.as<unsigned long long>();
  CK(cudaMemsetAsync(d_rec_count, 0, 2 * sizeof(unsigned long long), st));

  const int64_t dense_chunks = (s->total - 16) / 16;
  if (s->total >= (1ll << 32))
    return ctx->fail(KS_ERR_ARG, "one scan covers at most 2^32 bytes of sequence per GPU (shard the input)");
  rc = ensure_recs(ctx, std::max<size_t>((size_t)1 << 16, (size_t)(dense_chunks / 64)));
  if (rc) return rc;

  if (tab.tr) {
    CK(ctx->child_count.ensure(64));
    rc = ensure_children(ctx, std::max<size_t>((size_t)1 << 16, (size_t)(dense_chunks / 4)));
    if (rc) return rc;
  }
  unsigned long long level_start = 0;  // records before this level
  unsigned long long rec_total = 0;    // records after the last completed level
  int64_t nseg = 0, total_chunks = dense_chunks;
  int64_t dense_chunk0 = 0;
  if (s->window && !sh) return ctx->fail(KS_ERR_ARG, "a window set is scanned by shard (ks_dev_scan*_shard)");
  if (sh) {
    if (sh->chunk0 < 0 || sh->nchunks < 0 || sh->chunk0 + sh->nchunks > dense_chunks)
      return ctx->fail(KS_ERR_ARG, "shard range outside the buffer");
    if (s->window && sh->nchunks && (16 * sh->chunk0 < s->win_lo || 16 * (sh->chunk0 + sh->nchunks) + 16 > s->win_hi))
      return ctx->fail(KS_ERR_ARG, "shard range outside the resident window of the set");
    dense_chunk0 = sh->chunk0;
    total_chunks = sh->nchunks;
    CK(ctx->launch_rec.ensure(256));
  }
  // With min_width >= 15 no excursion inside one 16-position chunk can qualify: the walk then works on
  // per-chunk summaries (scan_walk_fast_kernel) and only a short list of chunks is walked position by
  // position.  A level whose list overflows is redone with the general walk.
  const bool fast_ok = !tab.tr && mw >= 15 && getenv("KS_NO_FAST_WALK") == nullptr;
  bool fast = fast_ok;
  // ... and with min_width >= 31 none inside 32 positions: level 0 then works on units of two chunks
  const bool pair_ok = fast_ok && mw >= 31 && getenv("KS_NO_PAIR") == nullptr;
  bool have_carry = false;  // the exchange runs once, also if level 0 has to be repeated with more room
  fx_t S_carry = 0;
  ExRec E_carry;
  memset(&E_carry, 0, sizeof E_carry);
  int level = 0;
  uint64_t revisit_chunks = 0;
  bool count_inscan = d_inscan != nullptr;
  for (;;) {
    const bool pair = fast && pair_ok && nseg == 0;
    const int64_t records = pair ? (total_chunks + 1) / 2 : total_chunks;
    size_t tiles = (size_t)((records + TILE_THREADS - 1) / TILE_THREADS);
    if (tiles == 0 && !(sh && level == 0)) break;
    if (tiles == 0) tiles = 1;  // an empty shard still takes part in the carry exchange
    if (tiles > 0xfffffff0ull) return ctx->fail(KS_ERR_ARG, "input too large for one scan launch");
    rc = ensure_tiles(ctx, tiles, tab.use_lut, nseg != 0, !fast);
    if (rc) return rc;
    if (tab.tr) CK(ctx->st_aux.ensure(tiles * TILE_THREADS * 4));
    size_t detail_cap = tiles + tiles * TILE_THREADS / 16 + 64;
    if (const char *e = getenv("KS_DETAIL_CAP")) detail_cap = (size_t)atoll(e);  // tests: force the overflow path
    if (fast) {
      CK(ctx->st_mn.ensure(tiles * TILE_THREADS * 8));
      CK(ctx->st_mx.ensure(tiles * TILE_THREADS * 8));
      CK(ctx->st_bm.ensure(tiles * TILE_THREADS * 8));
      CK(ctx->detail.ensure(detail_cap * sizeof(DetailEntry)));
      CK(ctx->detail_count.ensure(64));
      CK(cudaMemsetAsync(ctx->detail_count.p, 0, 4, st));
    }
    LevelArgs A;
    memset(&A, 0, sizeof A);
    A.pk = s->d_pk;
    A.brk = s->d_brk;
    A.ntiles = (int64_t)tiles;
    A.wfx = ctx->wfx.as<int64_t>();
    A.counts = tab.counts;
    A.cls = ctx->cls.as<uint16_t>();
    A.core = ctx->core.as<uint2>();
    A.hslots = ctx->lg_slots.as<HashSlot>();
    A.hmask = ctx->lg_mask;
    A.kmask64 = k < 32 ? ((((uint64_t)1) << (2 * k)) - 1) : ~0ull;
    A.pk_first = s->win_lo / 16;
    A.rk_pos = ctx->rk_pos.as<uint32_t>();
    A.rk_core = ctx->rk_core.as<uint32_t>();
    A.rk_p0 = ctx->rk_p0.as<uint32_t>();
    A.rk_x0 = ctx->sc_segx0.as<double>();
    A.rk_inc = ctx->sc_seginc.as<double>();
    A.rk_blob = ctx->rk_blob.p;
    A.rk_tail = ctx->rk_tail.as<int64_t>();
    A.rk_npieces = ctx->rk_npieces;
    A.rk_win_lo = ctx->rk_win_lo;
    A.rk_win_len = ctx->rk_win_len;
    A.rk_shift = ctx->rk_shift;
    A.rk_thr = tab.rk_thr;
    A.lut = ctx->lut_fx.as<int64_t>();
    A.core_lut = ctx->core_lut.as<int64_t>();
    A.lut_size = tab.lut_size;
    A.sp_count = ctx->lut_spc.as<uint32_t>();
    A.sp_val = ctx->lut_spv.as<int64_t>();
    A.sp_n = tab.sp_n;
    A.prm = d_prm;
    A.k = k;
    A.kmask = (uint32_t)(nk - 1);
    A.nseg = nseg;
    A.seg_start = ctx->seg_start.as<int64_t>();
    A.seg_len = ctx->seg_len.as<int64_t>();
    A.seg_chunk0 = ctx->seg_chunk0.as<uint64_t>();
    A.dense_start = 16 + 16 * dense_chunk0;
    A.pad_p0 = s->win_lo + 16;
    A.total_chunks = total_chunks;
    A.dense_first = dense_chunk0 == 0;
    const bool exchange = sh && level == 0;
    A.inscan = count_inscan ? d_inscan : nullptr;
    A.Q = (int64_t)(tiles * TILE_THREADS);
    A.st_c = ctx->st_c.as<uint32_t>();
    A.st_s = ctx->st_s.as<int64_t>();
    A.st_ea = ctx->st_ea.as<fx_t>();
    A.st_eb = ctx->st_eb.as<fx_t>();
    A.st_flags = ctx->st_flags.as<uint32_t>();
    A.st_p0 = ctx->st_p0.as<int64_t>();
    A.tile_xf = ctx->tile_xf.as<XfRec>();
    A.group_xf = ctx->group_xf.as<XfRec>();
    A.group_S = ctx->group_S.as<fx_t>();
    A.group_ex = ctx->group_ex.as<ExRec>();
    A.ngroups = (int64_t)((tiles + 31) / 32);
    const bool core_pipe = pair && tab.use_core && !count_inscan && getenv("KS_NO_CORE_PIPE") == nullptr;
    A.xf_log = core_pipe ? 5 : TILE_LOG;
    A.xf_ntiles = core_pipe ? (int64_t)tiles * TILE_WARPS : (int64_t)tiles;
    A.xf_ngroups = (A.xf_ntiles + 31) / 32;
    A.pending_list = ctx->pending_list.as<uint32_t>();
    A.pending_count = ctx->pending_count.as<unsigned int>();
    CK(cudaMemsetAsync(ctx->pending_count.p, 0, 4, st));
    A.tile_ex = ctx->tile_ex.as<ExRec>();
    A.pending = ctx->pending.as<ExPending>();
    A.S_start = 0;
    A.E_start.M = -(((fx_t)1) << 126);
    A.E_start.beg = -1; A.E_start.pk = -1; A.E_start.reset = 1; A.E_start.open = 0;
    A.launch_xf = exchange ? ctx->launch_rec.as<XfRec>() : nullptr;
    A.launch_ex = exchange ? reinterpret_cast<ExRec *>(ctx->launch_rec.as<char>() + 64) : nullptr;
    A.rec_beg = ctx->rec_beg.as<int64_t>();
    A.rec_pk = ctx->rec_pk.as<int64_t>();
    A.rec_c = ctx->rec_c.as<int64_t>();
    A.rec_mhi = ctx->rec_mhi.as<int64_t>();
    A.rec_mlo = ctx->rec_mlo.as<uint64_t>();
    A.rec_count = d_rec_count;
    A.rec_cap = ctx->rec_cap;
    A.inscan_mode = d_inscan != nullptr;
    CK(cudaMemsetAsync(d_rec_count + 1, 0, sizeof(unsigned long long), st));
    A.st_mn = ctx->st_mn.as<int64_t>();
    A.st_mx = ctx->st_mx.as<int64_t>();
    A.st_bm = ctx->st_bm.as<int64_t>();
    A.detail = ctx->detail.as<DetailEntry>();
    A.detail_count = ctx->detail_count.as<unsigned int>();
    A.detail_cap = (unsigned int)std::min<size_t>(detail_cap, 0xffffffffu);
    A.tr = tab.tr ? 1 : 0;
    A.buf = s->d_buf;
    A.nk = (uint32_t)nk;
    A.st_aux = ctx->st_aux.as<uint32_t>();
    A.child_pk = ctx->child_pk.as<int64_t>();
    A.child_c = ctx->child_c.as<int64_t>();
    A.child_count = ctx->child_count.as<unsigned long long>();
    A.child_cap = ctx->child_cap;
    if (tab.tr) CK(cudaMemsetAsync(ctx->child_count.p, 0, sizeof(unsigned long long), st));
    cudaEvent_t ps = ctx->prof_begin();
#define KS_GATHER(...) scan_gather_kernel<__VA_ARGS__><<<(unsigned)tiles, TILE_THREADS, 0, st>>>(A)
    if (tab.tr) KS_GATHER(0, true);
    else if (pair && tab.use_hash) KS_GATHER(4, false, true, false, true);
    else if (pair && tab.use_rank_core) KS_GATHER(3, false, true, true, true);
    else if (pair && tab.use_rank) KS_GATHER(3, false, true, false, true);
    else if (core_pipe) {
      if (!ctx->core_attr_set) {  // 34 KB of shared memory per CTA: ask for the large carve-out once
        CK(cudaFuncSetAttribute(scan_gather_core_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->core_attr_set = true;
      }
      scan_gather_core_kernel<<<(unsigned)std::min<size_t>(tiles, (size_t)148 * KS_CORE_MINBLOCKS), TILE_THREADS, 0, st>>>(A);
    } else if (pair && tab.use_core) KS_GATHER(2, false, true, true, true);
    else if (pair && tab.use_cls) KS_GATHER(2, false, true, false, true);
    else if (pair && tab.use_lut) KS_GATHER(1, false, true, false, true);
    else if (pair) KS_GATHER(0, false, true, false, true);
    else if (fast && tab.use_hash) KS_GATHER(4, false, true);
    else if (tab.use_hash) KS_GATHER(4);
    else if (fast && tab.use_rank_core) KS_GATHER(3, false, true, true);
    else if (fast && tab.use_rank) KS_GATHER(3, false, true);
    else if (tab.use_rank_core) KS_GATHER(3, false, false, true);
    else if (tab.use_rank) KS_GATHER(3);
    else if (fast && tab.use_core) KS_GATHER(2, false, true, true);
    else if (fast && tab.use_cls) KS_GATHER(2, false, true);
    else if (fast && tab.use_lut) KS_GATHER(1, false, true);
    else if (fast) KS_GATHER(0, false, true);
    else if (tab.use_cls) KS_GATHER(2);
    else if (tab.use_lut) KS_GATHER(1);
    else KS_GATHER(0);
#undef KS_GATHER
    group_scan_kernel<<<blocks_exact((size_t)A.xf_ngroups, 8), 256, 0, st>>>(A);
    group_top_kernel<<<1, TSCAN_THREADS, 0, st>>>(A);
    if (exchange && have_carry) {
      A.S_start = S_carry;
      A.launch_xf = nullptr;
      group_top_kernel<<<1, TSCAN_THREADS, 0, st>>>(A);
      LAUNCHED(1);
    } else if (exchange) {  // hand the shard's aggregate transform over, receive the state entering the shard
      XfRec mine;
      CK(cudaMemcpyAsync(&mine, A.launch_xf, sizeof mine, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      unsigned char carry[48];
      memset(carry, 0, sizeof carry);
      if (sh->fn(sh->user, 0, &mine, carry)) return ctx->fail(KS_ERR_ARG, "shard exchange (transform) failed");
      memcpy(&A.S_start, carry, sizeof(fx_t));
      S_carry = A.S_start;
      A.launch_xf = nullptr;
      group_top_kernel<<<1, TSCAN_THREADS, 0, st>>>(A);
      LAUNCHED(1);
    }
    if (tab.tr) scan_walk_kernel<0, true><<<(unsigned)tiles, TILE_THREADS, 0, st>>>(A);
    else if (pair) scan_walk_fast_kernel<true><<<(unsigned)tiles, TILE_THREADS, 0, st>>>(A);
    else if (fast) scan_walk_fast_kernel<false><<<(unsigned)tiles, TILE_THREADS, 0, st>>>(A);
    else if (tab.use_cls) scan_walk_kernel<2><<<(unsigned)tiles, TILE_THREADS, 0, st>>>(A);
    else if (tab.use_lut) scan_walk_kernel<1><<<(unsigned)tiles, TILE_THREADS, 0, st>>>(A);
    else scan_walk_kernel<0><<<(unsigned)tiles, TILE_THREADS, 0, st>>>(A);
    group_ex_kernel<<<blocks_exact((size_t)A.ngroups, 8), 256, 0, st>>>(A);
    if (fast) {
      // the list is short (one entry per tile at most, plus the rare wide excursions); the kernel reads
      // its length on the device and strides over it, so no host round trip sits between the kernels
      const unsigned dgrid = (unsigned)std::min<size_t>(blocks_exact(tiles + 1024, 128), 148u * 8u);
      if (pair && tab.use_hash) scan_detail_kernel<4, true><<<dgrid, 128, 0, st>>>(A);
      else if (pair && tab.use_rank) scan_detail_kernel<3, true><<<dgrid, 128, 0, st>>>(A);
      else if (pair && tab.use_cls) scan_detail_kernel<2, true><<<dgrid, 128, 0, st>>>(A);
      else if (pair && tab.use_lut) scan_detail_kernel<1, true><<<dgrid, 128, 0, st>>>(A);
      else if (pair) scan_detail_kernel<0, true><<<dgrid, 128, 0, st>>>(A);
      else if (tab.use_hash) scan_detail_kernel<4><<<dgrid, 128, 0, st>>>(A);
      else if (tab.use_rank) scan_detail_kernel<3><<<dgrid, 128, 0, st>>>(A);
      else if (tab.use_cls) scan_detail_kernel<2><<<dgrid, 128, 0, st>>>(A);
      else if (tab.use_lut) scan_detail_kernel<1><<<dgrid, 128, 0, st>>>(A);
      else scan_detail_kernel<0><<<dgrid, 128, 0, st>>>(A);
      LAUNCHED(1);
    }
    if (exchange && have_carry) {
      A.E_start = E_carry;
    } else if (exchange) {  // the same for the open-excursion state
      ex_top_kernel<<<1, 32, 0, st>>>(A);
      LAUNCHED(1);
      ExRec mine;
      CK(cudaMemcpyAsync(&mine, A.launch_ex, sizeof mine, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      unsigned char carry[48];
      memset(carry, 0, sizeof carry);
      if (sh->fn(sh->user, 1, &mine, carry)) return ctx->fail(KS_ERR_ARG, "shard exchange (excursion) failed");
      memcpy(&A.E_start, carry, sizeof(ExRec));
      E_carry = A.E_start;
      have_carry = true;
    }
    ex_fixup_kernel<<<148, 256, 0, st>>>(A);
    ctx->prof_end(level == 0 ? KS_PROF_SCAN0 : KS_PROF_SCANN, ps);
    LAUNCHED(6);
    CK(cudaGetLastError());
    struct { unsigned long long cnt, child_chunks, children; } hres;
    hres.children = 0;
    unsigned int n_detail = 0;
    DevScanParams hprm;
    CK(cudaMemcpyAsync(&hres.cnt, d_rec_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (fast) CK(cudaMemcpyAsync(&n_detail, ctx->detail_count.p, 4, cudaMemcpyDeviceToHost, st));
    if (tab.tr)
      CK(cudaMemcpyAsync(&hres.children, ctx->child_count.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (level == 0) CK(cudaMemcpyAsync(&hprm, d_prm, sizeof hprm, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (level == 0 && (hprm.err & 1))
      return ctx->fail(KS_ERR_RANGE, "a k-mer weight is +Inf or >= 2^40: outside the exact scan range");
    if (level == 0 && (hprm.err & 4))
      return ctx->fail(KS_ERR_RANGE, "a nonzero k-mer weight is more than 2^57 times smaller than the largest one: "
                                     "outside the exact scan range");
    if (level == 0 && tab.tr && (hprm.err & 2))
      return ctx->fail(KS_ERR_RANGE, "a transition or k-mer score is NaN");
    count_inscan = false;  // every position of this level has been counted, also if we must retry
    if (fast && n_detail > A.detail_cap) {
      // more undecided chunks than the list holds: redo the level with the general walk
      fast = false;
      CK(cudaMemcpyAsync(d_rec_count, &level_start, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      continue;
    }
    if (hres.cnt > ctx->rec_cap || hres.children > ctx->child_cap) {
      // record buffer too small: grow (keeping earlier levels), rewind the counter, redo the level
      if (hres.cnt > ctx->rec_cap) rc = ensure_recs(ctx, (size_t)hres.cnt + (size_t)hres.cnt / 8 + 1024);
      if (!rc && hres.children > ctx->child_cap)
        rc = ensure_children(ctx, (size_t)hres.children + (size_t)hres.children / 8 + 1024);
      if (rc) return rc;
      CK(cudaMemcpyAsync(d_rec_count, &level_start, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      continue;
    }
    fast = fast_ok;
    rec_total = hres.cnt;
    if (level > 0) revisit_chunks += (uint64_t)total_chunks;
    unsigned long long n_new = tab.tr ? hres.children : hres.cnt - level_start;
    ++level;
    if (n_new == 0) break;
    if (!tab.tr && hres.child_chunks == 0) break;  // no record of this level spawns a re-scan
    // child segments of the records this level emitted
    CK(ctx->seg_start.ensure(n_new * 8));
    CK(ctx->seg_len.ensure(n_new * 8));
    CK(ctx->seg_chunks.ensure(n_new * 8));
    CK(ctx->seg_chunk0.ensure((n_new + 1) * 8));
    CK(ctx->scan_tmp.ensure(exclusive_scan_scratch_elems(n_new) * 8));
    if (tab.tr)  // the re-scan requests were filtered when they were appended
      seg_build_kernel<<<blocks_exact(n_new, 256), 256, 0, st>>>(
          ctx->child_pk.as<int64_t>(), ctx->child_c.as<int64_t>(), 0, n_new, 0, 1, ctx->seg_start.as<int64_t>(),
          ctx->seg_len.as<int64_t>(), ctx->seg_chunks.as<uint64_t>());
    else
      seg_build_kernel<<<blocks_exact(n_new, 256), 256, 0, st>>>(
          ctx->rec_pk.as<int64_t>(), ctx->rec_c.as<int64_t>(), level_start, n_new, mw, d_inscan != nullptr,
          ctx->seg_start.as<int64_t>(), ctx->seg_len.as<int64_t>(), ctx->seg_chunks.as<uint64_t>());
    LAUNCHED(1);
    LAUNCHED((exclusive_scan<uint64_t, uint64_t>(ctx->seg_chunks.as<uint64_t>(), n_new,
                                                 ctx->seg_chunk0.as<uint64_t>(), ctx->scan_tmp.as<uint64_t>(), st)));
    CK(cudaGetLastError());
    uint64_t tot = 0;
    CK(cudaMemcpyAsync(&tot, ctx->seg_chunk0.as<uint64_t>() + n_new, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    level_start = hres.cnt;
    nseg = (int64_t)n_new;
    total_chunks = (int64_t)tot;
    if (d_inscan) count_inscan = true;
    if (total_chunks == 0) break;
  }
  ctx->last_levels = level;
  ctx->last_revisit_chunks = revisit_chunks;

  // order by start position and convert to the reference layout
  const unsigned long long n = rec_total;  // read back at the end of the last level
  if (n_spans) *n_spans = n;
  if (n == 0) return KS_OK;
  if (n > 0xfffffff0ull) return ctx->fail(KS_ERR_NOMEM, "too many spans");
  const uint32_t *d_perm = nullptr;
  if (n > 1 && n <= (unsigned long long)SMALL_SORT_MAX) {
    CK(ctx->sort_vals_a.ensure(n * 4));
    small_sort_kernel<<<1, 1024, 0, st>>>(ctx->rec_beg.as<int64_t>(), (unsigned int)n, ctx->sort_vals_a.as<uint32_t>());
    LAUNCHED(1);
    CK(cudaGetLastError());
    d_perm = ctx->sort_vals_a.as<uint32_t>();
  } else if (n > 1) {
    CK(ctx->sort_keys_a.ensure(n * 8));
    CK(ctx->sort_keys_b.ensure(n * 8));
    CK(ctx->sort_vals_a.ensure(n * 4));
    CK(ctx->sort_vals_b.ensure(n * 4));
    size_t nb = radix_nblocks(n);
    CK(ctx->sort_hist.ensure((256 * nb + 2) * 4));
    CK(ctx->sort_scan.ensure(exclusive_scan_scratch_elems(256 * nb) * 4));
    copy_keys_kernel<<<blocks_exact(n, 256), 256, 0, st>>>(ctx->rec_beg.as<int64_t>(), n,
                                                           ctx->sort_keys_a.as<uint64_t>());
    LAUNCHED(1);
    int nbits = 0;
    while (nbits < 63 && ((uint64_t)s->total >> nbits) != 0) ++nbits;
    RadixScratch rs{ctx->sort_hist.as<uint32_t>(), ctx->sort_scan.as<uint32_t>()};
    uint64_t *skeys = nullptr;
    uint32_t *svals = nullptr;
    LAUNCHED(radix_sort_pairs<uint64_t>(ctx->sort_keys_a.as<uint64_t>(), ctx->sort_vals_a.as<uint32_t>(),
                                        ctx->sort_keys_b.as<uint64_t>(), ctx->sort_vals_b.as<uint32_t>(), n,
                                        nbits, true, rs, st, &skeys, &svals));
    CK(cudaGetLastError());
    d_perm = svals;
  }
  CK(ctx->out_pos.ensure(n * 3 * 4));
  CK(ctx->out_score.ensure(n * 2 * 8));
  finalize_kernel<<<blocks_exact(n, 256), 256, 0, st>>>(
      d_perm, n, ctx->rec_beg.as<int64_t>(), ctx->rec_pk.as<int64_t>(), ctx->rec_mhi.as<int64_t>(),
      ctx->rec_mlo.as<uint64_t>(), s->d_starts, s->nseq, d_prm, ctx->out_pos.as<int32_t>(),
      ctx->out_score.as<double>(), tab.tr ? 1 : 0);
  LAUNCHED(1);
  CK(cudaGetLastError());
  if (host_out) {
    host_out->pos = (int32_t *)malloc(n * 3 * sizeof(int32_t));
    host_out->score = (double *)malloc(n * 2 * sizeof(double));
    if (!host_out->pos || !host_out->score) { ks_spans_free(host_out); return ctx->fail(KS_ERR_NOMEM, "out of host memory"); }
    host_out->n = (size_t)n;
    CK(cudaMemcpyAsync(host_out->pos, ctx->out_pos.p, n * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(host_out->score, ctx->out_score.p, n * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return KS_OK;
}

static int scan_table_impl(ks_ctx *ctx, const ks_seqset *s, int k, const double *d_W, double thr, int min_width,
                           double min_score, int32_t *d_inscan, ks_spans *host_out, uint64_t *n_spans,
                           const ShardCtl *sh) {
  if (!ctx) return KS_ERR_ARG;
  if (!s || !d_W) return ctx->fail(KS_ERR_ARG, "ks_dev_scan: null argument");
  int rc = check_k(ctx, k);
  if (rc) return rc;
  if (host_out) { host_out->pos = nullptr; host_out->score = nullptr; host_out->n = 0; }
  if (n_spans) *n_spans = 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nk = (size_t)1 << (2 * k);
  const uint64_t mw = (uint64_t)(int64_t)min_width;  // negative R integers wrap exactly like size_t (:243)
  // score table -> exact fixed point (W - thr as the reference computes it at :268)
  CK(ctx->wfx.ensure(nk * 8));
  CK(ctx->prm.ensure(sizeof(DevScanParams)));
  DevScanParams *d_prm = ctx->prm.as<DevScanParams>();
  CK(cudaMemsetAsync(d_prm, 0, sizeof(DevScanParams), st));
  cudaEvent_t pw = ctx->prof_begin();
  wmax_kernel<<<grid_for(nk, 256), 256, 0, st>>>(d_W, nk, thr, d_prm);
  wfx_kernel<<<grid_for(nk, 256), 256, 0, st>>>(d_W, nk, thr, ctx->wfx.as<int64_t>(), d_prm, mw, min_score);
  ctx->prof_end(KS_PROF_WFX, pw);
  LAUNCHED(2);
  CK(cudaGetLastError());
  ScanTable tab;
  return scan_core(ctx, s, k, tab, mw, d_inscan, host_out, n_spans, sh);
}

int ks_dev_scan(ks_ctx *ctx, const ks_seqset *s, int k, const double *d_W, double thr, int min_width,
                double min_score, int32_t *d_inscan, ks_spans *host_out, uint64_t *n_spans) {
  KS_TRY
  return scan_table_impl(ctx, s, k, d_W, thr, min_width, min_score, d_inscan, host_out, n_spans, nullptr);
  KS_CATCH(ctx)
}

int ks_dev_scan_shard(ks_ctx *ctx, const ks_seqset *s, int k, const double *d_W, double thr, int min_width,
                      double min_score, int64_t chunk0, int64_t nchunks, ks_exchange_fn fn, void *user,
                      ks_spans *host_out, uint64_t *n_spans) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!fn) return ctx->fail(KS_ERR_ARG, "ks_dev_scan_shard: null exchange function");
  ShardCtl sh;
  sh.chunk0 = chunk0; sh.nchunks = nchunks; sh.fn = fn; sh.user = user;
  return scan_table_impl(ctx, s, k, d_W, thr, min_width, min_score, nullptr, host_out, n_spans, &sh);
  KS_CATCH(ctx)
}

// Transition-score scan (tr_lr_regions_r core, :329-395): d_init / d_trans are double[4^k] in 2-bit code order.
int ks_dev_tr_lr_regions(ks_ctx *ctx, const ks_seqset *s, int k, const double *d_init, const double *d_trans,
                         int min_length, ks_spans *host_out, uint64_t *n_spans) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!s || !d_init || !d_trans) return ctx->fail(KS_ERR_ARG, "ks_dev_tr_lr_regions: null argument");
  if (k < 1 || k > 15) return ctx->fail(KS_ERR_ARG, "k should be a positive value less than MAX_K");
  if (min_length < 0) return ctx->fail(KS_ERR_ARG, "min_length should be a positive integer");
  if (host_out) { host_out->pos = nullptr; host_out->score = nullptr; host_out->n = 0; }
  if (n_spans) *n_spans = 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nk = (size_t)1 << (2 * k);
  // both tables -> one exact fixed-point table [trans | init] with a common scale; no score threshold
  CK(ctx->wfx.ensure(2 * nk * 8));
  CK(ctx->prm.ensure(sizeof(DevScanParams)));
  DevScanParams *d_prm = ctx->prm.as<DevScanParams>();
  CK(cudaMemsetAsync(d_prm, 0, sizeof(DevScanParams), st));
  const double no_min = -INFINITY;
  cudaEvent_t pw = ctx->prof_begin();
  wmax_kernel<<<grid_for(nk, 256), 256, 0, st>>>(d_trans, nk, 0.0, d_prm);
  wmax_kernel<<<grid_for(nk, 256), 256, 0, st>>>(d_init, nk, 0.0, d_prm);
  wfx_kernel<<<grid_for(nk, 256), 256, 0, st>>>(d_trans, nk, 0.0, ctx->wfx.as<int64_t>(), d_prm,
                                               (uint64_t)min_length, no_min);
  wfx_kernel<<<grid_for(nk, 256), 256, 0, st>>>(d_init, nk, 0.0, ctx->wfx.as<int64_t>() + nk, d_prm,
                                               (uint64_t)min_length, no_min);
  ctx->prof_end(KS_PROF_WFX, pw);
  LAUNCHED(4);
  CK(cudaGetLastError());
  ScanTable tab;
  tab.tr = true;
  return scan_core(ctx, s, k, tab, (uint64_t)min_length, nullptr, host_out, n_spans, nullptr);
  KS_CATCH(ctx)
}

// Scan with score = f(count): the count -> score function is the one the last
// ks_dev_scores(mode LOG2 | SIGN) on this ctx derived (cached per distinct count on the host).
// The kernel gathers the 4-byte count (table stays L2 resident up to k = 12) instead of an 8-byte score.
static int scan_counts_impl(ks_ctx *ctx, const ks_seqset *s, int k, const int32_t *d_counts, double thr,
                            int min_width, double min_score, ks_spans *host_out, uint64_t *n_spans,
                            const ShardCtl *sh) {
  if (!ctx) return KS_ERR_ARG;
  if (!s || !d_counts) return ctx->fail(KS_ERR_ARG, "ks_dev_scan_counts: null argument");
  int rc = check_k(ctx, k);
  if (rc) return rc;
  if (!ctx->lut_valid || ctx->lut_k != k)
    return ctx->fail(KS_ERR_ARG, "ks_dev_scan_counts: call ks_dev_scores with mode LOG2 or SIGN for this k first");
  if (host_out) { host_out->pos = nullptr; host_out->score = nullptr; host_out->n = 0; }
  if (n_spans) *n_spans = 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t mw = (uint64_t)(int64_t)min_width;
  cudaEvent_t pw = ctx->prof_begin();
  const size_t ng = ctx->lut_gcount.size();
  // host: range check, scale, parameters (O(#distinct counts))
  double wmax = 0;
  for (size_t g = 0; g < ng; ++g) {
    double w = ctx->lut_gval[g] - thr;
    if (w >= 0x1p40) return ctx->fail(KS_ERR_RANGE, "a k-mer score is +Inf or >= 2^40 (median frequency 0?): outside the exact scan range");
    if (w == w && w > -0x1p40 && fabs(w) > wmax) wmax = fabs(w);
  }
  DevScanParams hp;
  memset(&hp, 0, sizeof hp);
  hp.qs = qs_for_max(wmax);
  fx_t mu = fx_ceil_units(min_score, hp.qs);
  hp.min_width = mw;
  hp.min_lo = (uint64_t)(unsigned __int128)mu;
  hp.min_hi = (int64_t)(uint64_t)(((unsigned __int128)mu) >> 64);
  if (ctx->cls_counts == (const void *)d_counts && ctx->cls_n == ((size_t)1 << (2 * k)) && ng && ng <= 65535) {
    // class mode: one table entry per distinct count, gathered through the 2-byte class table
    // staged in pinned memory (HPIN_FX / HPIN_PRM are written by this function only); the last scan on this
    // context ended with a synchronisation, so nothing still reads these two regions
    rc = ensure_hpin(ctx);
    if (rc) return rc;
    int64_t *fx = reinterpret_cast<int64_t *>(ctx->hpin + ks_ctx::HPIN_FX);
    DevScanParams *hp_pin = reinterpret_cast<DevScanParams *>(ctx->hpin + ks_ctx::HPIN_PRM);
    for (size_t g = 0; g < ng; ++g) fx[g] = wfx_from_double(ctx->lut_gval[g] - thr, hp.qs);
    *hp_pin = hp;
    CK(ctx->prm.ensure(sizeof(DevScanParams)));
    CK(ctx->lut_fx.ensure(ng * 8 + 8));
    CK(cudaMemcpyAsync(ctx->prm.p, hp_pin, sizeof hp, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->lut_fx.p, fx, ng * 8, cudaMemcpyHostToDevice, st));
    if (ctx->core_valid) {  // the scores of the classes the core records name, in their byte order
      int64_t *cfx = reinterpret_cast<int64_t *>(ctx->hpin + ks_ctx::HPIN_COREFX);
      memset(cfx, 0, (size_t)CORE_ESCAPE * 8);
      for (size_t i = 0; i < ctx->core_groups.size(); ++i) cfx[i] = fx[ctx->core_groups[i]];
      CK(ctx->core_lut.ensure((size_t)CORE_ESCAPE * 8 + 8));
      CK(cudaMemcpyAsync(ctx->core_lut.p, cfx, (size_t)CORE_ESCAPE * 8, cudaMemcpyHostToDevice, st));
    }
    ctx->prof_end(KS_PROF_WFX, pw);
    ScanTable tab;
    tab.use_lut = true;
    tab.use_cls = true;
    tab.use_core = ctx->core_valid;
    tab.lut_size = (uint32_t)ng;
    rc = scan_core(ctx, s, k, tab, mw, nullptr, host_out, n_spans, sh);
    const int rcj = join_table(ctx);  // a score table written next to this scan (ks_ctx_side_table)
    return rc ? rc : rcj;
  }
  uint32_t maxc = ng ? ctx->lut_gcount[ng - 1] : 0;
  uint32_t lut_size = maxc < (1u << 16) ? maxc + 1 : (1u << 16);
  size_t sp_first = std::lower_bound(ctx->lut_gcount.begin(), ctx->lut_gcount.end(), lut_size) - ctx->lut_gcount.begin();
  uint32_t sp_n = (uint32_t)(ng - sp_first);
  CK(ctx->prm.ensure(sizeof(DevScanParams)));
  CK(ctx->lut_fx.ensure((size_t)lut_size * 8 + 8));
  CK(ctx->lut_spc.ensure((size_t)sp_n * 4 + 8));
  CK(ctx->lut_spv.ensure((size_t)sp_n * 8 + 8));
  CK(ctx->sc_gcount.ensure((ng + 1) * 4));
  CK(ctx->sc_lut.ensure(ng * 8 + 8));
  CK(cudaMemcpyAsync(ctx->prm.p, &hp, sizeof hp, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->sc_gcount.p, ctx->lut_gcount.data(), ng * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->sc_lut.p, ctx->lut_gval.data(), ng * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->lut_fx.p, 0, (size_t)lut_size * 8, st));
  if (ng) {
    lut_build_kernel<<<blocks_exact(ng, 256), 256, 0, st>>>(ctx->sc_gcount.as<uint32_t>(), ctx->sc_lut.as<double>(),
                                                           (uint32_t)ng, thr, hp.qs, lut_size, ctx->lut_fx.as<int64_t>(),
                                                           ctx->lut_spc.as<uint32_t>(), ctx->lut_spv.as<int64_t>(),
                                                           (uint32_t)sp_first);
    LAUNCHED(1);
  }
  ctx->prof_end(KS_PROF_WFX, pw);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));  // hp lives on this stack frame
  ScanTable tab;
  tab.use_lut = true;
  tab.counts = reinterpret_cast<const uint32_t *>(d_counts);
  tab.lut_size = lut_size;
  tab.sp_n = sp_n;
  return scan_core(ctx, s, k, tab, mw, nullptr, host_out, n_spans, sh);
}

int ks_dev_scan_counts(ks_ctx *ctx, const ks_seqset *s, int k, const int32_t *d_counts, double thr,
                       int min_width, double min_score, ks_spans *host_out, uint64_t *n_spans) {
  KS_TRY
  return scan_counts_impl(ctx, s, k, d_counts, thr, min_width, min_score, host_out, n_spans, nullptr);
  KS_CATCH(ctx)
}

int ks_dev_scan_counts_shard(ks_ctx *ctx, const ks_seqset *s, int k, const int32_t *d_counts, double thr,
                             int min_width, double min_score, int64_t chunk0, int64_t nchunks,
                             ks_exchange_fn fn, void *user, ks_spans *host_out, uint64_t *n_spans) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!fn) return ctx->fail(KS_ERR_ARG, "ks_dev_scan_counts_shard: null exchange function");
  ShardCtl sh;
  sh.chunk0 = chunk0; sh.nchunks = nchunks; sh.fn = fn; sh.user = user;
  return scan_counts_impl(ctx, s, k, d_counts, thr, min_width, min_score, host_out, n_spans, &sh);
  KS_CATCH(ctx)
}

// Scan in rank mode (the mode the reference codes, :268,602-612): score = rank - thr with the rank table the last
// ks_dev_scores(mode RANK) on this ctx derived.  The kernel gathers the 4-byte position of the k-mer in the rank
// order (4^k x 4 B, L2 resident up to k = 12, against 4^k x 8 B for the score table) and evaluates the rank from
// the linear pieces of the closed form -- the same fma that filled the rank table, so both paths agree bit for bit.
static int scan_ranks_impl(ks_ctx *ctx, const ks_seqset *s, int k, double thr, int min_width, double min_score,
                           ks_spans *host_out, uint64_t *n_spans, const ShardCtl *sh) {
  if (!ctx) return KS_ERR_ARG;
  if (!s) return ctx->fail(KS_ERR_ARG, "ks_dev_scan_ranks: null argument");
  int rc = check_k(ctx, k);
  if (rc) return rc;
  if (!ctx->rk_valid || ctx->rk_k != k)
    return ctx->fail(KS_ERR_ARG, "ks_dev_scan_ranks: call ks_dev_scores with mode RANK for this k first");
  if (host_out) { host_out->pos = nullptr; host_out->score = nullptr; host_out->n = 0; }
  if (n_spans) *n_spans = 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t mw = (uint64_t)(int64_t)min_width;
  cudaEvent_t pw = ctx->prof_begin();
  // ranks run from 0 (first k-mer in order) to rk_max, so max |rank - thr| sits at one of the two ends
  const double wmax = std::max(fabs(0.0 - thr), fabs(ctx->rk_max - thr));
  if (!(wmax < 0x1p40)) return ctx->fail(KS_ERR_RANGE, "threshold outside the exact scan range");
  rc = ensure_hpin(ctx);
  if (rc) return rc;
  DevScanParams *hp = reinterpret_cast<DevScanParams *>(ctx->hpin + ks_ctx::HPIN_PRM);
  memset(hp, 0, sizeof *hp);
  hp->qs = qs_for_max(wmax);
  fx_t mu = fx_ceil_units(min_score, hp->qs);
  hp->min_width = mw;
  hp->min_lo = (uint64_t)(unsigned __int128)mu;
  hp->min_hi = (int64_t)(uint64_t)(((unsigned __int128)mu) >> 64);
  CK(ctx->prm.ensure(sizeof(DevScanParams)));
  CK(cudaMemcpyAsync(ctx->prm.p, hp, sizeof *hp, cudaMemcpyHostToDevice, st));
  {  // finished scores of the rank-order positions outside the shared-memory window (depends on thr)
    const uint32_t ntail = ctx->rk_n - ctx->rk_win_len;
    CK(ctx->rk_tail.ensure((size_t)ntail * 8 + 8));
    if (ntail) {
      LevelArgs A;
      memset(&A, 0, sizeof A);
      A.prm = ctx->prm.as<DevScanParams>();
      A.rk_p0 = ctx->rk_p0.as<uint32_t>();
      A.rk_x0 = ctx->sc_segx0.as<double>();
      A.rk_inc = ctx->sc_seginc.as<double>();
      A.rk_npieces = ctx->rk_npieces;
      A.rk_win_lo = ctx->rk_win_lo;
      A.rk_win_len = ctx->rk_win_len;
      A.rk_thr = thr;
      rank_tail_kernel<<<grid_for(ntail, 256), 256, 0, st>>>(A, ntail, ctx->rk_tail.as<int64_t>());
      LAUNCHED(1);
      CK(cudaGetLastError());
    }
  }
  ScanTable tab;
  tab.use_rank = true;
  tab.rk_thr = thr;
  // a position table far beyond L2 (4^13 x 4 B = 268 MB): regroup it by (k-1)-mer so that the two positions a pair of
  // bases needs sit in one sector.  Rebuilt per scan: 0.3 ms at k = 13 against tens of ms of gathers.
  int core_min_k = 13;
  if (const char *e = getenv("KS_RANK_CORE_MIN_K")) core_min_k = atoi(e);
  if (k >= core_min_k && k >= 2) {
    const size_t ncore = (size_t)1 << (2 * k - 2);
    CK(ctx->rk_core.ensure(ncore * 32));
    rank_core_apply_kernel<<<grid_for(ncore, 256), 256, 0, st>>>(ctx->rk_pos.as<uint32_t>(), ncore, ctx->rk_core.as<uint4>());
    LAUNCHED(1);
    CK(cudaGetLastError());
    tab.use_rank_core = true;
  }
  ctx->prof_end(KS_PROF_WFX, pw);
  return scan_core(ctx, s, k, tab, mw, nullptr, host_out, n_spans, sh);
}

int ks_dev_scan_ranks(ks_ctx *ctx, const ks_seqset *s, int k, double thr, int min_width, double min_score,
                      ks_spans *host_out, uint64_t *n_spans) {
  KS_TRY
  return scan_ranks_impl(ctx, s, k, thr, min_width, min_score, host_out, n_spans, nullptr);
  KS_CATCH(ctx)
}

int ks_dev_scan_ranks_shard(ks_ctx *ctx, const ks_seqset *s, int k, double thr, int min_width, double min_score,
                            int64_t chunk0, int64_t nchunks, ks_exchange_fn fn, void *user, ks_spans *host_out,
                            uint64_t *n_spans) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!fn) return ctx->fail(KS_ERR_ARG, "ks_dev_scan_ranks_shard: null exchange function");
  ShardCtl sh;
  sh.chunk0 = chunk0; sh.nchunks = nchunks; sh.fn = fn; sh.user = user;
  return scan_ranks_impl(ctx, s, k, thr, min_width, min_score, host_out, n_spans, &sh);
  KS_CATCH(ctx)
}

// ------------------------------------------------------------------------------------------------
// Large k (1 <= k <= 31 accepted, meant for k >= 16; BASELINE.json configs[3]: k = 21): hash-table counting, rank
// over the k-mers that occur in (count, code) order (mode KS_MODE_RANK) or +-1 around the frequency `param` (mode
// KS_MODE_SIGN), scan with the scores kept in the table's slots.  One GPU; everything stays resident.
int ks_dev_large_regions(ks_ctx *ctx, const ks_seqset *s, int k, int mode, double param, double thr, int min_width,
                         double min_score, double *n_words, uint64_t *n_distinct, ks_spans *host_out,
                         uint64_t *n_spans) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!s) return ctx->fail(KS_ERR_ARG, "ks_dev_large_regions: null argument");
  if (k < 1 || k > 31) return ctx->fail(KS_ERR_ARG, "the large-k path takes k between 1 and 31 (got %d)", k);
  if (mode != KS_MODE_RANK && mode != KS_MODE_SIGN)
    return ctx->fail(KS_ERR_ARG, "the large-k path scores by weighted rank (mode 0) or +-1 around a frequency (mode 2)");
  if (mode == KS_MODE_SIGN && !isfinite(param))
    return ctx->fail(KS_ERR_ARG, "mode 2 at large k needs an explicit frequency threshold (the median of 4^k entries is 0)");
  if (s->window) return ctx->fail(KS_ERR_ARG, "the large-k path runs on whole sets (one GPU)");
  if (host_out) { host_out->pos = nullptr; host_out->score = nullptr; host_out->n = 0; }
  if (n_spans) *n_spans = 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int rc = ensure_packed(ctx, s);
  if (rc) return rc;
  // ---- hash table: power of two, load factor <= 0.75 even if every position holds a new k-mer ----
  uint64_t cap = 1024;
  while ((double)cap * 0.75 < (double)s->bases + 1.0) cap <<= 1;
  if (const char *e = getenv("KS_HASH_CAP")) cap = (uint64_t)atoll(e);  // tests: small tables, long probe chains
  if (cap > (1ull << 32)) return ctx->fail(KS_ERR_ARG, "input too large for the large-k path on one GPU (2^32 slots)");
  CK(ctx->lg_slots.ensure((size_t)cap * sizeof(HashSlot)));
  CK(ctx->lg_stats.ensure(64));
  CK(cudaMemsetAsync(ctx->lg_slots.p, 0, (size_t)cap * sizeof(HashSlot), st));
  CK(cudaMemsetAsync(ctx->lg_stats.p, 0, 64, st));
  HashArgs H;
  H.slots = ctx->lg_slots.as<HashSlot>();
  H.mask = cap - 1;
  H.stats = ctx->lg_stats.as<unsigned long long>();
  ctx->lg_mask = cap - 1;
  ctx->lg_k = k;
  ctx->lg_nd = 0;
  const int64_t nch = s->total / 16 - 1;  // chunks 1 .. total/16 - 1 hold the data (chunk 0 is the front pad)
  cudaEvent_t pe = ctx->prof_begin();
  hash_count_kernel<<<grid_for((size_t)nch, 256, 148u * 8u), 256, 0, st>>>(s->d_pk, s->d_brk, s->d_buf, 1, nch, k, H);
  ctx->prof_end(KS_PROF_COUNT, pe);
  LAUNCHED(1);
  CK(cudaGetLastError());
  cudaEvent_t psc = ctx->prof_begin();
  // ---- how many k-mers occur, and the largest count ----
  hash_stats_kernel<<<grid_for((size_t)cap, 256, 148u * 8u), 256, 0, st>>>(H);
  LAUNCHED(1);
  unsigned long long stats[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(stats, ctx->lg_stats.p, sizeof stats, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (stats[3] & 1ull) return ctx->fail(KS_ERR_NOMEM, "the k-mer hash table is full");
  if (stats[2] > 0x7fffffffull) return ctx->fail(KS_ERR_RANGE, "a k-mer count exceeds 2^31 - 1");
  const double total = (double)stats[0];
  const uint64_t nd = stats[1];
  if (n_words) *n_words = total;
  if (n_distinct) *n_distinct = nd;
  ctx->lg_nd = nd;
  ctx->lg_idx = nullptr;
  ctx->lg_cnt = nullptr;
  if (nd == 0) { ctx->prof_end(KS_PROF_SCORES, psc); return KS_OK; }
  int cbits = 0;
  while (cbits < 64 && (stats[2] >> cbits) != 0) ++cbits;
  // ---- order (count, code).  When count and code fit one 64-bit key, ONE stable radix sort of the composite key;
  //      otherwise codes first, then the counts (stable) with the code-order index as payload ----
  const bool composite = 2 * k + cbits <= 64 && getenv("KS_LARGE_TWO_PASS") == nullptr;
  CK(ctx->lg_comp_a.ensure((size_t)nd * 8));
  CK(ctx->lg_slot_a.ensure((size_t)nd * 4));
  CK(ctx->lg_comp_b.ensure((size_t)nd * 8));
  CK(ctx->lg_slot_b.ensure((size_t)nd * 4));
  unsigned long long *d_cursor = ctx->lg_stats.as<unsigned long long>() + 4;
  hash_compact_kernel<<<grid_for((size_t)cap, 256, 148u * 8u), 256, 0, st>>>(
      H, k, composite ? 1 : 0, ctx->lg_comp_a.as<uint64_t>(), ctx->lg_slot_a.as<uint32_t>(), nd, d_cursor);
  LAUNCHED(1);
  size_t nb = radix_nblocks((size_t)nd);
  CK(ctx->sort_hist.ensure((256 * nb + 2) * 4));
  CK(ctx->sort_scan.ensure(exclusive_scan_scratch_elems(256 * nb) * 4));
  RadixScratch rs{ctx->sort_hist.as<uint32_t>(), ctx->sort_scan.as<uint32_t>()};
  uint64_t *skeys = nullptr;
  uint32_t *sslots = nullptr;
  LAUNCHED(radix_sort_pairs<uint64_t>(ctx->lg_comp_a.as<uint64_t>(), ctx->lg_slot_a.as<uint32_t>(),
                                      ctx->lg_comp_b.as<uint64_t>(), ctx->lg_slot_b.as<uint32_t>(), (size_t)nd,
                                      composite ? 2 * k + cbits : 2 * k, false, rs, st, &skeys, &sslots));
  CK(cudaGetLastError());
  ctx->lg_sorted = skeys;
  uint32_t *scnt = nullptr, *sidx = nullptr;
  if (!composite) {
    CK(ctx->lg_cnt_a.ensure((size_t)nd * 4));
    CK(ctx->lg_cnt_b.ensure((size_t)nd * 4));
    CK(ctx->lg_idx_a.ensure((size_t)nd * 4));
    CK(ctx->lg_idx_b.ensure((size_t)nd * 4));
    large_counts_kernel<<<blocks_exact((size_t)nd, 256), 256, 0, st>>>(ctx->lg_slots.as<HashSlot>(), sslots, nd,
                                                                      ctx->lg_cnt_a.as<uint32_t>());
    LAUNCHED(1);
    LAUNCHED(radix_sort_pairs<uint32_t>(ctx->lg_cnt_a.as<uint32_t>(), ctx->lg_idx_a.as<uint32_t>(),
                                        ctx->lg_cnt_b.as<uint32_t>(), ctx->lg_idx_b.as<uint32_t>(), (size_t)nd, cbits,
                                        true, rs, st, &scnt, &sidx));
    CK(cudaGetLastError());
    if (cbits == 0 || sidx == nullptr) return ctx->fail(KS_ERR_CUDA, "internal: empty count sort");
    ctx->lg_idx = sidx;
    ctx->lg_cnt = scnt;
  }
  // ---- run-length table of the counts, pieces of the sequential accumulation (host, O(#distinct counts)) ----
  size_t gcap = (size_t)(sqrt(2.0 * (total > 0 ? total : 1.0)) + 16.0);
  CK(ctx->sc_small.ensure(64));
  uint32_t *d_ngroups = ctx->sc_small.as<uint32_t>() + 1;
  std::vector<uint32_t> gcount, gstart32;
  for (int attempt = 0; attempt < 2; ++attempt) {
    CK(ctx->sc_gcount.ensure((gcap + 1) * 4));
    CK(ctx->sc_gstart.ensure((gcap + 1) * 4));
    CK(cudaMemsetAsync(d_ngroups, 0, 4, st));
    if (composite)
      large_heads_kernel<<<blocks_exact((size_t)nd, 256), 256, 0, st>>>(skeys, nd, k, ctx->sc_gcount.as<uint32_t>(),
                                                                       ctx->sc_gstart.as<uint32_t>(), d_ngroups,
                                                                       (uint32_t)gcap);
    else
      rle_heads_kernel<<<blocks_exact((size_t)nd, 256), 256, 0, st>>>(scnt, (size_t)nd, ctx->sc_gcount.as<uint32_t>(),
                                                                     ctx->sc_gstart.as<uint32_t>(), d_ngroups,
                                                                     (uint32_t)gcap);
    LAUNCHED(1);
    uint32_t ng = 0;
    CK(cudaMemcpyAsync(&ng, d_ngroups, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ng > gcap) { gcap = ng; continue; }
    gcount.resize(ng);
    gstart32.resize(ng);
    CK(cudaMemcpyAsync(gcount.data(), ctx->sc_gcount.p, (size_t)ng * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(gstart32.data(), ctx->sc_gstart.p, (size_t)ng * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    break;
  }
  const size_t ng = gcount.size();
  {
    std::vector<uint32_t> ord(ng);
    for (size_t i = 0; i < ng; ++i) ord[i] = (uint32_t)i;
    std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return gstart32[a] < gstart32[b]; });
    std::vector<uint32_t> c2(ng), s2(ng);
    for (size_t i = 0; i < ng; ++i) { c2[i] = gcount[ord[i]]; s2[i] = gstart32[ord[i]]; }
    gcount.swap(c2);
    gstart32.swap(s2);
  }
  std::vector<uint64_t> gstart(ng + 1);
  for (size_t i = 0; i < ng; ++i) gstart[i] = gstart32[i];
  gstart[ng] = nd;
  std::vector<uint32_t> seg_first;
  std::vector<RankSeg> segs;
  build_rank_segments(gcount.data(), gstart.data(), ng, total, seg_first, segs);
  const size_t nsg = segs.size();
  std::vector<unsigned long long> j0(nsg);
  std::vector<double> x0(nsg), inc(nsg);
  for (size_t i = 0; i < nsg; ++i) { j0[i] = segs[i].j0; x0[i] = segs[i].x0; inc[i] = segs[i].inc; }
  gstart32.push_back((uint32_t)std::min<uint64_t>(nd, 0xffffffffull));
  // ---- scale of the exact scan, scores into the slots ----
  double wmax;
  if (mode == KS_MODE_RANK) {
    const double rmax = fma((double)(nd - 1 - gstart[ng - 1] - j0[nsg - 1]), inc[nsg - 1], x0[nsg - 1]);
    wmax = std::max(fabs(0.0 - thr), fabs(rmax - thr));
  } else {
    wmax = std::max(fabs(1.0 - thr), fabs(-1.0 - thr));
  }
  if (!(wmax < 0x1p40)) return ctx->fail(KS_ERR_RANGE, "threshold outside the exact scan range");
  const uint64_t mw = (uint64_t)(int64_t)min_width;
  DevScanParams hp;
  memset(&hp, 0, sizeof hp);
  hp.qs = qs_for_max(wmax);
  fx_t mu = fx_ceil_units(min_score, hp.qs);
  hp.min_width = mw;
  hp.min_lo = (uint64_t)(unsigned __int128)mu;
  hp.min_hi = (int64_t)(uint64_t)(((unsigned __int128)mu) >> 64);
  CK(ctx->prm.ensure(sizeof(DevScanParams)));
  CK(ctx->sc_gstart.ensure((ng + 1) * 4));
  CK(ctx->sc_segfirst.ensure((ng + 1) * 4));
  CK(ctx->sc_segj0.ensure(nsg * 8 + 8));
  CK(ctx->sc_segx0.ensure(nsg * 8 + 8));
  CK(ctx->sc_seginc.ensure(nsg * 8 + 8));
  const bool keep_ranks = nd <= (1ull << 28);  // the rank doubles in order are kept for inspection up to 2 GiB
  if (keep_ranks) CK(ctx->lg_ranks.ensure((size_t)nd * 8));
  else ctx->lg_ranks.release();
  CK(cudaMemcpyAsync(ctx->prm.p, &hp, sizeof hp, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->sc_gstart.p, gstart32.data(), (ng + 1) * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->sc_segfirst.p, seg_first.data(), (ng + 1) * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->sc_segj0.p, j0.data(), nsg * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->sc_segx0.p, x0.data(), nsg * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->sc_seginc.p, inc.data(), nsg * 8, cudaMemcpyHostToDevice, st));
  large_eval_kernel<<<blocks_exact((size_t)nd, 256), 256, 0, st>>>(
      skeys, sslots, sidx, scnt, nd, k, ctx->sc_gstart.as<uint32_t>(), (uint32_t)ng, ctx->sc_segfirst.as<uint32_t>(),
      ctx->sc_segj0.as<unsigned long long>(), ctx->sc_segx0.as<double>(), ctx->sc_seginc.as<double>(), mode, total,
      param, thr, hp.qs, ctx->lg_slots.as<HashSlot>(), keep_ranks ? ctx->lg_ranks.as<double>() : nullptr);
  LAUNCHED(1);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));  // host vectors above must outlive the copies
  ctx->prof_end(KS_PROF_SCORES, psc);
  ScanTable tab;
  tab.use_hash = true;
  return scan_core(ctx, s, k, tab, mw, nullptr, host_out, n_spans, nullptr);
  KS_CATCH(ctx)
}

// the sparse table of the last ks_dev_large_regions on this ctx, in (count, code) order: n_distinct entries each
int ks_large_table(ks_ctx *ctx, uint64_t *codes_out, uint32_t *counts_out, double *ranks_out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!ctx->lg_nd || !ctx->lg_sorted) return ctx->fail(KS_ERR_ARG, "ks_large_table: no large-k table on this context");
  CK(cudaSetDevice(ctx->device));
  const size_t nd = (size_t)ctx->lg_nd;
  std::vector<uint64_t> comp(nd);
  CK(cudaMemcpy(comp.data(), ctx->lg_sorted, nd * 8, cudaMemcpyDeviceToHost));
  const int sh = 2 * ctx->lg_k;
  const uint64_t kmask = (((uint64_t)1) << sh) - 1;
  if (ctx->lg_idx) {  // two-pass order: codes sit in code order, idx / cnt in (count, code) order
    std::vector<uint32_t> idx(nd), cnt(nd);
    CK(cudaMemcpy(idx.data(), ctx->lg_idx, nd * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cnt.data(), ctx->lg_cnt, nd * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nd; ++i) {
      if (codes_out) codes_out[i] = comp[idx[i]];
      if (counts_out) counts_out[i] = cnt[i];
    }
  } else {
    for (size_t i = 0; i < nd; ++i) {
      if (codes_out) codes_out[i] = comp[i] & kmask;
      if (counts_out) counts_out[i] = (uint32_t)(comp[i] >> sh);
    }
  }
  if (ranks_out) {
    if (!ctx->lg_ranks.p) return ctx->fail(KS_ERR_ARG, "ks_large_table: ranks are kept only up to 2^28 distinct k-mers");
    CK(cudaMemcpy(ranks_out, ctx->lg_ranks.p, nd * 8, cudaMemcpyDeviceToHost));
  }
  return KS_OK;
  KS_CATCH(ctx)
}

// host-buffer entry point of the large-k path
int ks_kmer_large_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k, int mode,
                          double param, double thr, int min_width, double min_score, double *n_words,
                          uint64_t *n_distinct, ks_spans *out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!seqs || !lens || nseq < 1)
    return ctx->fail(KS_ERR_ARG, "seq_r must be a character vector of length at least one");
  if (!out) return ctx->fail(KS_ERR_ARG, "null argument");
  CK(cudaSetDevice(ctx->device));
  ks_seqset *ss = nullptr;
  int rc = host_set_acquire(ctx, lens, nseq, &ss);
  if (rc) return rc;
  rc = upload_impl(ctx, ss, seqs, lens, nseq, 0, nullptr);
  if (rc) return rc;
  return ks_dev_large_regions(ctx, ss, k, mode, param, thr, min_width, min_score, n_words, n_distinct, out, nullptr);
  KS_CATCH(ctx)
}

// carry entering shard `rank`: fold of the aggregates of shards 0 .. rank-1 (host, exact integers)
int ks_fold_carry(int what, const void *all48, int nranks, int rank, void *carry_in48) {
  KS_TRY
  if (!all48 || !carry_in48 || rank < 0 || rank > nranks) return KS_ERR_ARG;
  memset(carry_in48, 0, 48);
  if (what == 0) {
    const XfRec *r = reinterpret_cast<const XfRec *>(all48);
    Xf f = xf_identity();
    for (int i = 0; i < rank; ++i) {
      Xf g; g.a = r[i].a; g.b = r[i].b; g.kill = r[i].kill;
      f = xf_compose(f, g);
    }
    fx_t S = xf_apply(f, 0);
    memcpy(carry_in48, &S, sizeof S);
    return KS_OK;
  }
  if (what == 1) {
    const ExRec *r = reinterpret_cast<const ExRec *>(all48);
    Ex acc = ex_identity();
    acc.reset = 1; acc.open = 0;  // left of the first shard nothing is open
    for (int i = 0; i < rank; ++i) {
      Ex g; g.M = r[i].M; g.beg = r[i].beg; g.pk = r[i].pk; g.reset = r[i].reset; g.open = r[i].open;
      acc = ex_combine(acc, g);
    }
    ExRec out;
    memset(&out, 0, sizeof out);
    out.M = acc.M; out.beg = acc.beg; out.pk = acc.pk; out.reset = 1; out.open = acc.open;
    memcpy(carry_in48, &out, sizeof out);
    return KS_OK;
  }
  return KS_ERR_ARG;
  KS_CATCH(((ks_ctx *)nullptr))
}

int ks_dev_pipeline(ks_ctx *ctx, const ks_seqset *s, int k, int mode, double param, double thr, int min_width,
                    double min_score, int32_t *d_counts, double *d_scores, double *n_words,
                    ks_spans *host_out, uint64_t *n_spans) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  double nw = 0;
  // the word count is read back together with the first table the score stage needs on the host
  int rc = dev_count_impl(ctx, s, k, d_counts, nullptr, false);
  if (rc) return rc;
  ctx->defer_table = getenv("KS_NO_SIDE_TABLE") == nullptr;
  rc = dev_scores_impl(ctx, k, d_counts, 0.0, true, mode, param, d_scores, &nw);
  ctx->defer_table = false;
  if (!rc) {
    if (n_words) *n_words = nw;
    if (mode == KS_MODE_LOG2 || mode == KS_MODE_SIGN)
      rc = ks_dev_scan_counts(ctx, s, k, d_counts, thr, min_width, min_score, host_out, n_spans);
    else if (mode == KS_MODE_RANK && ctx->rk_valid)  // the score stage just left the rank order on this context
      rc = ks_dev_scan_ranks(ctx, s, k, thr, min_width, min_score, host_out, n_spans);
    else
      rc = ks_dev_scan(ctx, s, k, d_scores, thr, min_width, min_score, nullptr, host_out, n_spans);
  }
  const int rcj = join_table(ctx);  // the score table written next to the scan: part of this call
  return rc ? rc : rcj;
  KS_CATCH(ctx)
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points
static int check_seqs(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq) {
  if (!seqs || !lens || nseq < 1)
    return ctx->fail(KS_ERR_ARG, "seq_r must be a character vector of length at least one");
  return KS_OK;
}

// upload (copy stream) with the pack+count pass running behind it (compute stream); returns the
// words counted.  The set is the context's cached one.
static int upload_and_count(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                            int32_t *d_counts, double *n_words, ks_seqset **set) {
  ks_seqset *ss = nullptr;
  int rc = host_set_acquire(ctx, lens, nseq, &ss);
  if (rc) return rc;
  rc = upload_impl(ctx, ss, seqs, lens, nseq, k, d_counts);
  if (rc) return rc;
  unsigned long long nw = 0;
  CK(cudaMemcpyAsync(&nw, ctx->nwords.p, sizeof nw, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n_words = (double)nw;
  *set = ss;
  return KS_OK;
}

// D2H of a table on the copy stream, behind everything the compute stream has done so far; it
// overlaps with the kernels that follow.  Completed by finish_copies().
static int copy_out_async(ks_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes) {
  CK(cudaEventRecord(ctx->ev_compute, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_compute, 0));
  if (host_ptr_is_pinned(host_dst) || bytes < (4u << 20) || !ctx->pinned) {
    CK(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    return KS_OK;
  }
  ctx->out_jobs.push_back({host_dst, dev_src, bytes});  // pageable destination: staged by start_staged_copies()
  return KS_OK;
}
// A device-to-host copy into pageable memory runs at a few GB/s; staged through the two pinned windows
// (the upload is done with them by now) with the final memcpy spread over host threads it runs near the
// PCIe rate, and in a helper thread it overlaps the scan.
static void start_staged_copies(ks_ctx *ctx) {
  if (ctx->out_jobs.empty()) return;
  ctx->out_rc = 0;
  ctx->out_thread = std::thread([ctx]() {
   try {
    cudaSetDevice(ctx->device);
    const size_t HALF = ctx->pinned_cap / 2;
    cudaEvent_t ev[2];
    cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    unsigned nthreads = std::thread::hardware_concurrency();
    nthreads = nthreads > 6 ? 6 : (nthreads < 1 ? 1 : nthreads);
    ctx->pool.ensure(nthreads);  // the main thread is in the scan by now and does not use the pool
    auto spread = [&](char *dst, const char *src, size_t len) {
      const unsigned n = ctx->pool.size();
      const size_t per = (len + n - 1) / n;
      ctx->pool.run([=](unsigned t) {
        const size_t a = (size_t)t * per;
        if (a >= len) return;
        const size_t b = a + per < len ? a + per : len;
        memcpy(dst + a, src + a, b - a);
      });
    };
    cudaError_t e = cudaSuccess;
    for (const auto &job : ctx->out_jobs) {
      size_t prev_off = 0, prev_len = 0;
      int half = 0;
      bool have_prev = false;
      for (size_t off = 0; off < job.bytes && e == cudaSuccess; off += HALF) {
        const size_t len = job.bytes - off < HALF ? job.bytes - off : HALF;
        char *win = (char *)ctx->pinned + (size_t)half * HALF;
        e = cudaMemcpyAsync(win, (const char *)job.src + off, len, cudaMemcpyDeviceToHost, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[half], ctx->copy_stream);
        if (have_prev && e == cudaSuccess) {  // the other window is complete: hand it over while this one fills
          e = cudaEventSynchronize(ev[half ^ 1]);
          if (e == cudaSuccess) spread((char *)job.dst + prev_off, (char *)ctx->pinned + (size_t)(half ^ 1) * HALF, prev_len);
        }
        prev_off = off; prev_len = len; have_prev = true;
        half ^= 1;
      }
      if (have_prev && e == cudaSuccess) {
        e = cudaEventSynchronize(ev[half ^ 1]);
        if (e == cudaSuccess) spread((char *)job.dst + prev_off, (char *)ctx->pinned + (size_t)(half ^ 1) * HALF, prev_len);
      }
      if (e != cudaSuccess) break;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (e != cudaSuccess) ctx->out_rc = (int)e;
   } catch (...) {  // nothing may escape a thread
    ctx->out_rc = (int)cudaErrorMemoryAllocation;
   }
  });
}
static int finish_copies(ks_ctx *ctx) {
  if (ctx->out_thread.joinable()) ctx->out_thread.join();
  ctx->out_jobs.clear();
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
  if (ctx->out_rc) {
    cudaError_t e = (cudaError_t)ctx->out_rc;
    ctx->out_rc = 0;
    return ctx->fail(KS_ERR_CUDA, "copy of a result table to the host failed: %s", cudaGetErrorString(e));
  }
  return KS_OK;
}

int ks_kmer_counts(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                   int32_t *counts_out, double *n_words) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  int rc = check_seqs(ctx, seqs, lens, nseq);
  if (rc) return rc;
  if (k < 1 || k > 15) return ctx->fail(KS_ERR_ARG, "k must be a positive integer less than 1+MAX_K");
  if (!counts_out || !n_words) return ctx->fail(KS_ERR_ARG, "null output");
  CK(cudaSetDevice(ctx->device));
  size_t n = (size_t)1 << (2 * k);
  CK(ctx->tmp_counts.ensure(n * 4));
  ks_seqset *ss = nullptr;
  rc = upload_and_count(ctx, seqs, lens, nseq, k, ctx->tmp_counts.as<int32_t>(), n_words, &ss);
  if (rc) return rc;
  rc = copy_out_async(ctx, counts_out, ctx->tmp_counts.p, n * 4);
  start_staged_copies(ctx);
  int rc2 = finish_copies(ctx);
  return rc ? rc : rc2;
  KS_CATCH(ctx)
}

int ks_kmer_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k, const double *W,
                    int min_width, double min_score, double *nuc, int32_t *inscan_counts_out, ks_spans *out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  int rc = check_seqs(ctx, seqs, lens, nseq);
  if (rc) return rc;
  if (k >= 16 || k < 1)
    return ctx->fail(KS_ERR_ARG, "kmer sizes larger than or equal to %d not currently supported", 16);
  if (!W || !out) return ctx->fail(KS_ERR_ARG, "null argument");
  CK(cudaSetDevice(ctx->device));
  size_t n = (size_t)1 << (2 * k);
  if (nuc) {
    *nuc = 0;
    for (int i = 0; i < nseq; ++i)
      if (lens[i] >= k) *nuc += (double)lens[i];  // :533-535
  }
  cudaStream_t st = ctx->stream;
  CK(ctx->tmp_scores.ensure(n * 8));
  CK(ctx->tmp_inscan.ensure(n * 4));
  CK(cudaMemcpyAsync(ctx->tmp_scores.p, W, n * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->tmp_inscan.p, 0, n * 4, st));
  ks_seqset *ss = nullptr;
  rc = host_set_acquire(ctx, lens, nseq, &ss);
  if (rc) return rc;
  rc = upload_impl(ctx, ss, seqs, lens, nseq, 0, nullptr);
  if (rc) return rc;
  rc = ks_dev_scan(ctx, ss, k, ctx->tmp_scores.as<double>(), 0.0, min_width, min_score,
                   inscan_counts_out ? ctx->tmp_inscan.as<int32_t>() : nullptr, out, nullptr);
  if (rc) return rc;
  if (inscan_counts_out) {
    rc = copy_out_async(ctx, inscan_counts_out, ctx->tmp_inscan.p, n * 4);
    start_staged_copies(ctx);
    int rc2 = finish_copies(ctx);
    return rc ? rc : rc2;
  }
  return KS_OK;
  KS_CATCH(ctx)
}

int ks_kmer_mode_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k, int mode,
                         double param, double thr, int min_width, double min_score, double *n_words,
                         int32_t *counts_out, double *scores_out, ks_spans *out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  int rc = check_seqs(ctx, seqs, lens, nseq);
  if (rc) return rc;
  rc = check_k(ctx, k);
  if (rc) return rc;
  if (!out) return ctx->fail(KS_ERR_ARG, "null argument");
  if (mode < KS_MODE_RANK || mode > KS_MODE_RANK_REL) return ctx->fail(KS_ERR_ARG, "unknown score mode %d", mode);
  CK(cudaSetDevice(ctx->device));
  size_t n = (size_t)1 << (2 * k);
  CK(ctx->tmp_counts.ensure(n * 4));
  CK(ctx->tmp_scores.ensure(n * 8));
  int32_t *d_counts = ctx->tmp_counts.as<int32_t>();
  double *d_scores = ctx->tmp_scores.as<double>();
  ks_seqset *ss = nullptr;
  double nw = 0;
  rc = upload_and_count(ctx, seqs, lens, nseq, k, d_counts, &nw, &ss);   // counting hides behind the H2D
  if (rc) return rc;
  if (n_words) *n_words = nw;
  const bool count_fn = (mode == KS_MODE_LOG2 || mode == KS_MODE_SIGN);
  rc = ks_dev_scores(ctx, k, d_counts, nw, mode, param, (count_fn && !scores_out) ? nullptr : d_scores);
  // the table copies go out only now: the score stage reads small tables back, and a 64 MiB copy in front
  // of them on the device-to-host engine would stall it for a millisecond; both copies overlap the scan
  if (!rc && counts_out) rc = copy_out_async(ctx, counts_out, d_counts, n * 4);
  if (!rc && scores_out) rc = copy_out_async(ctx, scores_out, d_scores, n * 8);
  start_staged_copies(ctx);
  if (!rc) {
    if (count_fn) rc = ks_dev_scan_counts(ctx, ss, k, d_counts, thr, min_width, min_score, out, nullptr);
    else if (mode == KS_MODE_RANK && ctx->rk_valid) rc = ks_dev_scan_ranks(ctx, ss, k, thr, min_width, min_score, out, nullptr);
    else rc = ks_dev_scan(ctx, ss, k, d_scores, thr, min_width, min_score, nullptr, out, nullptr);
  }
  int rc2 = finish_copies(ctx);
  return rc ? rc : rc2;
  KS_CATCH(ctx)
}

int ks_kmer_low_comp_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                             int min_width, double min_score, double thr, double n_out[2], int32_t *counts_out,
                             double *ranks_out, ks_spans *out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!(thr > 0 && thr < 1)) return ctx->fail(KS_ERR_ARG, "the threshold must be between 0 and 1");
  double nw = 0;
  int rc = ks_kmer_mode_regions(ctx, seqs, lens, nseq, k, KS_MODE_RANK, 0.0, thr, min_width, min_score, &nw,
                                counts_out, ranks_out, out);
  if (n_out) { n_out[0] = nw; n_out[1] = 0; }  // :613
  return rc;
  KS_CATCH(ctx)
}

// transition-score scan from host buffers
int ks_tr_lr_regions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq, int k,
                     const double *init_scores, const double *trans_scores, int min_length, ks_spans *out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  int rc = check_seqs(ctx, seqs, lens, nseq);
  if (rc) return rc;
  if (k < 1 || k > 15) return ctx->fail(KS_ERR_ARG, "k should be a positive value less than MAX_K");
  if (min_length < 0) return ctx->fail(KS_ERR_ARG, "min_length should be a positive integer");
  if (!init_scores || !trans_scores || !out) return ctx->fail(KS_ERR_ARG, "null argument");
  CK(cudaSetDevice(ctx->device));
  const size_t nk = (size_t)1 << (2 * k);
  cudaStream_t st = ctx->stream;
  CK(ctx->tr_tables.ensure(2 * nk * 8));
  double *d_init = ctx->tr_tables.as<double>(), *d_trans = d_init + nk;
  CK(cudaMemcpyAsync(d_init, init_scores, nk * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_trans, trans_scores, nk * 8, cudaMemcpyHostToDevice, st));
  ks_seqset *ss = nullptr;
  rc = host_set_acquire(ctx, lens, nseq, &ss);
  if (rc) return rc;
  rc = upload_impl(ctx, ss, seqs, lens, nseq, 0, nullptr);
  if (rc) return rc;
  return ks_dev_tr_lr_regions(ctx, ss, k, d_init, d_trans, min_length, out, nullptr);
  KS_CATCH(ctx)
}

// ------------------------------------------------------------------------------------------------
// windowed occurrence histograms
uint32_t ks_kmer_code(const char *s, int k) {
  uint32_t code = 0;
  if (!s) return 0;
  size_t i = 0;
  auto is_n = [](char c) { return (c | 0x20) == 'n'; };
  while (s[i]) {  // every N-free piece restarts the code; the first piece holding k bases ends the search
    code = 0;
    int got = 0;
    for (; got < k && s[i] && !is_n(s[i]); ++got, ++i) code = (code << 2) | (((unsigned char)s[i] >> 1) & 3u);
    if (got == k || !s[i]) break;
    while (s[i] && is_n(s[i])) ++i;
  }
  return code;
}

int ks_dev_window_dist(ks_ctx *ctx, const ks_seqset *s, int k, const uint32_t *codes, int kmer_n, int window,
                       int32_t *d_dist, int32_t *d_pos) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  if (!s || !codes || !d_dist) return ctx->fail(KS_ERR_ARG, "ks_dev_window_dist: null argument");
  if (k < 1 || k > 15) return ctx->fail(KS_ERR_ARG, "kmer sizes larger than or equal to %d not currently supported", 16);
  if (kmer_n < 1) return ctx->fail(KS_ERR_ARG, "kmers_r should be a character vector with at least one element");
  if (window < 2 * k) return ctx->fail(KS_ERR_ARG, "The window size must be at least two times k");
  if (s && s->window) return ctx->fail(KS_ERR_ARG, "ks_dev_window_dist: not for window sets");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int rc = ensure_packed(ctx, s);
  if (rc) return rc;
  const int64_t nch = s->total / 16;  // every chunk of the layout, front pad and tail included
  const int64_t mstride = nch + 2, pstride = nch + 1;
  const size_t bins = (size_t)window + 1;
  // selected k-mers are handled in batches that keep the per-chunk scratch below ~2 GiB
  int batch = (int)std::min<int64_t>(kmer_n, std::max<int64_t>(1, (2ll << 30) / (7 * nch + 1)));
  CK(ctx->win_match.ensure((size_t)batch * mstride * sizeof(uint16_t)));
  CK(ctx->win_cnt.ensure((size_t)(batch + 1) * nch));
  CK(ctx->win_pre.ensure((size_t)(batch + 1) * pstride * sizeof(uint32_t)));
  CK(ctx->win_scratch.ensure((size_t)(batch + 1) * exclusive_scan_scratch_elems((size_t)nch) * sizeof(uint32_t)));
  CK(ctx->win_codes.ensure((size_t)kmer_n * sizeof(uint32_t)));
  CK(cudaMemcpyAsync(ctx->win_codes.p, codes, (size_t)kmer_n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_dist, 0, bins * (size_t)kmer_n * sizeof(int32_t), st));
  if (d_pos) CK(cudaMemsetAsync(d_pos, 0, (size_t)kmer_n * (size_t)s->total * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ctx->win_match.p, 0, (size_t)batch * mstride * sizeof(uint16_t), st));  // spare columns stay 0
  // sequences exactly `window` long are left out by the reference (:775) but hold one window here
  std::vector<int64_t> fix;
  for (int q = 0; q < s->nseq; ++q)
    if (s->lens[q] == window) fix.push_back(s->starts[q]);
  if (!fix.empty()) {
    CK(ctx->win_fix.ensure(fix.size() * sizeof(int64_t)));
    CK(cudaMemcpyAsync(ctx->win_fix.p, fix.data(), fix.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  }
  uint8_t *d_cnt = ctx->win_cnt.as<uint8_t>();
  uint32_t *d_pre = ctx->win_pre.as<uint32_t>();
  uint8_t *d_brk_cnt = d_cnt + (size_t)batch * nch;
  uint32_t *d_brk_pre = d_pre + (size_t)batch * pstride;
  const uint32_t kmask = (uint32_t)(((uint64_t)1 << (2 * k)) - 1);
  for (int lo = 0; lo < kmer_n; lo += batch) {
    const int nb = std::min(batch, kmer_n - lo);
    win_match_kernel<<<grid_for((size_t)nch, WIN_THREADS, 148u * 8u), WIN_THREADS, (size_t)nb * sizeof(uint32_t), st>>>(
        s->d_pk, s->d_brk, nch, k, kmask, ctx->win_codes.as<uint32_t>() + lo, nb, ctx->win_match.as<uint16_t>(),
        mstride, d_cnt, lo == 0 ? d_brk_cnt : nullptr);
    LAUNCHED(1);
    // the break row sits right behind the k-mer rows of a full batch: one batched scan covers it too
    const bool with_brk = lo == 0;
    const int rows = nb + ((with_brk && nb == batch) ? 1 : 0);
    int nl = exclusive_scan_rows<uint8_t, uint32_t>(d_cnt, (size_t)nch, rows, (size_t)nch, d_pre, (size_t)pstride,
                                                    ctx->win_scratch.as<uint32_t>(), st);
    if (with_brk && nb != batch)
      nl += exclusive_scan<uint8_t, uint32_t>(d_brk_cnt, (size_t)nch, d_brk_pre, ctx->win_scratch.as<uint32_t>(), st);
    LAUNCHED(nl);
    WinArgs A;
    A.match = ctx->win_match.as<uint16_t>();
    A.pre = d_pre;
    A.brk = s->d_brk;
    A.brk_pre = d_brk_pre;
    A.mstride = mstride;
    A.pstride = pstride;
    A.nch = nch;
    A.k = k;
    A.window = window;
    A.kmer_n = nb;
    A.hist = d_dist + (size_t)lo * bins;
    A.pos = d_pos ? d_pos + (size_t)lo * (size_t)s->total : nullptr;
    A.pos_stride = s->total;
    // a CTA handles WIN_KB selected k-mers (gridDim.y batches): its bins stay small enough for 4+ CTAs per SM
    size_t smem = bins * (size_t)std::min(nb, WIN_KB) * sizeof(int32_t);
    A.use_smem = smem <= 160u * 1024u;
    if (A.use_smem && smem > 48u * 1024u)
      CK(cudaFuncSetAttribute(win_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 wgrid(grid_for((size_t)nch, WIN_THREADS, 148u * 4u), (unsigned)((nb + WIN_KB - 1) / WIN_KB));
    if (A.use_smem) win_hist_kernel<true><<<wgrid, WIN_THREADS, smem, st>>>(A);
    else win_hist_kernel<false><<<wgrid, WIN_THREADS, 0, st>>>(A);
    LAUNCHED(1);
    if (!fix.empty()) {
      int nt = (int)fix.size() * nb;
      win_fix_kernel<<<(nt + 127) / 128, 128, 0, st>>>(A, ctx->win_fix.as<int64_t>(), (int)fix.size());
      LAUNCHED(1);
    }
    CK(cudaGetLastError());
  }
  return KS_OK;
  KS_CATCH(ctx)
}

int ks_windowed_kmer_count_distributions(ks_ctx *ctx, const char *const *seqs, const int64_t *lens, int nseq,
                                         int k, const uint32_t *codes, int kmer_n, int window,
                                         int32_t *dist_out, int32_t *included_out, int32_t *const *pos_out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  int rc = check_seqs(ctx, seqs, lens, nseq);
  if (rc) return rc;
  if (!codes || kmer_n < 1) return ctx->fail(KS_ERR_ARG, "kmers_r should be a character vector with at least one element");
  if (k < 1 || k > 15) return ctx->fail(KS_ERR_ARG, "kmer sizes larger than or equal to %d not currently supported", 16);
  if (window < 2 * k) return ctx->fail(KS_ERR_ARG, "The window size must be at least two times k");
  if (!dist_out || !included_out) return ctx->fail(KS_ERR_ARG, "null output");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // sequences that are left out (:775) are uploaded as empty strings
  std::vector<int64_t> eff(lens, lens + nseq);
  bool any = false;
  for (int q = 0; q < nseq; ++q) {
    included_out[q] = lens[q] > window;
    if (!included_out[q]) eff[q] = 0;
    else any = true;
  }
  const size_t bins = (size_t)window + 1;
  memset(dist_out, 0, bins * (size_t)kmer_n * sizeof(int32_t));
  if (!any) return KS_OK;
  ks_seqset *ss = nullptr;
  rc = host_set_acquire(ctx, eff.data(), nseq, &ss);
  if (rc) return rc;
  rc = upload_impl(ctx, ss, seqs, eff.data(), nseq, 0, nullptr);
  if (rc) return rc;
  CK(ctx->win_hist.ensure(bins * (size_t)kmer_n * sizeof(int32_t)));
  int32_t *d_hist = ctx->win_hist.as<int32_t>();
  if (!pos_out) {
    rc = ks_dev_window_dist(ctx, ss, k, codes, kmer_n, window, d_hist, nullptr);
    if (rc) return rc;
  } else {
    // per-position values: as many selected k-mers per round as fit ~2 GiB of int32 rows
    int per = (int)std::min<int64_t>(kmer_n, std::max<int64_t>(1, (2ll << 30) / (4 * ss->total)));
    CK(ctx->win_pos.ensure((size_t)per * (size_t)ss->total * sizeof(int32_t)));
    int32_t *d_pos = ctx->win_pos.as<int32_t>();
    for (int lo = 0; lo < kmer_n; lo += per) {
      const int nb = std::min(per, kmer_n - lo);
      rc = ks_dev_window_dist(ctx, ss, k, codes + lo, nb, window, d_hist + (size_t)lo * bins, d_pos);
      if (rc) return rc;
      for (int q = 0; q < nseq; ++q) {
        if (!included_out[q] || !pos_out[q]) continue;
        for (int i = 0; i < nb; ++i)
          CK(cudaMemcpyAsync(pos_out[q] + (size_t)(lo + i) * (size_t)lens[q],
                             d_pos + (size_t)i * (size_t)ss->total + ss->starts[q], (size_t)lens[q] * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, st));
      }
      CK(cudaStreamSynchronize(st));
    }
  }
  CK(cudaMemcpyAsync(dist_out, d_hist, bins * (size_t)kmer_n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return KS_OK;
  KS_CATCH(ctx)
}

int ks_kmer_scores(ks_ctx *ctx, int k, const int32_t *counts, double total, int mode, double param,
                   double *scores_out) {
  KS_TRY
  if (!ctx) return KS_ERR_ARG;
  int rc = check_k(ctx, k);
  if (rc) return rc;
  if (!counts || !scores_out) return ctx->fail(KS_ERR_ARG, "null argument");
  CK(cudaSetDevice(ctx->device));
  size_t n = (size_t)1 << (2 * k);
  CK(ctx->tmp_counts.ensure(n * 4));
  CK(ctx->tmp_scores.ensure(n * 8));
  CK(cudaMemcpyAsync(ctx->tmp_counts.p, counts, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  rc = ks_dev_scores(ctx, k, ctx->tmp_counts.as<int32_t>(), total, mode, param, ctx->tmp_scores.as<double>());
  if (rc) return rc;
  CK(cudaMemcpyAsync(scores_out, ctx->tmp_scores.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return KS_OK;
  KS_CATCH(ctx)
}

}  // extern "C"

#include "ks_multi.inc"

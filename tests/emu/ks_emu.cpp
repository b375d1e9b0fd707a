// ks_emu.cpp -- HOST EMULATION of the GPU algorithm, TEST INFRASTRUCTURE ONLY.
//
// It drives the very same per-chunk functions the sm_100a kernels use (csrc/ks_chunk.cuh,
// csrc/ks_rankseg.h), folding chunks left to right where the kernels use warp/block scans and a
// decoupled look-back.  Because the scan arithmetic is exact (integer), the fold order does not
// matter, so the emulation must agree with the GPU bit for bit.  It lets the CPU-only test tier
// check the level-wise restatement (chunk transforms, excursion carry, restart-at-peak child
// segments, exact rank pieces) against the oracle without a GPU.  The product never loads it.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "../../kmer_spans_b200/csrc/ks_chunk.cuh"
#include "../../kmer_spans_b200/csrc/ks_pairgeom.h"
#include "../../kmer_spans_b200/csrc/ks_layout.h"
#include "../../kmer_spans_b200/csrc/ks_rankseg.h"

using namespace ks;

// K1 on the host: pack the whole buffer (chunk c = positions [16c, 16c+16))
struct Packed {
  std::vector<uint32_t> pk;
  std::vector<uint16_t> brk, nul;
};
static Packed pack_buffer(const uint8_t *buf, int64_t ntot) {
  Packed P;
  int64_t nch = ntot / 16 + 4;
  P.pk.assign(nch, 0); P.brk.assign(nch, 0xffff); P.nul.assign(nch, 0xffff);
  for (int64_t c = 0; c < ntot / 16; ++c) {
    uint32_t w[4], pk, b, n;
    memcpy(w, buf + 16 * c, 16);
    pack16(w, pk, b, n);
    P.pk[c] = pk; P.brk[c] = (uint16_t)b; P.nul[c] = (uint16_t)n;
  }
  return P;
}
// 32 positions [p0-16, p0+16) at an arbitrary p0 >= 16, as the kernels assemble them
static void window(const Packed &P, int64_t p0, uint64_t &X, uint32_t &brk32, uint32_t &nul32) {
  int64_t wq = p0 >> 4;
  int r = (int)(p0 & 15);
  uint32_t hi = P.pk[wq - 1], mid = P.pk[wq], lo = P.pk[wq + 1];
  uint32_t xh = r ? ((hi << (2 * r)) | (mid >> (32 - 2 * r))) : hi;
  uint32_t xl = r ? ((mid << (2 * r)) | (lo >> (32 - 2 * r))) : mid;
  X = ((uint64_t)xh << 32) | xl;
  uint64_t b48 = (uint64_t)P.brk[wq - 1] | ((uint64_t)P.brk[wq] << 16) | ((uint64_t)P.brk[wq + 1] << 32);
  uint64_t n48 = (uint64_t)P.nul[wq - 1] | ((uint64_t)P.nul[wq] << 16) | ((uint64_t)P.nul[wq + 1] << 32);
  brk32 = (uint32_t)(b48 >> r);
  nul32 = (uint32_t)(n48 >> r);
}

// the bucketed count of ks_count.cuh folded on the host: one sub-key per pair of consecutive counted k-mers, filed
// under the bucket PairGeom names; phase 2 = two tables per bucket (c.b in the bucket's slice of the count table,
// a.c in a second table), then the fold.  Un-paired k-mers, and pairs beyond `row_cap` in their bucket (the staging
// rows of the kernel are finite), are counted directly, decoded back from (bucket, sub-key) as the overflow path does.
template <int K>
static void count_pairs_k(const uint8_t *buf, int64_t ntot, int32_t *counts, uint64_t *nwords, uint32_t row_cap) {
  typedef PairGeom<K> G;
  const uint32_t kmask = G::KMASK;
  const size_t nk = (size_t)kmask + 1;
  std::vector<uint32_t> table_a(nk, 0), fill(G::NB, 0);
  uint64_t n = 0;
  Packed P = pack_buffer(buf, ntot);
  for (int64_t p0 = 16; p0 < ntot; p0 += 16) {
    uint32_t code[16], counted, brk32, nul32;
    uint64_t X;
    window(P, p0, X, brk32, nul32);
    uint32_t next_nul = (p0 + 16 < ntot) ? (buf[p0 + 16] == 0) : 1;
    decode_count(X, brk32, nul32, next_nul, K, kmask, code, counted);
    n += (uint64_t)__builtin_popcount(counted);
    for (int i = 0; i < 8; ++i) {
      const uint32_t m2 = (counted >> (2 * i)) & 3u;
      if (m2 == 3u) {
        const uint32_t y = (uint32_t)(X >> (28 - 4 * i));  // a.c.b, as the kernel reads it from the packed window
        const uint32_t bucket = G::bucket4(y) >> 2, sub = G::sub(y);
        if (G::code_ac(bucket, sub) != code[2 * i] || G::code_cb(bucket, sub) != code[2 * i + 1]) { *nwords = ~0ull; return; }
        if (sub == 0xffffu || fill[bucket] >= row_cap) {  // filler look-alike / full row: directly
          counts[G::code_ac(bucket, sub)]++;
          counts[G::code_cb(bucket, sub)]++;
        } else {
          ++fill[bucket];
          counts[(size_t)bucket * G::ENTRIES + (sub & G::LOW)]++;      // tabB -> the bucket's slice
          table_a[(size_t)bucket * G::ENTRIES + (sub >> 2)]++;         // tabA -> second table
        }
      } else {
        if (m2 & 1u) counts[code[2 * i]]++;
        if (m2 & 2u) counts[code[2 * i + 1]]++;
      }
    }
    if ((p0 & 0xfff0) == 0) std::fill(fill.begin(), fill.end(), 0u);  // a new "tile": the rows are empty again
  }
  for (size_t x = 0; x < nk; ++x) counts[x] += (int32_t)table_a[G::fold_index((uint32_t)x)];
  *nwords = n;
}
extern "C" {

int64_t emu_layout_total(const int64_t *lens, int nseq, int64_t *starts) {
  return ks_layout_total(lens, nseq, starts);
}

// counting pass over the whole buffer
void emu_count(const uint8_t *buf, int64_t ntot, int k, int32_t *counts, uint64_t *nwords) {
  uint32_t kmask = (1u << (2 * k)) - 1u;
  uint64_t n = 0;
  Packed P = pack_buffer(buf, ntot);
  for (int64_t p0 = 16; p0 < ntot; p0 += 16) {
    uint32_t code[16], counted, brk32, nul32;
    uint64_t X;
    window(P, p0, X, brk32, nul32);
    uint32_t next_nul = (p0 + 16 < ntot) ? (buf[p0 + 16] == 0) : 1;
    decode_count(X, brk32, nul32, next_nul, k, kmask, code, counted);
    for (int j = 0; j < 16; ++j)
      if (counted & (1u << j)) { counts[code[j]]++; ++n; }
  }
  *nwords = n;
}

int emu_count_pairs(const uint8_t *buf, int64_t ntot, int k, int32_t *counts, uint64_t *nwords, uint32_t row_cap) {
  switch (k) {
    case 8: count_pairs_k<8>(buf, ntot, counts, nwords, row_cap); return 0;
    case 9: count_pairs_k<9>(buf, ntot, counts, nwords, row_cap); return 0;
    case 10: count_pairs_k<10>(buf, ntot, counts, nwords, row_cap); return 0;
    case 11: count_pairs_k<11>(buf, ntot, counts, nwords, row_cap); return 0;
    case 12: count_pairs_k<12>(buf, ntot, counts, nwords, row_cap); return 0;
    case 13: count_pairs_k<13>(buf, ntot, counts, nwords, row_cap); return 0;
  }
  return 1;
}

// exact ranks: stable order, run-length groups, linear pieces, closed-form evaluation
void emu_rank_exact(const int32_t *counts, int k, double total, double *ranks) {
  size_t n = (size_t)1 << (2 * k);
  if (total == 0) {
    for (size_t i = 0; i < n; ++i) ranks[i] = NAN;
    ranks[0] = 0;
    return;
  }
  std::vector<uint32_t> idx(n);
  for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return counts[a] < counts[b]; });
  std::vector<uint32_t> gcount;
  std::vector<uint64_t> gstart;
  for (size_t i = 0; i < n; ++i)
    if (i == 0 || counts[idx[i]] != counts[idx[i - 1]]) { gcount.push_back((uint32_t)counts[idx[i]]); gstart.push_back(i); }
  gstart.push_back(n);
  std::vector<uint32_t> seg_first;
  std::vector<RankSeg> segs;
  build_rank_segments(gcount.data(), gstart.data(), gcount.size(), total, seg_first, segs);
  size_t g = 0;
  for (size_t p = 0; p < n; ++p) {
    while (gstart[g + 1] <= p) ++g;
    uint64_t j = p - gstart[g];
    uint32_t lo = seg_first[g], hi = seg_first[g + 1];
    // last piece with j0 <= j
    while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (segs[mid].j0 <= j) lo = mid; else hi = mid; }
    ranks[idx[p]] = fma((double)(j - segs[lo].j0), segs[lo].inc, segs[lo].x0);
  }
}

int emu_num_rank_segments(const int32_t *counts, int k, double total) {
  size_t n = (size_t)1 << (2 * k);
  std::vector<int32_t> c(counts, counts + n);
  std::sort(c.begin(), c.end());
  std::vector<uint32_t> gcount;
  std::vector<uint64_t> gstart;
  for (size_t i = 0; i < n; ++i)
    if (i == 0 || c[i] != c[i - 1]) { gcount.push_back((uint32_t)c[i]); gstart.push_back(i); }
  gstart.push_back(n);
  std::vector<uint32_t> seg_first;
  std::vector<RankSeg> segs;
  build_rank_segments(gcount.data(), gstart.data(), gcount.size(), total, seg_first, segs);
  return (int)segs.size();
}

struct Rec { int64_t beg, pk, c; fx_t M; };
struct Emit {
  std::vector<Rec> *out;
  void operator()(int64_t beg, int64_t pk, int64_t c, fx_t M) { out->push_back({beg, pk, c, M}); }
};

// Returns 0, or 1 when a weight is +inf / >= 2^40 (rejected by the product too).
// Outputs (malloc'ed, caller frees with emu_free): beg, pk (global positions), score.
// fast != 0 (and min_width >= 15): the summary-based walk of scan_walk_fast_kernel / scan_detail_kernel --
// chunk summaries from FastChunk / GeneralChunk, open-excursion elements from fast_walk_element, and the
// position-by-position walk only for chunks whose entering excursion closes and might qualify.
static int emu_scan_impl(const uint8_t *buf, int64_t ntot, int k, const double *W, double thr, uint64_t min_width,
                         double min_score, int32_t *inscan, int64_t *n_out, int64_t **beg_out, int64_t **pk_out,
                         double **score_out, int *levels_out, int64_t *revisits_out, int fast, int64_t *detail_out);
int emu_scan(const uint8_t *buf, int64_t ntot, int k, const double *W, double thr, uint64_t min_width,
             double min_score, int32_t *inscan, int64_t *n_out, int64_t **beg_out, int64_t **pk_out,
             double **score_out, int *levels_out, int64_t *revisits_out) {
  return emu_scan_impl(buf, ntot, k, W, thr, min_width, min_score, inscan, n_out, beg_out, pk_out, score_out,
                       levels_out, revisits_out, 0, nullptr);
}
// detail_out: number of chunks that needed the position-by-position walk
int emu_scan_fast(const uint8_t *buf, int64_t ntot, int k, const double *W, double thr, uint64_t min_width,
                  double min_score, int64_t *n_out, int64_t **beg_out, int64_t **pk_out, double **score_out,
                  int *levels_out, int64_t *detail_out) {
  return emu_scan_impl(buf, ntot, k, W, thr, min_width, min_score, nullptr, n_out, beg_out, pk_out, score_out,
                       levels_out, nullptr, 1, detail_out);
}
static int emu_scan_impl(const uint8_t *buf, int64_t ntot, int k, const double *W, double thr, uint64_t min_width,
                         double min_score, int32_t *inscan, int64_t *n_out, int64_t **beg_out, int64_t **pk_out,
                         double **score_out, int *levels_out, int64_t *revisits_out, int fast, int64_t *detail_out) {
  if (fast && min_width < 15) fast = 0;
  int64_t n_detail = 0;
  size_t nk = (size_t)1 << (2 * k);
  uint32_t kmask = (uint32_t)(nk - 1);
  // table: w = fl(W - thr) as the reference computes it (:268), then exact fixed point
  double wmax = 0;
  for (size_t i = 0; i < nk; ++i) {
    double w = W[i] - thr;
    if (w >= 0x1p40) return 1;
    if (w == w && w > -0x1p40 && fabs(w) > wmax) wmax = fabs(w);
  }
  int qs = qs_for_max(wmax);
  std::vector<int64_t> wfx(nk);
  for (size_t i = 0; i < nk; ++i) wfx[i] = wfx_from_double(W[i] - thr, qs);
  ScanParams prm;
  prm.min_width = min_width;
  prm.min_units = fx_ceil_units(min_score, qs);

  Packed P = pack_buffer(buf, ntot);
  std::vector<Rec> all;
  std::vector<std::pair<int64_t, int64_t>> segs;  // start, len
  segs.push_back({16, ntot - 16});
  int level = 0;
  int64_t revisits = 0;
  while (!segs.empty()) {
    std::vector<Rec> recs;
    Emit emit{&recs};
    for (auto &sg : segs) {
      int64_t nchunks = (level == 0) ? sg.second / 16 : segment_chunks(sg.second);
      fx_t S = 0;
      Ex E = ex_identity();
      E.reset = 1; E.open = 0;
      if (fast && level == 0 && min_width >= 31 && !getenv("KS_NO_PAIR")) {
        // units of two chunks (scan_gather_kernel<..., kPair>): one merged summary per 32 positions
        for (int64_t u = 0; u < (nchunks + 1) / 2; ++u) {
          const int64_t p0u = sg.first + 32 * u;
          UnitSummary U;
          int64_t sv[2][16];
          uint32_t lv[2] = {0, 0};
          int nsub = 0;
          for (int h = 0; h < 2; ++h) {
            const int64_t ci = 2 * u + h;
            if (ci >= nchunks) break;
            ++nsub;
            const int64_t p0 = p0u + 16 * h;
            uint32_t code[16], scored, brk32, nul32;
            uint64_t X;
            window(P, p0, X, brk32, nul32);
            decode_scan(X, brk32, k, kmask, 16, code, scored);
            for (int j = 0; j < 16; ++j) {
              sv[h][j] = 0;
              if (scored & (1u << j)) {
                int64_t v = wfx[code[j]];
                if (v != WFX_KILL) { sv[h][j] = v; lv[h] |= 1u << j; }
              }
            }
            ChunkSummary sm;
            UnitSummary uc;
            bool general = scored != 0xffffu;
            if (!general) {
              FastChunk fc;
              fc.init();
              for (int j = 0; j < 16; ++j) fc.step(j, wfx[code[j]]);
              general = fc.bad();
              if (!fc.bad()) { sm = fc.summary(); uc = unit_from_chunk(fc.a(), fc.b(), 0, 0xffffu, sm); }
            }
            if (general) {
              GeneralChunk gc;
              gc.init();
              for (int j = 0; j < 16; ++j) gc.step<true>(j, (scored & (1u << j)) ? wfx[code[j]] : WFX_KILL);
              if (gc.live != lv[h]) return 2;
              sm = gc.summary();
              uc = unit_from_chunk(gc.ta, gc.tb, gc.kill, gc.live, sm);
            }
            U = h == 0 ? uc : unit_merge(U, uc);
          }
          const bool head = u == 0;
          if (head) unit_make_head(U);
          Ex ex;
          bool closing = false;
          fast_walk_unit(S, head, U, p0u, ex, closing);
          // reference: the two chunks walked position by position with their true entering states
          fx_t Sc = S;
          Ex ex_ref = ex_identity();
          bool zero_ref = false;
          {
            std::vector<Rec> scratch;
            Emit none{&scratch};
            for (int h = 0; h < nsub; ++h) {
              Ex e1;
              fx_t pm;
              int64_t pp;
              int z;
              chunk_walk(sv[h], lv[h], Sc, p0u + 16 * h, prm, none, e1, pm, pp, z);
              ex_ref = h == 0 ? e1 : ex_combine(ex_ref, e1);
              zero_ref = zero_ref || z >= 0;
              Sc = xf_apply(chunk_transform(sv[h], lv[h]), Sc);
            }
            if (!scratch.empty()) return 7;  // an excursion inside 32 positions cannot qualify
          }
          if (ex_ref.reset != ex.reset || ex_ref.open != ex.open || ex_ref.beg != ex.beg || ex_ref.pk != ex.pk ||
              ex_ref.M != ex.M)
            return 5;
          if (closing != (!head && S > 0 && zero_ref)) return 6;
          Xf fu;
          fu.a = (fx_t)U.ta; fu.b = (fx_t)U.tb; fu.kill = U.kill;
          if (xf_apply(fu, S) != Sc) return 8;  // merged transform == the chunks' transforms chained
          if (closing && !fast_walk_cannot_qualify(E, S, p0u, U.mx, prm, 32)) {
            ++n_detail;  // scan_detail_kernel<..., kPair>: first chunk, then the second if no zero was met
            fx_t Sd = S, accM = -(((fx_t)1) << 126);
            int64_t accPk = -1;
            int fz = -1;
            int64_t pz = 0;
            for (int h = 0; h < nsub && fz < 0; ++h) {
              Ex exd;
              fx_t preM;
              int64_t prePk;
              chunk_walk(sv[h], lv[h], Sd, p0u + 16 * h, prm, emit, exd, preM, prePk, fz);
              if (preM > accM) { accM = preM; accPk = prePk; }
              pz = p0u + 16 * h;
              if (fz < 0) { for (int j = 0; j < 16; ++j) Sd += (fx_t)sv[h][j]; }
            }
            if (fz < 0) return 3;
            chunk_finish_entering(S, E, accM, accPk, fz, pz, prm, emit);
          }
          S = Sc;
          E = ex_combine(E, ex);
        }
        continue;
      }
      for (int64_t ci = 0; ci < nchunks; ++ci) {
        int64_t p0 = sg.first + 16 * ci;
        int64_t rem = sg.first + sg.second - p0;
        int n_in = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        uint32_t code[16], scored, brk32, nul32;
        uint64_t X;
        window(P, p0, X, brk32, nul32);
        decode_scan(X, brk32, k, kmask, n_in, code, scored);
        int64_t s[16];
        uint32_t live = 0;
        for (int j = 0; j < 16; ++j) {
          s[j] = 0;
          if (scored & (1u << j)) {
            if (inscan) inscan[code[j]]++;
            if (level > 0) ++revisits;
            int64_t v = wfx[code[j]];
            if (v != WFX_KILL) { s[j] = v; live |= 1u << j; }
          }
        }
        Xf f;
        Ex ex;
        fx_t preM;
        int64_t prePk;
        int fz;
        if (fast) {
          // what the gather kernel leaves behind for this chunk
          ChunkSummary sm;
          bool general = scored != 0xffffu;
          uint32_t live2 = 0;
          if (!general) {
            FastChunk fc;
            fc.init();
            for (int j = 0; j < 16; ++j) fc.step(j, wfx[code[j]]);
            general = fc.bad();
            if (!fc.bad()) { f.a = (fx_t)fc.a(); f.b = (fx_t)fc.b(); f.kill = 0; sm = fc.summary(); live2 = 0xffffu; }
          }
          if (general) {
            GeneralChunk gc;
            gc.init();
            for (int j = 0; j < 16; ++j) gc.step<true>(j, (scored & (1u << j)) ? wfx[code[j]] : WFX_KILL);
            f.a = (fx_t)gc.ta; f.b = (fx_t)gc.tb; f.kill = gc.kill; sm = gc.summary(); live2 = gc.live;
          }
          if (live2 != live) return 2;  // the two formulations must see the same live positions
          // what the walk kernel derives from it
          bool closing = false;
          fast_walk_element(S, ci == 0, live, sm, p0, ex, closing);
          {  // every summary-derived element must equal the one the position-by-position walk finds
            std::vector<Rec> scratch;
            Emit none{&scratch};
            Ex ex_ref;
            fx_t pm;
            int64_t pp;
            int z;
            chunk_walk(s, live, S, p0, prm, none, ex_ref, pm, pp, z);
            if (ex_ref.reset != ex.reset || ex_ref.open != ex.open || ex_ref.beg != ex.beg || ex_ref.pk != ex.pk ||
                ex_ref.M != ex.M)
              return 5;
            if (closing != (ci != 0 && S > 0 && z >= 0)) return 6;
          }
          if (closing && !fast_walk_cannot_qualify(E, S, p0, sm.mx, prm)) {
            ++n_detail;  // scan_detail_kernel: position-by-position walk of this chunk only
            Ex ex_detail;
            chunk_walk(s, live, S, p0, prm, emit, ex_detail, preM, prePk, fz);
            if (fz < 0) return 3;
            chunk_finish_entering(S, E, preM, prePk, fz, p0, prm, emit);
            if (ex_detail.reset != ex.reset || ex_detail.open != ex.open || ex_detail.beg != ex.beg ||
                ex_detail.pk != ex.pk || ex_detail.M != ex.M)
              return 4;  // summary-derived element must equal the walked one
          }
        } else {
          f = chunk_transform(s, live);
          chunk_walk(s, live, S, p0, prm, emit, ex, preM, prePk, fz);
          chunk_finish_entering(S, E, preM, prePk, fz, p0, prm, emit);
        }
        S = xf_apply(f, S);
        E = ex_combine(E, ex);
      }
    }
    segs.clear();
    for (auto &r : recs) {
      int64_t st, ln;
      if (child_segment(r.pk, r.c, min_width, inscan != nullptr, st, ln)) segs.push_back({st, ln});
      all.push_back(r);
    }
    ++level;
    if (level > 100000) break;
  }
  std::sort(all.begin(), all.end(), [](const Rec &a, const Rec &b) { return a.beg < b.beg; });
  int64_t n = (int64_t)all.size();
  *n_out = n;
  *beg_out = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
  *pk_out = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
  *score_out = (double *)malloc(sizeof(double) * (n + 1));
  for (int64_t i = 0; i < n; ++i) {
    (*beg_out)[i] = all[i].beg;
    (*pk_out)[i] = all[i].pk;
    (*score_out)[i] = fx_to_double(all[i].M, qs);
  }
  if (levels_out) *levels_out = level;
  if (revisits_out) *revisits_out = revisits;
  if (detail_out) *detail_out = n_detail;
  return 0;
}

// ---- transition-score scan (tr_lr_regions_r): the level loop over chunk_walk_tr, folded on the host ----
struct TrEmit {
  std::vector<Rec> *regions;
  std::vector<std::pair<int64_t, int64_t>> *children;  // (start, len) of the re-scans asked for
  void out(int64_t beg, int64_t pk, fx_t M) { regions->push_back({beg, pk, 0, M}); }
  void child(int64_t pk, int64_t c, uint64_t min_width) {
    if (tr_child_wanted(pk, c, min_width)) children->push_back({pk + 1, c - pk});
  }
};

// init / trans: double[4^k] in 2-bit code order.  Outputs as emu_scan, coordinates 0-based global positions
// (the caller adds the 1-based convention).  Returns 1 for +inf / >= 2^40, 2 for NaN entries.
int emu_tr_scan(const uint8_t *buf, int64_t ntot, int k, const double *init, const double *trans, int min_len,
                int64_t *n_out, int64_t **beg_out, int64_t **pk_out, double **score_out, int *levels_out) {
  size_t nk = (size_t)1 << (2 * k);
  uint32_t kmask = (uint32_t)(nk - 1);
  double wmax = 0;
  for (int t = 0; t < 2; ++t) {
    const double *W = t ? init : trans;
    for (size_t i = 0; i < nk; ++i) {
      double w = W[i];
      if (w != w) return 2;
      if (w >= 0x1p40) return 1;
      if (w > -0x1p40 && fabs(w) > wmax) wmax = fabs(w);
    }
  }
  int qs = qs_for_max(wmax);
  std::vector<int64_t> wfx(2 * nk);  // [trans | init], as the kernel's table
  for (size_t i = 0; i < nk; ++i) { wfx[i] = wfx_from_double(trans[i], qs); wfx[nk + i] = wfx_from_double(init[i], qs); }
  ScanParams prm;
  prm.min_width = (uint64_t)min_len;
  prm.min_units = fx_ceil_units(-INFINITY, qs);

  Packed P = pack_buffer(buf, ntot);
  std::vector<Rec> all;
  std::vector<std::pair<int64_t, int64_t>> segs, next;
  segs.push_back({16, ntot - 16});
  int level = 0;
  while (!segs.empty()) {
    next.clear();
    TrEmit emit{&all, &next};
    for (auto &sg : segs) {
      int64_t nchunks = (level == 0) ? sg.second / 16 : segment_chunks(sg.second);
      fx_t S = 0;
      Ex E = ex_identity();
      E.reset = 1; E.open = 0;
      for (int64_t ci = 0; ci < nchunks; ++ci) {
        int64_t p0 = sg.first + 16 * ci;
        int64_t rem = sg.first + sg.second - p0;
        int n_in = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        uint32_t brk32, nul32;
        uint64_t X;
        window(P, p0, X, brk32, nul32);
        // what scan_gather_kernel<0, true> decodes: the k-mer ENDING at every position, the first k-mer of a
        // run (initial score), and the runs that are not looked at (terminator within two bytes, :340-341)
        uint32_t code[16];
        for (int j = 0; j < 16; ++j) code[j] = (uint32_t)(X >> (30 - 2 * j)) & kmask;
        const uint32_t runk = run_ending(~brk32, k);
        const uint32_t first32 = runk & (brk32 << k);
        const uint32_t inside = n_in >= 16 ? 0xffffu : ((1u << n_in) - 1u);
        uint32_t dead = 0;
        uint32_t F = (first32 >> 15) & 0x1ffffu;
        while (F) {
          int b = __builtin_ctz(F);
          F &= F - 1;
          int64_t f = p0 - 1 + b;
          uint8_t z1 = buf[f + 1], z2 = buf[f + 2];
          if (b >= 1 && (z1 == 0 || z2 == 0)) dead |= 1u << (b - 1);
          if (b <= 15 && z2 == 0) dead |= 1u << b;
        }
        const uint32_t tr_first = (first32 >> 16) & inside & ~dead;
        const uint32_t real = ((run_ending(~brk32, k + 1) >> 16) & inside & ~dead) | tr_first;
        int64_t s[16];
        uint32_t live = 0;
        for (int j = 0; j < 16; ++j) {
          s[j] = 0;
          if (real & (1u << j)) {
            int64_t v = wfx[code[j] + ((tr_first >> j) & 1u ? nk : 0)];
            if (v != WFX_KILL) { s[j] = v; live |= 1u << j; }
          }
        }
        Xf f = chunk_transform(s, live);
        Ex ex;
        fx_t preM;
        int64_t prePk;
        int fz;
        chunk_walk_tr(s, live, tr_first, real, S, p0, prm, emit, ex, preM, prePk, fz);
        chunk_finish_entering_tr(S, E, preM, prePk, fz, p0, real, prm, emit);
        S = xf_apply(f, S);
        E = ex_combine(E, ex);
      }
    }
    segs.swap(next);
    ++level;
    if (level > 100000) break;
  }
  std::sort(all.begin(), all.end(), [](const Rec &a, const Rec &b) { return a.beg < b.beg; });
  int64_t n = (int64_t)all.size();
  *n_out = n;
  *beg_out = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
  *pk_out = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
  *score_out = (double *)malloc(sizeof(double) * (n + 1));
  for (int64_t i = 0; i < n; ++i) {
    (*beg_out)[i] = all[i].beg;
    (*pk_out)[i] = all[i].pk;
    (*score_out)[i] = fx_to_double(all[i].M, qs);
  }
  if (levels_out) *levels_out = level;
  return 0;
}

void emu_free(void *p) { free(p); }

double emu_fx_roundtrip(double d, int qs) { return fx_to_double((fx_t)wfx_from_double(d, qs), qs); }
}

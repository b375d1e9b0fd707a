// ks_rankseg.h -- bit-exact closed form of the reference's sequential rank accumulation.
//
// rank_kmers_w (/root/reference/src/kmer_spans.c:198-200) walks the k-mers in stable
// (count, index) order and accumulates   r <- fl(r + fl(count / T))   one k-mer at a time, in
// double.  Inside a tie group the addend d is constant, and while r stays inside one binade
// [2^e, 2^(e+1)) every step adds the same multiple of ulp(r): fl(r + d) = r + rn_u(d).  So the
// 4^k-step sequential sum collapses into a few linear pieces per group ("segments"):
//       rank(j) = x0 + (j - j0) * inc          (exact in double, no rounding)
// Host-side control-plane work: O(#distinct counts + #binades), fed by the device run-length
// table of the sorted counts; the per-k-mer evaluation runs on the GPU (rank_eval_kernel).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <vector>

namespace ks {

struct RankSeg {
  uint64_t j0;  // first ordinal (within the group) covered by this piece
  double x0;    // rank at ordinal j0
  double inc;   // exact per-step increment
};

static inline double pow2i(int e) {
  uint64_t b = (uint64_t)(1023 + e) << 52;
  double p;
  memcpy(&p, &b, 8);
  return p;
}

// Appends the pieces of one tie group (h members, addend d) starting from rank x; returns the
// rank after the last member's addend has been applied (= first rank of the next group).
static inline double rank_group_segments(double x, double d, uint64_t h, std::vector<RankSeg> &segs) {
  uint64_t j = 0;
  while (j < h) {
    if (d == 0.0) { segs.push_back({j, x, 0.0}); return x; }
    bool bulk = false;
    if (x > 0.0 && isfinite(x) && isfinite(d) && d > 0.0) {
      int ex = 0;
      (void)frexp(x, &ex);
      ex -= 1;  // x in [2^ex, 2^(ex+1))
      if (ex - 52 > -1000 && ex < 1000) {
        double u = pow2i(ex - 52), top = pow2i(ex + 1);
        double dq = d / u;  // exact (power-of-two scaling), may be huge or fractional
        if (dq < 9007199254740992.0) {
          double q = floor(dq), frac = dq - q;
          double incq;
          bool ok = true;
          if (frac < 0.5) incq = q;
          else if (frac > 0.5) incq = q + 1;
          else {  // exactly half way: ties-to-even depends on the parity of x / u
            double mx = x / u;  // 53-bit integer, exact
            if (fmod(mx, 2.0) != 0.0) ok = false;  // odd: take one real step first
            incq = (fmod(q, 2.0) == 0.0) ? q : q + 1;
          }
          if (ok) {
            double inc = incq * u;  // exact
            if (inc == 0.0) { segs.push_back({j, x, 0.0}); return x; }
            double room = (top - x - d) / inc;  // conservative: every exact sum stays below top
            double mm = floor(room) - 1.0;
            if (mm >= 1.0) {
              uint64_t m = mm > 1.8e19 ? (uint64_t)-1 : (uint64_t)mm;
              if (m > h - j) m = h - j;
              segs.push_back({j, x, inc});
              x = x + (double)m * inc;  // exact
              j += m;
              bulk = true;
            }
          }
        }
      }
    }
    if (!bulk) {
      segs.push_back({j, x, 0.0});
      x = x + d;  // one real step
      j += 1;
    }
  }
  return x;
}

// Groups: distinct counts ascending (gcount) with their first sorted position (gstart, ngroups+1
// entries, gstart[ngroups] = 4^k).  Fills seg_first (ngroups+1) and segs.
static inline void build_rank_segments(const uint32_t *gcount, const uint64_t *gstart, size_t ngroups,
                                       double total, std::vector<uint32_t> &seg_first,
                                       std::vector<RankSeg> &segs) {
  seg_first.assign(ngroups + 1, 0);
  segs.clear();
  double x = 0.0;
  for (size_t g = 0; g < ngroups; ++g) {
    seg_first[g] = (uint32_t)segs.size();
    double d = (double)(int32_t)gcount[g] / total;  // :200, int -> double then divide
    x = rank_group_segments(x, d, gstart[g + 1] - gstart[g], segs);
  }
  seg_first[ngroups] = (uint32_t)segs.size();
}

}  // namespace ks

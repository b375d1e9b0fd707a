// ks_chunk.cuh -- per-chunk logic of the kmer_spans hot path, shared by the sm_100a kernels
// (ks_kernels.cuh) and by the host-side emulation used ONLY by tests (tests/emu/ks_emu.cpp).
//
// A chunk is 16 consecutive byte positions of the concatenated sequence buffer, owned by one
// thread.  Everything here is straight-line code over register arrays (fully unrolled).
//
// Reference behaviour restated here (file:line in /root/reference/src/kmer_spans.c):
//   * 2-bit code (c>>1)&3, only N/n (and the terminator) break a run            :34-35,111-132
//   * counting rule incl. the exact-k-at-terminator drop                          :135-155
//   * scored indices of a run [a,b): i = a+k .. b-1, w_i = W[code(i-1)] - thr     :261-296
//   * S' = (S + w > 0) ? S + w : 0, start / leftmost peak / close bookkeeping     :269-291
//
// Arithmetic: the scan runs in EXACT fixed point.  Every per-k-mer score (W[code]-thr, the
// double the reference computes at :268) is converted once to a signed multiple of
// q = 2^-QS (QS chosen per table so that max|w| < 2^(57-QS), i.e. 57-bit table entries: exact for
// every weight whose last mantissa bit is >= 2^-QS, truncated toward zero below); running sums
// are 128-bit integers in units of q, partial sums inside one chunk are 64-bit.  Integer addition is associative, so tile, block, warp and
// multi-GPU decompositions all give the same bits, and the max-plus transforms
// x -> max(x + a, b) compose exactly.  DESIGN.md discusses the parity consequences.
#pragma once
#include <stdint.h>
#include <limits.h>
#include <string.h>

#ifdef __CUDACC__
#define KS_HD __host__ __device__ __forceinline__
#else
#define KS_HD inline
#endif

namespace ks {

constexpr int CHUNK = 16;
typedef __int128 fx_t;                 // running sums, units of q = 2^-QS
constexpr int64_t WFX_KILL = INT64_MIN;  // table sentinel: NaN / -inf weight => state forced to 0


// ---------------------------------------------------------------------------------------------
// K1: 2-bit packing with N masking.  16 ASCII bytes (4 little-endian words) ->
//   pk  : 16 x 2 bits, FIRST base most significant (so that (pk_prev << 32 | pk_cur) >> s yields
//         rolling k-mer codes directly)
//   brk : bit j set iff byte j breaks a run (N, n or the terminator)        (:35,112,123)
//   nul : bit j set iff byte j is the terminator (needed only by the counting rule :143-144)
KS_HD uint32_t zero_bytes(uint32_t v) {  // 0x80 in every byte of v that is zero (exact)
  uint32_t t = (v & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return ~(t | v | 0x7f7f7f7fu);
}
KS_HD uint32_t nib_of_flags(uint32_t m80) {  // 0x80-per-byte flags -> 4 bits, byte 0 -> bit 0
  return (((m80 >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xfu;
}
KS_HD void pack16(const uint32_t w[4], uint32_t &pk, uint32_t &brk, uint32_t &nul) {
  pk = 0; brk = 0; nul = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t t = (w[i] >> 1) & 0x03030303u;             // (c >> 1) & 3 per byte
    uint32_t b = (t * 0x40100401u) >> 24;               // 4 bases, byte 0 most significant
    pk |= b << (24 - 8 * i);
    uint32_t z = zero_bytes(w[i]);
    uint32_t n = zero_bytes((w[i] | 0x20202020u) ^ 0x6e6e6e6eu);
    nul |= nib_of_flags(z) << (4 * i);
    brk |= nib_of_flags(z | n) << (4 * i);
  }
}

// run masks: bit b of the result is set iff bits b-len+1 .. b of ok are all set (len in 1..16)
KS_HD uint32_t run_ending(uint32_t ok, int len) {
  uint32_t r = ok;
  int have = 1;
  while (have * 2 <= len) { r &= r << have; have *= 2; }
  if (have < len) r &= r << (len - have);
  return r;
}

// decode for the SCAN from packed input.  X = (pk of the 16 positions before the chunk) << 32 | pk of
// the chunk; brk32 = brk of the 16 positions before (bits 0..15) | brk of the chunk << 16.
// code[j] = k-mer ending at position j-1; bit j of `scored` set iff position j and the k positions
// before it are all non-break and j < n_in.
KS_HD void decode_scan(uint64_t X, uint32_t brk32, int k, uint32_t kmask, int n_in, uint32_t code[CHUNK],
                       uint32_t &scored) {
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) code[j] = (uint32_t)(X >> (32 - 2 * j)) & kmask;
  uint32_t r = run_ending(~brk32, k + 1) >> 16;
  uint32_t inside = n_in >= 16 ? 0xffffu : ((1u << n_in) - 1u);
  scored = r & inside;
}

// decode for COUNTING (from the ASCII chunk, packed on the fly): code[j] = k-mer ENDING at position
// j; bit j of `counted` set iff sequence_kmer_count (:135-155) counts it: all k positions
// non-break, and not "first k-mer of its run with the terminator right behind it" (:143-144).
// nul32 like brk32; next_nul = position 16 (first of the next chunk) is the terminator.
KS_HD void decode_count(uint64_t X, uint32_t brk32, uint32_t nul32, uint32_t next_nul, int k, uint32_t kmask,
                        uint32_t code[CHUNK], uint32_t &counted) {
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) code[j] = (uint32_t)(X >> (30 - 2 * j)) & kmask;
  uint32_t runk = run_ending(~brk32, k);
  uint32_t first = runk & (brk32 << k);                       // the run starts exactly k positions back
  uint32_t nulnext = (nul32 >> 1) | (next_nul ? 0x80000000u : 0u);  // bit b: position b+1 is the terminator
  counted = (runk & ~(first & nulnext)) >> 16;
}

// ---------------------------------------------------------------------------------------------
// double -> table entry (signed multiple of 2^-qs).  Exact when the double's last mantissa bit
// is >= 2^-qs, truncated toward zero otherwise.  NaN and anything <= -2^62 map to WFX_KILL.
KS_HD int64_t wfx_from_double(double d, int qs) {
  if (!(d == d) || d <= -0x1p40) return WFX_KILL;  // NaN, -inf, or "as good as -inf"
  uint64_t bits;
  memcpy(&bits, &d, 8);
  int e = (int)((bits >> 52) & 0x7ff);
  uint64_t m = bits & ((1ull << 52) - 1);
  bool neg = (bits >> 63) != 0;
  if (e == 0) e = 1; else m |= 1ull << 52;  // subnormal: m * 2^-1074
  int sh = e - 1075 + qs;  // value * 2^qs = m * 2^sh
  uint64_t v;
  if (sh >= 0) {
    if (sh > 4) return INT64_MAX;  // cannot happen when qs = qs_for_max(max |w|)
    v = m << sh;
  } else if (sh > -64) {
    v = m >> (-sh);
  } else {
    v = 0;
  }
  return neg ? -(int64_t)v : (int64_t)v;
}

// number of fraction bits for a table whose largest finite |w| is wmax: max|w| < 2^(57-qs).
// 57-bit scores keep every partial sum of one 16-position chunk below 2^61, so the per-position
// arithmetic inside a chunk is plain 64-bit; only the carries between chunks are 128-bit.
KS_HD int qs_for_max(double wmax) {
  uint64_t bits;
  memcpy(&bits, &wmax, 8);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;  // wmax in [2^e, 2^(e+1))
  if (wmax == 0.0) e = -1;
  int E = e + 1;  // wmax < 2^E
  int qs = 57 - E;
  if (qs > 1100) qs = 1100;  // covers every finite double (subnormals included)
  return qs;
}

// 128-bit running sum (units of 2^-qs) -> nearest double (ties to even).
KS_HD double fx_to_double(fx_t v, int qs) {
  bool neg = v < 0;
  unsigned __int128 u = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
  if (u == 0) return 0.0;
  uint64_t hi = (uint64_t)(u >> 64), lo = (uint64_t)u;
  int msb;
  if (hi) { msb = 64; uint64_t t = hi; while (t >>= 1) ++msb; }
  else { msb = 0; uint64_t t = lo; while (t >>= 1) ++msb; }
  uint64_t m;
  int sh = 0;
  if (msb <= 52) {
    m = lo;
  } else {
    sh = msb - 52;
    unsigned __int128 q = u >> sh;
    unsigned __int128 rem = u & ((((unsigned __int128)1) << sh) - 1);
    unsigned __int128 half = ((unsigned __int128)1) << (sh - 1);
    m = (uint64_t)q;
    if (rem > half || (rem == half && (m & 1))) ++m;
  }
  // m * 2^(sh - qs); m <= 2^53 so (double)m is exact, scaling by a power of two is exact
  double r = (double)m;
  int ex = sh - qs;
  // scale in two safe steps (|ex| <= 190)
  uint64_t pb;
  double p;
  int e1 = ex / 2, e2 = ex - e1;
  pb = (uint64_t)(1023 + e1) << 52; memcpy(&p, &pb, 8); r *= p;
  pb = (uint64_t)(1023 + e2) << 52; memcpy(&p, &pb, 8); r *= p;
  return neg ? -r : r;
}

// min_score (double) -> smallest integer number of units u with u * 2^-qs >= min_score.
// Saturates far outside the reachable range of sums.
KS_HD fx_t fx_ceil_units(double x, int qs) {
  const fx_t BIG = ((fx_t)1) << 126;
  if (!(x == x)) return BIG;  // NaN: nothing qualifies (M >= NaN is false)
  uint64_t bits;
  memcpy(&bits, &x, 8);
  int e = (int)((bits >> 52) & 0x7ff);
  uint64_t m = bits & ((1ull << 52) - 1);
  bool neg = (bits >> 63) != 0;
  if (e == 0 && m == 0) return 0;
  if (e - 1023 + qs >= 100) return neg ? -BIG : BIG;        // beyond any reachable sum (< 2^102 units)
  if (e == 0) e = 1; else m |= 1ull << 52;                  // subnormal: m * 2^-1074
  int sh = e - 1075 + qs;  // <= 47
  unsigned __int128 v;
  bool inexact = false;
  if (sh >= 0) {
    v = ((unsigned __int128)m) << sh;
  } else if (sh > -64) {
    v = m >> (-sh);
    inexact = (m & ((1ull << (-sh)) - 1)) != 0;
  } else {
    v = 0;
    inexact = true;
  }
  fx_t r = (fx_t)v;
  if (neg) return -r;          // ceil of a negative value: truncation toward zero
  return inexact ? r + 1 : r;  // ceil of a positive value
}

// ---------------------------------------------------------------------------------------------
// max-plus transform x -> kill ? b : max(x + a, b).  (a = -inf is carried as the flag `kill`.)
struct Xf {
  fx_t a, b;
  uint32_t kill;
};
constexpr int XF_IDENT_SHIFT = 124;
KS_HD Xf xf_identity() { Xf f; f.a = 0; f.b = -(((fx_t)1) << XF_IDENT_SHIFT); f.kill = 0; return f; }
KS_HD fx_t fx_max(fx_t x, fx_t y) { return x > y ? x : y; }
// f first, then g
KS_HD Xf xf_compose(const Xf &f, const Xf &g) {
  if (g.kill) return g;
  Xf r;
  r.kill = f.kill;
  r.a = f.a + g.a;
  r.b = fx_max(f.b + g.a, g.b);
  return r;
}
KS_HD fx_t xf_apply(const Xf &f, fx_t x) { return f.kill ? f.b : fx_max(x + f.a, f.b); }

// chunk transform.  s[j] valid where bit j of live is set; other positions force the state to 0.
// 64-bit inside the chunk (|s| < 2^57, 16 terms), widened at the end.
template <class Scores>
KS_HD Xf chunk_transform(const Scores &s, uint32_t live) {
  int64_t a = 0, b = -(1ll << 62);
  uint32_t kill = 0;
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) {
    if (live & (1u << j)) {
      a += s[j];
      int64_t t = b + s[j];
      b = t > 0 ? t : 0;
    } else {
      kill = 1; a = 0; b = 0;
    }
  }
  Xf f;
  f.a = (fx_t)a;
  f.b = (fx_t)b;
  f.kill = kill;
  return f;
}

// ---------------------------------------------------------------------------------------------
// Open-excursion state carried across chunk boundaries (segmented max-scan element).
//   reset = 1: the chunk determines the state at its end by itself:
//              open = 1 -> (beg, M, pk) of the excursion open at the chunk end; open = 0 -> closed
//   reset = 0: the whole chunk lies inside one excursion entering from the left (no zero, no
//              start); (M, pk) = maximum over the chunk, beg inherited.
struct Ex {
  fx_t M;
  int64_t beg, pk;
  uint32_t reset, open;
};
KS_HD Ex ex_identity() { Ex e; e.M = -(((fx_t)1) << 126); e.beg = -1; e.pk = -1; e.reset = 0; e.open = 1; return e; }
// l first, then r.  Strict > keeps the LEFTMOST maximum (reference :287).
KS_HD Ex ex_combine(const Ex &l, const Ex &r) {
  if (r.reset) return r;
  Ex o = l;
  if (r.M > l.M) { o.M = r.M; o.pk = r.pk; }
  return o;
}

struct ScanParams {
  uint64_t min_width;  // compared as in the reference (:279): (uint64)(pk - beg) >= min_width
  fx_t min_units;      // M >= min_score, in units of 2^-qs
};

KS_HD bool qualifies(const ScanParams &p, int64_t beg, int64_t pk, fx_t M) {
  return (uint64_t)(pk - beg) >= p.min_width && M >= p.min_units;
}

// Walk a chunk with its true incoming state S_in (> 0 means an excursion enters from the left).
// Emits every excursion that both starts and closes inside the chunk; returns in `pre` the
// (M, pk) over the positions before the first zero (the part belonging to the entering
// excursion), `first_zero` = position index (0..15) of that zero or -1, and in `ex` the
// segmented-scan element of this chunk.  p0 = global position of chunk byte 0.
//
// 64-bit inside the chunk: the walk runs on sigma = min(S_in, 2^62).  A chunk moves the state by
// less than 2^61, so with S_in >= 2^62 no position reaches 0 either way and all comparisons agree;
// maxima are shifted back by S_in - sigma when they leave the chunk.
template <class Scores, class Emit>
KS_HD void chunk_walk(const Scores &s, uint32_t live, fx_t S_in, int64_t p0,
                      const ScanParams &prm, Emit &emit, Ex &ex, fx_t &preM, int64_t &prePk,
                      int &first_zero) {
  const fx_t cap = ((fx_t)1) << 62;
  const int64_t sigma = S_in > cap ? (1ll << 62) : (int64_t)S_in;
  const fx_t shift = S_in - (fx_t)sigma;
  const int64_t NEG = -(1ll << 62);
  int64_t S = sigma;
  int64_t M = NEG;
  int64_t pk = -1, beg = -1;
  int64_t pM = NEG, pPk = -1;
  bool started = false;  // a start happened inside this chunk
  first_zero = -1;
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) {
    int64_t Sn = 0;
    if (live & (1u << j)) {
      Sn = S + s[j];
      Sn = Sn > 0 ? Sn : 0;
    }
    if (S == 0 && Sn > 0) { started = true; beg = p0 + j; pk = p0 + j; M = Sn; }
    if (Sn == 0) {
      if (S > 0 && started) {  // close at p0 + j of an excursion that started in this chunk
        if (qualifies(prm, beg, pk, (fx_t)M)) emit(beg, pk, (int64_t)(p0 + j), (fx_t)M);
      }
      if (first_zero < 0) { first_zero = j; pM = M; pPk = pk; }
    } else if (Sn > M) {
      M = Sn; pk = p0 + j;
    }
    S = Sn;
  }
  if (first_zero < 0) { pM = M; pPk = pk; }
  preM = pM == NEG ? -(((fx_t)1) << 126) : (fx_t)pM + shift;
  prePk = pPk;
  // segmented-scan element
  if (S > 0) {
    ex.open = 1;
    ex.pk = pk;
    if (started) { ex.reset = 1; ex.beg = beg; ex.M = (fx_t)M; }
    else { ex.reset = 0; ex.beg = -1; ex.M = (fx_t)M + shift; }
  } else {
    ex.reset = 1; ex.open = 0; ex.M = -(((fx_t)1) << 126); ex.beg = -1; ex.pk = -1;
  }
}

// After the exclusive segmented scan: finish the excursion that ENTERED this chunk (S_in > 0)
// and closes at its first zero.
template <class Emit>
KS_HD void chunk_finish_entering(fx_t S_in, const Ex &ex_in, fx_t preM, int64_t prePk, int first_zero,
                                 int64_t p0, const ScanParams &prm, Emit &emit) {
  if (S_in > 0 && first_zero >= 0) {
    fx_t M = ex_in.M;
    int64_t pk = ex_in.pk;
    if (preM > M) { M = preM; pk = prePk; }
    if (qualifies(prm, ex_in.beg, pk, M)) emit(ex_in.beg, pk, (int64_t)(p0 + first_zero), M);
  }
}

// ---------------------------------------------------------------------------------------------
// Chunk summaries for the fast walk (min_width >= 15: an excursion inside one chunk cannot qualify).
// With the prefix sums P_j of a chunk and its zero-start trajectory Bz, the trajectory from the entering
// state is S_j = max(S_in + P_j, Bz_j): it reaches 0 inside the chunk iff a position is forced to 0 or
// S_in + min P <= 0, and from there on it IS Bz.  So per chunk it is enough to keep
//   mn, mx + leftmost position am   min / max of P
//   bm, bbeg, bpk, open             peak, start and leftmost peak position of the excursion Bz has open at
//                                   the chunk end
// bits = am << 19 | bbeg << 23 | bpk << 27 | open << 31 (stored next to the live mask in st_flags).
struct ChunkSummary {
  int64_t mn, mx, bm;
  uint32_t bits;
};

// any chunk: the plain recurrences, position by position (v = WFX_KILL forces the state to 0)
struct GeneralChunk {
  int64_t ta, tb, mn, mx, Bz, bM;
  uint32_t kill, live, am, bbeg, bpk;
  KS_HD void init() {
    ta = 0; tb = -(1ll << 62); kill = 0; live = 0;
    mn = 1ll << 62; mx = -(1ll << 62); Bz = 0; bM = 0; am = 0; bbeg = 0; bpk = 0;
  }
  template <bool kSumm>
  KS_HD void step(int j, int64_t v) {
    const bool ok = v != WFX_KILL;
    if (ok) {
      live |= 1u << j;
      ta += v;
      int64_t t = tb + v;
      tb = t > 0 ? t : 0;
    } else {
      kill = 1; ta = 0; tb = 0;
    }
    if (kSumm) {
      mn = ta < mn ? ta : mn;
      if (ta > mx) { mx = ta; am = (uint32_t)j; }
      int64_t t = Bz + v;
      const int64_t Bn = (ok && t > 0) ? t : 0;
      if (Bz == 0 && Bn > 0) { bbeg = (uint32_t)j; bpk = (uint32_t)j; bM = Bn; }
      else if (Bn > bM) { bM = Bn; bpk = (uint32_t)j; }
      Bz = Bn;
    }
  }
  KS_HD ChunkSummary summary() const {
    ChunkSummary r;
    r.mn = mn; r.mx = mx; r.bm = bM;
    r.bits = (am << 19) | (bbeg << 23) | (bpk << 27) | (Bz > 0 ? 0x80000000u : 0u);
    return r;
  }
};

// a chunk whose 16 positions are all scored and none is forced to 0 (almost every chunk of a genome):
// everything follows from the prefix sums alone, one 64-bit add and four compares per position:
//   mnT = min_j P_j            -> transform b = P_15 - mnT, zero test of the fast walk
//   Bz_j = P_j - min(0, mnT_j); Bz_j == 0 iff P_j <= mnT_{j-1} and P_j <= 0
//   bMx = max of P since the last zero of Bz (position bpk), jz1 = last zero + 1
//   mx = max_j P_j with its leftmost position am
// bad: a WFX_KILL entry was met (the only table entry whose high word is INT32_MIN) -> use GeneralChunk
struct FastChunk {
  int64_t P, mnT, bMx, mx;
  uint32_t am, bpk, jz1;
  int32_t hmin;  // smallest high word met: INT32_MIN iff a WFX_KILL entry was among the values
  KS_HD void init() { P = 0; mnT = 1ll << 62; bMx = 0; mx = -(1ll << 62); am = 0; bpk = 0; jz1 = 0; hmin = 0; }
  KS_HD bool bad() const { return hmin == INT32_MIN; }
  KS_HD void step(int j, int64_t v) {
    const int32_t vh = (int32_t)(v >> 32);
    hmin = vh < hmin ? vh : hmin;
    P += v;
    const bool newmin = P <= mnT;
    mnT = newmin ? P : mnT;
    const bool isz = newmin && P <= 0;
    const bool up = isz || P > bMx;
    bMx = up ? P : bMx;
    bpk = up ? (uint32_t)j : bpk;
    jz1 = isz ? (uint32_t)(j + 1) : jz1;
    if (P > mx) { mx = P; am = (uint32_t)j; }
  }
  KS_HD int64_t a() const { return P; }
  KS_HD int64_t b() const { return P - mnT; }
  KS_HD ChunkSummary summary() const {
    ChunkSummary r;
    const int64_t m0 = mnT < 0 ? mnT : 0;
    r.mn = mnT; r.mx = mx; r.bm = bMx - m0;
    const bool open = jz1 != (uint32_t)CHUNK;
    r.bits = (am << 19) | ((jz1 & 15u) << 23) | (bpk << 27) | (open ? 0x80000000u : 0u);
    return r;
  }
};

// open-excursion element of a chunk (or a unit of two chunks) from its summary and the state entering it
// (S_in >= 0); `closing`: an excursion enters and returns to 0 inside.  am / bbeg / bpk are offsets from p0.
KS_HD void fast_walk_element_at(fx_t S_in, bool head, bool all_live, int64_t mn, int64_t mx, int64_t bm, uint32_t am,
                                uint32_t bbeg, uint32_t bpk, bool open, int64_t p0, Ex &ex, bool &closing) {
  const bool zero = head || S_in <= 0 || !all_live || S_in + (fx_t)mn <= 0;
  if (!zero) {
    ex.reset = 0; ex.open = 1; ex.beg = -1;
    ex.M = S_in + (fx_t)mx;
    ex.pk = p0 + am;
  } else if (open) {
    ex.reset = 1; ex.open = 1;
    ex.beg = p0 + bbeg;
    ex.pk = p0 + bpk;
    ex.M = (fx_t)bm;
  } else {
    ex.reset = 1; ex.open = 0; ex.M = -(((fx_t)1) << 126); ex.beg = -1; ex.pk = -1;
  }
  closing = !head && S_in > 0 && zero;
}
KS_HD void fast_walk_element(fx_t S_in, bool head, uint32_t live, const ChunkSummary &sm, int64_t p0, Ex &ex,
                             bool &closing) {
  fast_walk_element_at(S_in, head, live == 0xffffu, sm.mn, sm.mx, sm.bm, (sm.bits >> 19) & 15u, (sm.bits >> 23) & 15u,
                       (sm.bits >> 27) & 15u, (sm.bits & 0x80000000u) != 0, p0, ex, closing);
}
// the entering excursion of a closing chunk cannot qualify: its peak lies at or before the last position of the
// chunk (p0 + 15; p0 + 31 for a unit of two chunks) and is at most max(M so far, S_in + max P)
KS_HD bool fast_walk_cannot_qualify(const Ex &e_in, fx_t S_in, int64_t p0, int64_t mx, const ScanParams &prm,
                                    int span = CHUNK) {
  if ((uint64_t)(p0 + span - 1 - e_in.beg) < prm.min_width) return true;
  return fx_max(e_in.M, S_in + (fx_t)mx) < prm.min_units;
}

// ---------------------------------------------------------------------------------------------
// Units of two chunks (min_width >= 31: an excursion inside 32 positions cannot qualify either).  The gather kernel
// then leaves one record per 32 positions -- half the block scan, half the records, half the walk.  A unit is the
// merge of the summaries of its two chunks; offsets are relative to the first position of the unit.
struct UnitSummary {
  int64_t ta, tb;  // transform x -> kill ? tb : max(x + ta, tb), 64-bit (32 terms below 2^57)
  int64_t mn, mx, bm;
  uint32_t kill, all_live;  // all_live: every position scored, none forced to 0
  uint32_t am, bbeg, bpk, open;
};
KS_HD UnitSummary unit_from_chunk(int64_t ta, int64_t tb, uint32_t kill, uint32_t live, const ChunkSummary &sm) {
  UnitSummary u;
  u.ta = ta; u.tb = tb; u.kill = kill; u.all_live = live == 0xffffu ? 1u : 0u;
  u.mn = sm.mn; u.mx = sm.mx; u.bm = sm.bm;
  u.am = (sm.bits >> 19) & 15u; u.bbeg = (sm.bits >> 23) & 15u; u.bpk = (sm.bits >> 27) & 15u; u.open = sm.bits >> 31;
  return u;
}
// l = positions 0..15, r = positions 16..31
KS_HD UnitSummary unit_merge(const UnitSummary &l, const UnitSummary &r) {
  UnitSummary u;
  if (r.kill) {
    u.ta = r.ta; u.tb = r.tb; u.kill = 1;
  } else {
    const int64_t t = l.tb + r.ta;
    u.ta = l.ta + r.ta; u.tb = t > r.tb ? t : r.tb; u.kill = l.kill;
  }
  u.all_live = l.all_live & r.all_live;
  // extrema of the prefix sums (exact when all positions are live; an upper bound for mx otherwise, which is all
  // the walk asks of it then)
  const int64_t cmn = l.ta + r.mn, cmx = l.ta + r.mx;
  u.mn = cmn < l.mn ? cmn : l.mn;
  if (cmx > l.mx) { u.mx = cmx; u.am = CHUNK + r.am; } else { u.mx = l.mx; u.am = l.am; }
  // zero-start trajectory: s1 enters the second chunk; if it reaches 0 there, the second chunk's own trajectory
  // takes over, otherwise the excursion open at the end of the first chunk runs on
  const int64_t s1 = l.kill ? l.tb : (l.ta > l.tb ? l.ta : l.tb);
  if (s1 <= 0 || !r.all_live || s1 + r.mn <= 0) {
    u.open = r.open; u.bm = r.bm; u.bbeg = CHUNK + r.bbeg; u.bpk = CHUNK + r.bpk;
  } else {
    const int64_t cand = s1 + r.mx;
    u.open = 1; u.bbeg = l.bbeg;
    if (cand > l.bm) { u.bm = cand; u.bpk = CHUNK + r.am; } else { u.bm = l.bm; u.bpk = l.bpk; }
  }
  return u;
}
// the unit opens a scan: whatever entered is replaced by state 0
KS_HD void unit_make_head(UnitSummary &u) {
  const int64_t v = u.kill ? u.tb : (u.ta > u.tb ? u.ta : u.tb);
  u.kill = 1; u.ta = 0; u.tb = v;
}
KS_HD void fast_walk_unit(fx_t S_in, bool head, const UnitSummary &u, int64_t p0, Ex &ex, bool &closing) {
  fast_walk_element_at(S_in, head, u.all_live != 0, u.mn, u.mx, u.bm, u.am, u.bbeg, u.bpk, u.open != 0, p0, ex, closing);
}

// ---------------------------------------------------------------------------------------------
// Transition-score scan (find_kmer_tr_lr_regions, /root/reference/src/kmer_spans.c:329-395).
// Same clamped scan, different bookkeeping:
//   * `first` marks positions that carry a run's INITIAL score (the first k-mer of the run, :344-354);
//     the reference books everything that happens there one position later (i = a + k)
//   * `real` marks positions the reference's loop visits; a close at a real position is tested on
//     width only (:377) and ALWAYS makes the scan resume behind the peak (:382-388) -> emit.child;
//     a close forced by the end of the run is the terminal test (:392-393): reported, not re-scanned
// emit.out(beg, pk, M) reports a region, emit.child(pk, c) asks for the re-scan of (pk, c].
template <class Scores, class Emit>
KS_HD void chunk_walk_tr(const Scores &s, uint32_t live, uint32_t first, uint32_t real, fx_t S_in, int64_t p0,
                         const ScanParams &prm, Emit &emit, Ex &ex, fx_t &preM, int64_t &prePk,
                         int &first_zero) {
  const fx_t cap = ((fx_t)1) << 62;
  const int64_t sigma = S_in > cap ? (1ll << 62) : (int64_t)S_in;
  const fx_t shift = S_in - (fx_t)sigma;
  const int64_t NEG = -(1ll << 62);
  int64_t S = sigma;
  int64_t M = NEG;
  int64_t pk = -1, beg = -1;
  int64_t pM = NEG, pPk = -1;
  bool started = false;
  first_zero = -1;
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) {
    int64_t Sn = 0;
    if (live & (1u << j)) {
      Sn = S + s[j];
      Sn = Sn > 0 ? Sn : 0;
    }
    const int64_t pos = p0 + j + ((first >> j) & 1u);
    if (S == 0 && Sn > 0) { started = true; beg = pos; pk = pos; M = Sn; }
    if (Sn == 0) {
      if (S > 0 && started) {
        if (qualifies(prm, beg, pk, (fx_t)M)) emit.out(beg, pk, (fx_t)M);
        if (real & (1u << j)) emit.child(pk, (int64_t)(p0 + j), prm.min_width);
      }
      if (first_zero < 0) { first_zero = j; pM = M; pPk = pk; }
    } else if (Sn > M) {
      M = Sn; pk = pos;
    }
    S = Sn;
  }
  if (first_zero < 0) { pM = M; pPk = pk; }
  preM = pM == NEG ? -(((fx_t)1) << 126) : (fx_t)pM + shift;
  prePk = pPk;
  if (S > 0) {
    ex.open = 1;
    ex.pk = pk;
    if (started) { ex.reset = 1; ex.beg = beg; ex.M = (fx_t)M; }
    else { ex.reset = 0; ex.beg = -1; ex.M = (fx_t)M + shift; }
  } else {
    ex.reset = 1; ex.open = 0; ex.M = -(((fx_t)1) << 126); ex.beg = -1; ex.pk = -1;
  }
}

template <class Emit>
KS_HD void chunk_finish_entering_tr(fx_t S_in, const Ex &ex_in, fx_t preM, int64_t prePk, int first_zero,
                                    int64_t p0, uint32_t real, const ScanParams &prm, Emit &emit) {
  if (S_in > 0 && first_zero >= 0) {
    fx_t M = ex_in.M;
    int64_t pk = ex_in.pk;
    if (preM > M) { M = preM; pk = prePk; }
    if (qualifies(prm, ex_in.beg, pk, M)) emit.out(ex_in.beg, pk, M);
    if (real & (1u << first_zero)) emit.child(pk, (int64_t)(p0 + first_zero), prm.min_width);
  }
}
// re-scan of (pk, c]: a region inside it has pk' - beg' <= c - pk - 2, so shorter tails are dropped
KS_HD bool tr_child_wanted(int64_t pk, int64_t c, uint64_t min_width) {
  int64_t room = c - pk - 2;
  return room >= 0 && (uint64_t)room >= min_width;
}

// Next-level segment spawned by a qualifying excursion (beg, pk, c): the reference restarts the
// scan with S = 0 at pk + 1 (:281-282,303) and by the re-synchronisation lemma is back
// on the parent trajectory at c, so the child scan covers [pk + 1, c].  Without in-scan counting a
// child that is too short to hold a qualifying excursion is dropped.
KS_HD bool child_segment(int64_t pk, int64_t c, uint64_t min_width, bool inscan, int64_t &start,
                         int64_t &len) {
  start = pk + 1;
  len = c - pk;
  if (inscan) return true;
  int64_t room = c - pk - 2;  // largest possible pk' - beg' inside (pk, c)
  return room >= 0 && (uint64_t)room >= min_width;
}
// chunks a segment occupies: one spare position after its end so that an excursion still open at
// the last position is closed inside the segment's own chunks
KS_HD int64_t segment_chunks(int64_t len) { return len / CHUNK + 1; }

}  // namespace ks

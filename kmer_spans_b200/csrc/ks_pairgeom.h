// ks_pairgeom.h -- geometry of the bucketed count (ks_count.cuh): which bucket a PAIR of consecutive k-mers is
// filed under and what its 16-bit sub-key holds.  Host + device (tests/emu folds the same functions on the CPU).
#pragma once
#include "ks_chunk.cuh"

namespace ks {

// buckets: 1024 for k <= 12; 4096 at k = 13, where 1024 would leave 18-bit sub-keys (one CTA of 1024 threads per SM
// then holds the 4096 staging rows)
constexpr int bk_log(int k) { return k >= 13 ? 12 : 10; }
constexpr int bk_threads(int k) { return k >= 13 ? 1024 : 256; }
constexpr int BK_MAX_BUCKETS = 1 << 12;
constexpr int BK_ROUNDS = 6;    // chunks per thread and tile, loaded in two batches of three
constexpr int bk_tile_chunks(int k) { return bk_threads(k) * BK_ROUNDS; }  // 12 pairs per bucket and tile

// The three bases around a pair, read straight from the packed window: Y = a.c.b (2k + 2 bits), c the shared (k-1)-mer.
//   bucket  = leading LOG bits of c           = (Y >> (REST + 2)) & (2^LOG - 1),  REST = 2k - 2 - LOG bits of c remain
//   sub-key = a | rest of c | b  (REST + 4 bits) = (Y & LOW) | ((Y >> LOG) & (3 << (REST + 2))),  LOW = REST + 2 ones
// so that phase 2 indexes its c.b table with sub & LOW and its a.c table with sub >> 2.
template <int K>
struct PairGeom {
  static constexpr int LOG = bk_log(K), NB = 1 << LOG, THREADS = bk_threads(K), TILE_CHUNKS = bk_tile_chunks(K);
  static constexpr int REST = 2 * K - 2 - LOG;
  static constexpr uint32_t LOW = (1u << (REST + 2)) - 1u;
  static constexpr uint32_t KMASK = (uint32_t)(((uint64_t)1 << (2 * K)) - 1u);
  static constexpr uint32_t ENTRIES = 1u << (REST + 2);  // per bucket and table = 4^K / buckets
  static KS_HD uint32_t bucket4(uint32_t y) { return (y >> REST) & ((uint32_t)(NB - 1) << 2); }  // bucket * 4
  static KS_HD uint32_t sub(uint32_t y) { return (y & LOW) | ((y >> LOG) & (3u << (REST + 2))); }
  static KS_HD uint32_t code_ac(uint32_t bucket, uint32_t sub) {
    return ((sub >> (REST + 2)) << (2 * K - 2)) | (bucket << REST) | ((sub >> 2) & ((1u << REST) - 1u));
  }
  static KS_HD uint32_t code_cb(uint32_t bucket, uint32_t sub) {
    return (bucket << (REST + 2)) | (sub & LOW);
  }
  // phase 2 keeps, per bucket, tabB[sub & LOW] (the k-mers c.b: the bucket's own slice of the count table) and
  // tabA[sub >> 2] = [a][rest of c] (the k-mers a.c); the fold adds tabA of bucket(c) into counts[a.c]:
  static KS_HD size_t fold_index(uint32_t code) {  // where the count of the k-mer `code` = a.c sits in the second table
    const uint32_t a = code >> (2 * K - 2), c = code & (KMASK >> 2);
    return (size_t)(c >> REST) * ENTRIES + ((size_t)a << REST) + (c & ((1u << REST) - 1u));
  }
};

}  // namespace ks

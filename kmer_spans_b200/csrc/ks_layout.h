// ks_layout.h -- layout of the concatenated sequence buffer in HBM (host + device, plain C++).
//
//   [16 zero bytes][seq 0][0][seq 1][0] ... [seq n-1][0][zero padding to a multiple of 16][16 zero bytes]
//
// A NUL byte plays the role of the C-string terminator the reference stops at
// (/root/reference/src/kmer_spans.c:121,140,261); sequences are R strings and cannot contain one.
// Global position g of base i of sequence s is start[s] + i; start[0] = 16, so the 16 bytes in
// front of every chunk exist, and every chunk load stays inside the allocation (+KS_SLACK).
#pragma once
#include <stdint.h>
#define KS_FRONT_PAD 16
#define KS_SLACK 64

static inline int64_t ks_layout_total(const int64_t *lens, int nseq, int64_t *starts /*nseq+1 or NULL*/) {
  int64_t cur = KS_FRONT_PAD;
  for (int i = 0; i < nseq; ++i) {
    if (starts) starts[i] = cur;
    cur += lens[i] + 1;
  }
  if (starts) starts[nseq] = cur;
  int64_t tot = ((cur + 15) / 16) * 16 + 16;
  return tot;
}

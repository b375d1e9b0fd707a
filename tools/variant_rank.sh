#!/bin/bash
# usage: tools/variant_rank.sh "<nvcc extra flags>" <tag>  -- rebuild with the flags, time the rank-mode pipeline
cd "$(dirname "$0")/.."
KS_NVCC_EXTRA="$1" python kmer_spans_b200/build.py --force > gpurun_out/build_$2.log 2>&1
echo "$2: $(python tools/mode_timing.py 250000000 12 2>&1 | grep 'rank thr=.75')"

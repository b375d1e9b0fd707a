#!/bin/bash
# usage: tools/variant_modes.sh "<nvcc extra flags>" <tag>  -- rebuild with the flags, print per-mode timings
cd "$(dirname "$0")/.."
KS_NVCC_EXTRA="$1" python kmer_spans_b200/build.py --force > gpurun_out/build_$2.log 2>&1 || { tail -5 gpurun_out/build_$2.log; exit 1; }
echo "== $2"; python tools/mode_timing.py 250000000 12 | grep -E "rank thr=.75|log2" | cut -c1-220

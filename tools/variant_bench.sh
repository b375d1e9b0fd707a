#!/bin/bash
# usage: tools/variant_bench.sh "<nvcc extra flags>" <tag>   -- rebuild the library with the flags, run a short bench
set -e
cd "$(dirname "$0")/.."
KS_NVCC_EXTRA="$1" python kmer_spans_b200/build.py --force > gpurun_out/build_$2.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_$2.log 2> gpurun_out/bench_$2.err || tail -5 gpurun_out/bench_$2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$2.log"))
print("$2", round(d["value"],2), "Gb/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY

#!/bin/bash
# usage: tools/variant_ncu.sh <kernel regex> -- gpu__time_duration of the matching kernels for every tools/bin/lib_*.so
cd "$(dirname "$0")/.."
cp kmer_spans_b200/csrc/libkspans_cuda.so /tmp/lib_default.so
for f in tools/bin/lib_*.so; do
  cp "$f" kmer_spans_b200/csrc/libkspans_cuda.so
  t=$(basename $f .so)
  ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"$1" -c 4 --csv --log-file gpurun_out/vn_$t.csv python tools/prof_run.py 250000000 2 > gpurun_out/vn_$t.log 2>&1
  echo "== $t"; grep -v "^==" gpurun_out/vn_$t.csv | awk -F'","' 'NR>1{print $5, $(NF-2), $NF}' | tr -d '"' 
done
cp /tmp/lib_default.so kmer_spans_b200/csrc/libkspans_cuda.so

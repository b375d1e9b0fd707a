#!/bin/bash
# usage: tools/variant_libs.sh <script.py> [args] -- runs the script once per prebuilt library in tools/bin/lib_*.so
cd "$(dirname "$0")/.."
cp kmer_spans_b200/csrc/libkspans_cuda.so /tmp/lib_default.so
for f in tools/bin/lib_*.so; do
  cp "$f" kmer_spans_b200/csrc/libkspans_cuda.so
  echo "== $f"
  python "$@" 2>&1 | tail -${TAILN:-4}
done
cp /tmp/lib_default.so kmer_spans_b200/csrc/libkspans_cuda.so

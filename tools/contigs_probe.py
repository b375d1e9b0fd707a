"""config-5-like input: many short contigs through the host-buffer entry point (staging path)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from kmer_spans_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seqs = api.SeqBatch.from_list(synth.contigs(n, seed=5, k=10))
bases = int(seqs.lens.sum())
ctx = api.Context(0)
for mode, thr in ((0, 0.75), (1, 0.0), (2, 0.0)):
    for rep in range(2):
        t0 = time.perf_counter()
        r = ctx.kmer_mode_regions(seqs, 10, mode, 100, 20.0, thr=thr, want_tables=False)
        dt = time.perf_counter() - t0
    print("contigs %d bases %.1f Mb mode %d: %.1f ms end to end (%.2f Gbases/s), %d spans, levels %s" % (
        n, bases / 1e6, mode, dt * 1e3, bases / dt / 1e9, len(r["pos"]), ctx.scan_stats()))

"""where does the end-to-end (host buffers) time go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kmer_spans_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
seq = synth.config2(n, 2)[0]
pinned = torch.from_numpy(seq).pin_memory().numpy()
ctx = api.Context(0)
counts_host = torch.empty(4 ** 12, dtype=torch.int32).pin_memory().numpy()
for rep in range(3):
    t0 = time.perf_counter(); ss = ctx.upload([pinned]); t1 = time.perf_counter()
    ss.free(); t2 = time.perf_counter()
    r = ctx.kmer_mode_regions([pinned], 12, 1, 100, 20.0, want_tables=False, counts_out=counts_host); t3 = time.perf_counter()
    r = ctx.kmer_mode_regions([seq], 12, 1, 100, 20.0, want_tables=False, counts_out=counts_host); t4 = time.perf_counter()
    print("upload %.1f ms  free %.1f ms  mode_regions(pinned) %.1f ms  mode_regions(pageable) %.1f ms" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))

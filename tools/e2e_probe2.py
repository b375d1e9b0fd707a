"""end-to-end (host buffers) breakdown: raw H2D, count only, full call with / without the count table D2H"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from kmer_spans_b200 import api, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
seq = synth.config2(n, 2)[0]
pinned_t = torch.from_numpy(seq).pin_memory()
pinned = pinned_t.numpy()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx = api.Context(0)
counts_host = torch.empty(4 ** 12, dtype=torch.int32).pin_memory().numpy()
def tm(f, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
print("raw H2D 250 MB pinned: %.2f ms" % tm(lambda: dev.copy_(pinned_t, non_blocking=True)))
d2h = torch.empty(4 ** 12, dtype=torch.int32, device="cuda"); hh = torch.empty(4 ** 12, dtype=torch.int32).pin_memory()
print("raw D2H 64 MiB pinned: %.2f ms" % tm(lambda: hh.copy_(d2h, non_blocking=True)))
print("upload only: %.2f ms" % tm(lambda: ctx.upload([pinned]).free()))
print("kmer_counts (pinned in, pageable out): %.2f ms" % tm(lambda: ctx.kmer_counts([pinned], 12, with_f=False)))
print("mode_regions no tables: %.2f ms" % tm(lambda: ctx.kmer_mode_regions([pinned], 12, 1, 100, 20.0, want_tables=False)))
print("mode_regions counts_out pinned: %.2f ms" % tm(lambda: ctx.kmer_mode_regions([pinned], 12, 1, 100, 20.0, want_tables=False, counts_out=counts_host)))
co = np.zeros(4 ** 12, np.int32); so = np.zeros(4 ** 12, np.float64)
print("mode_regions log2, counts + scores to pageable numpy: %.2f ms" % tm(lambda: ctx.kmer_mode_regions([pinned], 12, 1, 100, 20.0, want_tables=True)))
print("low_comp (rank), counts + ranks to pageable numpy (the R call): %.2f ms" % tm(lambda: ctx.kmer_low_comp_regions([seq], 12, 100, 20.0)))

"""device-resident ms per pass for each score mode (config-2 workload)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kmer_spans_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 12
seq = synth.config2(n, 2)[0]
ctx = api.Context(0)
ss = ctx.upload([seq])
counts = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
scores = torch.zeros(4 ** k, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ctx.set_profile(True)
for mode, thr, name in ((1, 0.0, "log2"), (0, 0.75, "rank thr=.75"), (2, 0.0, "sign"), (0, 0.5, "rank thr=.5")):
    for i in range(2):
        ctx.dev_pipeline(ss, k, mode, 100, 20.0, thr=thr, d_counts=counts.data_ptr(), d_scores=scores.data_ptr())
    ctx.profile(reset=True)
    ctx.timer_start()
    reps = 3
    for i in range(reps):
        r = ctx.dev_pipeline(ss, k, mode, 100, 20.0, thr=thr, d_counts=counts.data_ptr(), d_scores=scores.data_ptr())
    ms = ctx.timer_stop() / reps
    prof = ctx.profile(reset=True)
    lv, rv = ctx.scan_stats()
    print("%-14s %.3f ms/pass  %.1f Gbases/s  spans %d levels %d revisit_chunks %d | %s" % (
        name, ms, n / ms / 1e6, r["n_spans"], lv, rv, {a: round(b[0] / reps, 3) for a, b in prof.items()}))

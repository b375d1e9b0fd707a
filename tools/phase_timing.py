"""KS_EXP_TIMING builds: average cycles per tile spent in each phase of scan level 0."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kmer_spans_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
k = 12
seq = synth.config2(n, 2)[0]
ctx = api.Context(0)
ss = ctx.upload([seq])
counts = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
scores = torch.zeros(4 ** k, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
out = (C.c_uint64 * 16)()
ctx.lib.ks_ctx_debug_counters(ctx.h, out, 1)
for i in range(3):
    ctx.dev_pipeline(ss, k, mode, 100, 20.0, thr=0.75 if mode == 0 else 0.0, d_counts=counts.data_ptr(), d_scores=scores.data_ptr())
    if i == 0:
        ctx.lib.ks_ctx_debug_counters(ctx.h, out, 1)
ctx.lib.ks_ctx_debug_counters(ctx.h, out, 0)
names = ["load+decode+gather", "transform+warp scan+bar", "lookback / prefix", "barrier wait", "walk+ex+finish"]
for base, who in ((0, "thread 0 (warp 0)"), (8, "thread 255 (warp 7)")):
    nt = out[base + 5]
    print(who, "tiles", nt)
    tot = sum(out[base + i] for i in range(5))
    for i, nm in enumerate(names):
        print("   %-28s %8.0f cycles/tile  %5.1f%%" % (nm, out[base + i] / max(nt, 1), 100.0 * out[base + i] / max(tot, 1)))
    print("   total %.0f cycles/tile" % (tot / max(nt, 1)))

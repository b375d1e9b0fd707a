"""KS_EXP_TIMING builds: average cycles per tile (thread 0) in each phase of scan level 0."""
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
names = ["A: map + packed loads + decode + issue gathers", "C: wait counts + LUT", "L: transform + warp scan", "L: stash + barrier",
         "L: prefix + publish + barrier", "F: look-back (warp 0)", "F: barrier + unstash", "F: walk",
         "F: ex scan + publish + barriers", "F: finish entering", "next_tile (2 barriers + atomic)"]
ntiles = max(out[15], 1)
tot = sum(out[i] for i in range(11))
print("mode", mode, "tile fetches", out[15], " total %.0f cycles/tile" % (tot / ntiles))
for i, nm in enumerate(names):
    print("   %-34s %8.0f cycles/tile  %5.1f%%" % (nm, out[i] / ntiles, 100.0 * out[i] / max(tot, 1)))

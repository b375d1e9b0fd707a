"""Small driver for ncu: config-2 workload, a few passes of the device-resident pipeline.
usage: python tools/prof_run.py [n_bases] [passes] [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kmer_spans_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
k = 12
seq = synth.config2(n, 2)[0]
ctx = api.Context(0)
ss = ctx.upload([seq])
counts = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
scores = torch.zeros(4 ** k, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for i in range(passes):
    r = ctx.dev_pipeline(ss, k, mode, 100, 20.0, thr=0.75 if mode == 0 else 0.0, d_counts=counts.data_ptr(),
                         d_scores=scores.data_ptr())
print("spans", r["n_spans"], "launches", ctx.launches(), "levels", ctx.scan_stats())

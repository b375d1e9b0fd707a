"""device-resident ms per pass, log2 mode only (config-2 workload); one line"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kmer_spans_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 12
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
te = int(os.environ.get('KS_SYNTH_TANDEM_EVERY', '0'))
seq = synth.genome(n, 2, tandem_every=te, n_blocks=(5, 50_000)) if te else synth.config2(n, 2)[0]
ctx = api.Context(0)
ss = ctx.upload([seq])
counts = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
scores = torch.zeros(4 ** k, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ctx.set_profile(True)
kw = dict(thr=0.75 if mode == 0 else 0.0, d_counts=counts.data_ptr(), d_scores=scores.data_ptr())
for i in range(3):
    ctx.dev_pipeline(ss, k, mode, 100, 20.0, **kw)
ctx.profile(reset=True)
ctx.timer_start()
reps = 10
for i in range(reps):
    r = ctx.dev_pipeline(ss, k, mode, 100, 20.0, **kw)
ms = ctx.timer_stop() / reps
prof = ctx.profile(reset=True)
print("mode %d %.3f ms/pass  %.1f Gbases/s  spans %d | %s" % (
    mode, ms, n / ms / 1e6, r["n_spans"], {a: round(b[0] / reps, 3) for a, b in prof.items()}))

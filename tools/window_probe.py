"""windowed occurrence histograms at config-2 size: device-resident time with CUDA
events, end-to-end time through the host entry point, the CPU reference on a bounded sample, and a
full-size parity property (every window is counted once: column sums = number of windows)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from kmer_spans_b200 import api, synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
window = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
k = 2
kmers = [a + b for a in "ACTG" for b in "ACTG"]
seq = synth.config2(n)[0]
ctx = api.Context(0)
codes = np.array([ctx.lib.ks_kmer_code(x.encode(), k) for x in kmers], np.uint32)
ss = ctx.upload([seq])
d_dist = torch.empty((len(kmers), window + 1), dtype=torch.int32, device="cuda")
times = []
for rep in range(6):
    ctx.timer_start()
    ctx.dev_window_dist(ss, k, codes, window, d_dist.data_ptr())
    times.append(ctx.timer_stop())
dev_ms = float(np.median(times[1:]))
dist = d_dist.cpu().numpy()
# property: N-free stretches of the synthetic genome are known -> number of windows per k-mer
is_n = np.flatnonzero(seq == ord("N"))
cuts = np.concatenate(([-1], is_n, [n]))
runs = np.diff(cuts) - 1
want_windows = int(np.maximum(runs - window + 1, 0).sum())
assert (dist.sum(axis=1) == want_windows).all(), (dist.sum(axis=1), want_windows)
# every base pair of a window belongs to exactly one dinucleotide: sum_i sum_c c * dist[i, c] = windows * (window - 1)
assert int((dist * np.arange(window + 1)).sum()) == want_windows * (window - k + 1)
for rep in range(2):  # the second call re-uses the pinned staging and device buffers
    t0 = time.perf_counter()
    r = ctx.window_kmer_dist([seq], kmers, window, freq=False)
    e2e_ms = (time.perf_counter() - t0) * 1e3
assert np.array_equal(r["dist"].T, dist)
out = dict(bases=n, k=k, kmers=len(kmers), window=window, device_ms=dev_ms, device_gbases_s=n / dev_ms / 1e6,
           e2e_ms=e2e_ms, e2e_gbases_s=n / e2e_ms / 1e6, windows=want_windows)
try:
    from oracle.ksoracle import Ref
    ref = Ref()
    m = min(n, 5_000_000)
    sample = seq[:m].tobytes()
    t0 = time.perf_counter()
    a = ref.call_window_dist([sample], [x.encode() for x in kmers], k, window, 0)
    cpu_s = time.perf_counter() - t0
    got = ctx.window_kmer_dist([sample], kmers, window, freq=False)
    assert np.array_equal(got["dist"].T, a["dist"]), "sample parity"
    out.update(cpu_sample_bases=m, cpu_s=cpu_s, cpu_gbases_s=m / cpu_s / 1e9, cpu_kind="reference", cpu_cores=1)
except FileNotFoundError:
    pass
print(json.dumps(out))

"""transition-score scan at config-2 size: device-resident time (CUDA events),
levels, CPU reference on a bounded sample with parity, and size-independent properties at full size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from kmer_spans_b200 import api, synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 50
nk = 4 ** k
seq = synth.config2(n)[0]
ctx = api.Context(0)
# log-ratio style tables from the sequence's own spectrum: over-represented k-mers score up
counts = ctx.kmer_counts([seq], k, with_f=False)["counts"].astype(np.float64)
f = (counts + 1.0) / (counts.sum() + nk)
init = np.log2(f * nk)
trans = np.log2(f * nk) - 0.25
ss = ctx.upload([seq])
d_init = torch.from_numpy(init).cuda()
d_trans = torch.from_numpy(trans).cuda()
from kmer_spans_b200._lib import KsSpans
import ctypes as C
times = []
for rep in range(5):
    sp = KsSpans()
    ctx.timer_start()
    ctx._ck(ctx.lib.ks_dev_tr_lr_regions(ctx.h, ss.h, k, C.c_void_p(d_init.data_ptr()), C.c_void_p(d_trans.data_ptr()),
                                         min_len, C.byref(sp), None))
    times.append(ctx.timer_stop())
    pos, score = api._spans_to_numpy(ctx.lib, sp)
dev_ms = float(np.median(times[1:]))
levels, revisit = ctx.scan_stats()
# properties: 1-based ids, ascending starts, width rule, positive scores, inside the sequence
assert (pos[:, 0] == 1).all() and (np.diff(pos[:, 1]) > 0).all()
assert ((pos[:, 2] - pos[:, 1]) >= min_len).all() and (score[:, 0] > 0).all() and (score[:, 1] == 0).all()
assert pos[:, 1].min() >= 1 and pos[:, 2].max() <= n
out = dict(bases=n, k=k, min_length=min_len, device_ms=dev_ms, device_gbases_s=n / dev_ms / 1e6, spans=int(len(pos)),
           levels=int(levels), revisited_positions=int(revisit) * 16)
try:
    from oracle.ksoracle import Ref, Oracle
    ref, orc = Ref(), Oracle()
    m = min(n, 10_000_000)
    sample = seq[:m].tobytes()
    kms = [orc.kmer_seq(k, c).encode() for c in range(nk)]
    t0 = time.perf_counter()
    a = ref.call_tr_lr([sample], k, min_len, kms, init, trans)
    cpu_s = time.perf_counter() - t0
    g = ctx.lr_regions([sample], (k, min_len), kms, init, trans)
    assert g["pos"].tolist() == a["pos"].tolist(), "sample parity (coordinates)"
    np.testing.assert_allclose(g["score"], a["score"], rtol=1e-9)
    out.update(cpu_sample_bases=m, cpu_s=cpu_s, cpu_gbases_s=m / cpu_s / 1e9, cpu_kind="reference", cpu_cores=1,
               sample_spans=int(len(a["pos"])))
except FileNotFoundError:
    pass
print(json.dumps(out))

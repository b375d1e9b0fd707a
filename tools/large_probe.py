"""BASELINE.json configs[3] probe: k = 21 on the first chromosomes of the config-3 genome (hash-table counting,
rank over the k-mers that occur, scan).  usage: python tools/large_probe.py [gigabases] [k]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kmer_spans_b200 import api, synth  # noqa: E402

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
k = int(sys.argv[2]) if len(sys.argv) > 2 else 21
element = synth.random_bases(np.random.default_rng(0xE1E), 300)
seqs, tot = [], 0
for i, mb in enumerate(synth.HUMAN_MB):
    if tot >= gb * 1e9:
        break
    seqs.append(synth.genome(mb * 1_000_000, 100 + i, element=element))
    tot += mb * 1_000_000
ctx = api.Context(0)
ss = ctx.upload(seqs)
out = {"bases": tot, "sequences": len(seqs), "k": k}
for rep in range(2):
    ctx.set_profile(True)
    ctx.profile(reset=True)
    ctx.timer_start()
    r = ctx.dev_large_regions(ss, k, 0, 100, 20.0, thr=0.75)
    ms = ctx.timer_stop()
    prof = ctx.profile(reset=True)
    free, total = torch.cuda.mem_get_info()
    out["pass%d" % rep] = {"ms": ms, "gbases_per_s": tot / ms / 1e6, "distinct": r["nd"], "words": r["n"],
                           "spans": r["n_spans"], "stages_ms": {a: b[0] for a, b in prof.items() if b[1]},
                           "device_memory_used_gb": (total - free) / 1e9}
# algorithmic bytes of the hash path: 1.5 N + 24 N
peak = 6544.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ms = out["pass1"]["ms"]
out["roofline_frac_bytes_hash"] = (25.5 * tot) / (ms * 1e-3) / 1e9 / peak
print(json.dumps(out, indent=1))

import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from kmer_spans_b200 import api, synth
n=250_000_000; window=1000; k=2
kmers=[a+b for a in "ACTG" for b in "ACTG"]
seq=synth.config2(n)[0]
ctx=api.Context(0)
codes=np.array([ctx.lib.ks_kmer_code(x.encode(),k) for x in kmers],np.uint32)
ss=ctx.upload([seq])
d=torch.empty((16,window+1),dtype=torch.int32,device="cuda")
for i in range(2):
    ctx.dev_window_dist(ss,k,codes,window,d.data_ptr()); ctx.sync()

"""torchrun check of the split (intra-sequence) multi-GPU path against the oracle:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/split_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from kmer_spans_b200 import api, synth
from kmer_spans_b200 import dist as ksd

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ctx = api.Context(int(os.environ["LOCAL_RANK"]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
seqs = [synth.genome(n, 2, n_blocks=(5, 50_000)), synth.genome(n // 7, 3)]
ok = True
for k, mode, thr in ((12, 0, 0.75), (12, 1, 0.0), (10, 2, 0.0)):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    r = ksd.run_split_nccl(ctx, dist, seqs, k, mode, 100, 20.0, thr=thr)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        from oracle.ksoracle import Oracle
        want = Oracle().mode_regions([s.tobytes() for s in seqs], k, mode, 100, 20.0, thr=thr)
        same = r["pos"].tolist() == want["pos"].tolist() and np.allclose(r["score"], want["score"], rtol=1e-9)
        same = same and (r["counts"].cpu().numpy() == want["counts"]).all()
        print("world %d k=%d mode=%d: %d spans, match=%s, %.1f ms" % (world, k, mode, len(want["pos"]), same, dt * 1e3), flush=True)
        ok = ok and same
if rank == 0:
    print("SPLIT_CHECK", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()

"""N ranks (torchrun), whole sequences per rank: dist.run_sharded (count -> NCCL all-reduce on the ctx stream
-> scores -> scan) against the CPU oracle on the union of the sequences."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from kmer_spans_b200 import dist as ksd, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k, n_seq = 10, 3 * world
seqs = [synth.genome(2_000_000 + 100_000 * i, 40 + i) for i in range(n_seq)]
plan = ksd.plan_shards([len(s) for s in seqs], world)
mine = plan[rank]
stages = ksd.GpuStages(local)
ok = True
for mode, thr in ((1, 0.0), (0, 0.75), (2, 0.0)):
    for rep in range(3):  # repeated: a race between the collective and the kernels would show as flaky results
        r = ksd.run_sharded(stages, dist, [seqs[i] for i in mine], mine, k, mode, 100, 20.0, thr=thr)
        if rank == 0:
            from oracle.ksoracle import Oracle
            want = Oracle().mode_regions([s.tobytes() for s in seqs], k, mode, 100, 20.0, thr=thr)
            same = r["pos"].tolist() == want["pos"].tolist() and np.allclose(r["score"], want["score"], rtol=1e-9) \
                and r["n"] == want["n"] and np.array_equal(r["counts"].cpu().numpy(), want["counts"])
            ok = ok and same
            print("mode %d rep %d: %d spans, n %.0f, identical to the oracle: %s" % (mode, rep, len(r["pos"]), r["n"], same))
dist.barrier()
if rank == 0:
    print("SHARDED CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()

"""BASELINE.json configs[2]: 24 human-like chromosomes (3.1 Gb), k=13, weighted-rank score (thr 0.75), sharded by
whole chromosomes over the ranks (LPT), count tables all-reduced over NCCL.
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config3_run.py [scale]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from kmer_spans_b200 import synth
from kmer_spans_b200 import dist as ksd

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
os.environ["NCCL_DEBUG"] = "WARN"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lens = [int(mb * 1_000_000 * scale) for mb in synth.HUMAN_MB]
plan = ksd.plan_shards(lens, world)
mine = plan[rank]
element = synth.random_bases(np.random.default_rng(0xE1E), 300)
t0 = time.perf_counter()
seqs = [synth.genome(lens[i], 100 + i, element=element) for i in mine]
tgen = time.perf_counter() - t0
stages = ksd.GpuStages(lr)
K, MODE, THR = 13, 0, 0.75
res = None
times = []
for rep in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = ksd.run_sharded(stages, dist if world > 1 else None, seqs, mine, K, MODE, 100, 20.0, thr=THR)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
if rank == 0:
    pos = res["pos"]
    ok_sorted = bool(np.all(np.lexsort((pos[:, 1], pos[:, 0])) == np.arange(len(pos))))
    out = {"config": "BASELINE configs[2] x %.3g" % scale, "world": world, "bases": int(sum(lens)), "k": K,
           "mode": "rank thr 0.75", "spans": int(len(pos)), "sorted": ok_sorted, "n_words": res["n"],
           "counts_checksum_ok": bool(int(res["counts"].to(torch.int64).sum().item()) == int(res["n"])),
           "seconds_end_to_end_from_host": times, "gbases_per_s_best": sum(lens) / min(times) / 1e9,
           "shard_bases": [int(sum(lens[i] for i in p)) for p in plan], "gen_s_rank0": tgen}
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()

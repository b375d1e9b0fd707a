"""CPU tier: the multi-GPU host logic (shard plan, count all_reduce, span merge) with
world_size 2 over gloo.  The per-rank compute stages are stand-ins backed by the oracle -- test
code only; the product stages (dist.GpuStages) run the CUDA kernels."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_shards_balanced():
    from kmer_spans_b200 import synth
    from kmer_spans_b200.dist import plan_shards
    lens = [mb * 1_000_000 for mb in synth.HUMAN_MB]
    for world in (1, 2, 4, 8):
        plan = plan_shards(lens, world)
        assert sorted(i for p in plan for i in p) == list(range(len(lens)))
        loads = [sum(lens[i] for i in p) for p in plan]
        assert max(loads) <= 1.08 * sum(lens) / world
    assert plan_shards([5, 1], 4) == [[0], [1], [], []]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from kmer_spans_b200 import dist as ksd
    from oracle.ksoracle import Oracle
    from tests.test_oracle import planted
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    rng = np.random.default_rng(42)
    seqs = [planted(rng, int(rng.integers(200, 9000))) for _ in range(9)] + [b"AC"]
    k, mode, mw, ms, thr = 6, 0, 10, 3, 0.7

    class OracleStages:  # test stand-in with the GpuStages interface
        def load(self, s):
            self.seqs = s

        def alloc_tables(self, k):
            self.counts = torch.zeros(4 ** k, dtype=torch.int32)
            return self.counts

        def count(self, k):
            if not self.seqs:
                return 0.0
            n, c = orc.kmer_counts(self.seqs, k)
            self.counts.copy_(torch.from_numpy(c))
            return n

        def scores_from_counts(self, k, total, mode, param):
            self.W = orc.scores(self.counts.numpy(), k, total, mode, param)

        def scan(self, k, thr, min_w, min_score, fetch=True):
            if not self.seqs:
                return np.zeros((0, 3), np.int32), np.zeros((0, 2))
            r = orc.kmer_regions(self.seqs, k, self.W - thr, min_w, min_score)
            return r["pos"], r["score"]

    plan = ksd.plan_shards([len(s) for s in seqs], world)
    mine = plan[rank]
    res = ksd.run_sharded(OracleStages(), dist, [seqs[i] for i in mine], mine, k, mode, mw, ms, thr=thr)
    if rank == 0:
        want = orc.low_comp(seqs, k, mw, ms, thr)
        ok = (res["counts"].numpy() == want["counts"]).all() and res["n"] == want["n"][0] \
            and res["pos"].tolist() == want["pos"].tolist() and len(want["pos"]) > 3
        q.put(bool(ok))
    dist.destroy_process_group()


def test_sharded_pipeline_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

"""Multi-GPU composition of the hot path: one process per GPU (torch.distributed, NCCL over NVLink).

The path shards by whole sequences (chromosomes / contigs), balanced by length:
  count      local int32[4^k] table per rank             -> all_reduce(SUM) over NCCL   (real exchange)
  n_words    local scalar                                -> all_reduce(SUM)
  scores     recomputed redundantly from the reduced table (deterministic, identical on all ranks)
  scan       each rank scans its own sequences; runs never cross sequences, so no carry crosses
             a shard boundary and no halo is needed
  spans      per-rank lists keep the ORIGINAL 0-based seq_id; rank 0 gathers and orders them
This mirrors what kmer_low_comp_regions does over all sequences of one call
(/root/reference/src/kmer_spans.c:592-612: one shared count table, then a scan per sequence).

The stage functions are injectable so that the sharding / reduction / merge logic can be exercised
with world_size-2 gloo on CPU (tests/test_dist.py) -- the stand-in stages there are test code.
"""
import os

import numpy as np


def plan_shards(lengths, world):
    """Longest-processing-time assignment of whole sequences to ranks.
    Returns a list (per rank) of sequence indices in ascending order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(x) for x in out]


def merge_spans(parts):
    """parts: list of (pos (n,3) with GLOBAL seq ids, score (n,2)); returns them ordered by
    (seq_id, start) -- the reference's discovery order."""
    pos = np.concatenate([p for p, _ in parts]) if parts else np.zeros((0, 3), np.int32)
    score = np.concatenate([s for _, s in parts]) if parts else np.zeros((0, 2))
    if len(pos):
        o = np.lexsort((pos[:, 1], pos[:, 0]))
        pos, score = pos[o], score[o]
    return pos, score


class GpuStages:
    """The product stages: kernels of libkspans_cuda.so on this rank's GPU; tables live in torch
    CUDA tensors so that torch.distributed can reduce them in place."""

    def __init__(self, device):
        import torch
        from . import api
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.ctx = api.Context(device)
        self.ctx.side_table(True)  # the score table is read after the scan: let it be written next to it
        self.ss = None

    def load(self, seqs):
        if self.ss is not None:
            self.ss.free()
        self.ss = self.ctx.upload(seqs) if len(seqs) else None

    def alloc_tables(self, k, dist=None):
        """count / score tables, kept between calls.  With a process group (world > 1) the count table and the
        word count live in ONE symmetric-memory buffer mapped by every rank, so that reduce_counts() can sum
        them over NVLink / NVSwitch peer memory (csrc/ks_xgpu.cuh); if that cannot be set up the tables are
        plain tensors and reduce_counts() uses NCCL."""
        t = self.torch
        world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        key = (k, world, id(dist.group.WORLD) if world > 1 else 0)
        if getattr(self, "_tables_k", None) == key:
            return self.counts
        n = 4 ** k
        self._peer = None
        if world > 1 and os.environ.get("KS_PEER_SUM", "1") != "0":
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(n + 2, dtype=t.int32, device=self.device)
                hdl = symm.rendezvous(buf, dist.group.WORLD)
                buf.zero_()
                mc = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
                if os.environ.get("KS_PEER_SUM", "1") == "p2p":  # measurement: plain peer loads / stores
                    mc = 0
                self._peer = dict(hdl=hdl, buf=buf, ptrs=[int(p) for p in hdl.buffer_ptrs], mc=mc,
                                  rank=dist.get_rank(), n_u64=n // 2 + 1)
                self.counts = buf[:n]
                self.nwords = buf[n:n + 2].view(t.int64)
            except Exception as e:  # no symmetric memory on this system: NCCL does the sum
                self._peer = None
                self._peer_error = repr(e)
            # the ranks must agree on the path: one that fell back alone would sit in an NCCL all-reduce while the
            # others wait at the symmetric-memory barrier.  [peer path ok, multicast ok], minimum over the ranks
            flag = t.tensor([1 if self._peer is not None else 0,
                             1 if (self._peer is not None and self._peer["mc"]) else 0],
                            dtype=t.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok_peer, ok_mc = (int(v) for v in flag.tolist())
            if not ok_peer:
                if self._peer is not None:
                    self._peer_error = "another rank could not map the symmetric buffer"
                self._peer = None
            elif not ok_mc:
                self._peer["mc"] = 0
        if self._peer is None:
            self.counts = t.zeros(n, dtype=t.int32, device=self.device)
            self.nwords = t.zeros(1, dtype=t.int64, device=self.device)
        self.scores = t.empty(n, dtype=t.float64, device=self.device)
        t.cuda.synchronize(self.device)
        self._tables_k = key
        return self.counts

    def reduce_counts(self, dist, events=None):
        """sum of self.counts / self.nwords over all ranks, ordered on the ctx stream.  events: optional list that
        receives four CUDA events (before, after the first barrier, after the sum kernel, after the second barrier)"""
        t = self.torch

        def mark():
            if events is not None:
                e = t.cuda.Event(enable_timing=True)
                e.record(self.stream())
                events.append(e)
        with t.cuda.stream(self.stream()):
            mark()
            if self._peer is not None:
                p = self._peer
                p["hdl"].barrier(channel=0)   # every rank has finished counting
                mark()
                self.ctx.dev_xsum(p["ptrs"], p["rank"], p["mc"], p["n_u64"])
                mark()
                p["hdl"].barrier(channel=1)   # every slice has been written everywhere
            else:
                mark()
                dist.all_reduce(self.counts, op=dist.ReduceOp.SUM)
                dist.all_reduce(self.nwords, op=dist.ReduceOp.SUM)
                mark()
            mark()

    def peer_sum_kind(self):
        if getattr(self, "_peer", None) is None:
            return "nccl all_reduce"
        return "one kernel over symmetric memory, " + ("NVSwitch multicast (multimem.ld_reduce / multimem.st)"
                                                       if self._peer["mc"] else "peer loads / stores")

    def load_and_count(self, seqs, k):
        """upload into the resident set (buffers re-used) with the pack+count pass running behind the copies;
        counts -> self.counts, words -> self.nwords; nothing synchronised"""
        if not len(seqs):
            self.load(seqs)
            self.count_async(k)
            return
        if self.ss is None:
            self.ss = self.ctx.upload(seqs)
            self.count_async(k)
            return
        self.ss.reupload(seqs, k, self.counts.data_ptr(), self.nwords.data_ptr())

    def count(self, k):
        if self.ss is None:
            self.counts.zero_()
            self.torch.cuda.synchronize(self.device)
            return 0.0
        return self.ctx.dev_count(self.ss, k, self.counts.data_ptr())

    # ---- the same stages without host round trips (one stream carries kernels and collectives) ----
    def stream(self):
        """the ctx stream as a torch stream: collectives issued under it are ordered with the kernels"""
        if getattr(self, "_ext", None) is None:
            self._ext = self.torch.cuda.ExternalStream(int(self.ctx.stream), device=self.device)
        return self._ext

    def count_async(self, k):
        """counts into self.counts, words into self.nwords (int64 tensor, 1 element), nothing synchronised"""
        t = self.torch
        if getattr(self, "nwords", None) is None:
            self.nwords = t.zeros(1, dtype=t.int64, device=self.device)
            t.cuda.synchronize(self.device)
        if self.ss is None:
            with t.cuda.stream(self.stream()):
                self.counts.zero_()
                self.nwords.zero_()
            return
        self.ctx.dev_count_async(self.ss, k, self.counts.data_ptr(), self.nwords.data_ptr())

    def scores_from_counts_dev(self, k, mode, param):
        """scores from self.counts / self.nwords as they are on the ctx stream; returns the total"""
        self.mode = mode
        return self.ctx.dev_scores_devtotal(k, self.counts.data_ptr(), self.nwords.data_ptr(), mode,
                                            self.scores.data_ptr(), param)

    def scores_from_counts(self, k, total, mode, param):
        self.torch.cuda.synchronize(self.device)  # the reduced table is ready before our stream reads it
        self.mode = mode
        self.ctx.dev_scores(k, self.counts.data_ptr(), total, mode, self.scores.data_ptr(), param)

    def scan(self, k, thr, min_w, min_score, fetch=True):
        if self.ss is None:
            return (0, None) if not fetch else (np.zeros((0, 3), np.int32), np.zeros((0, 2)))
        if self.mode in (1, 2):  # score is a function of the count: gather counts + LUT
            r = self.ctx.dev_scan_counts(self.ss, k, self.counts.data_ptr(), thr, min_w, min_score, fetch_spans=fetch)
        else:
            r = self.ctx.dev_scan(self.ss, k, self.scores.data_ptr(), thr, min_w, min_score, 0, fetch_spans=fetch)
        if not fetch:
            return r["n_spans"], None
        return r["pos"], r["score"]


def run_sharded(stages, dist, seqs_local, local_ids, k, mode, min_w, min_score, thr=0.0, param=float("nan"),
                gather=True, fetch=True, counts_host=None):
    """One pass of count -> all_reduce -> scores -> scan over this rank's sequences.
    stages: GpuStages (or a test stand-in with the same methods); dist: torch.distributed or None.
    local_ids: original indices of seqs_local.  Returns dict(n, counts, pos, score) (spans on rank 0
    when gather, else this rank's).  counts_host: optional pinned int32 tensor that receives the reduced count
    table; the copy runs on a side stream behind the score stage and overlaps the scan."""
    import torch
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    if hasattr(stages, "load_and_count"):
        # kernels and collectives on one stream: no host synchronisation until the score stage reads back
        counts = stages.alloc_tables(k, dist)
        stages.load_and_count(seqs_local, k)
        if world > 1:
            stages.reduce_counts(dist)   # the one real exchange of this path
        total = stages.scores_from_counts_dev(k, mode, param)
    else:
        stages.load(seqs_local)
        counts = stages.alloc_tables(k)
        n_local = stages.count(k)
        n_t = torch.tensor([n_local], dtype=torch.float64, device=counts.device)
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
            dist.all_reduce(n_t, op=dist.ReduceOp.SUM)
        total = float(n_t.item())
        stages.scores_from_counts(k, total, mode, param)
    side = None
    if counts_host is not None and hasattr(stages, "stream"):
        side = getattr(stages, "_side", None)
        if side is None:
            side = stages._side = torch.cuda.Stream(device=counts.device)
        ev = torch.cuda.Event()
        ev.record(stages.stream())
        side.wait_event(ev)
        with torch.cuda.stream(side):
            counts_host.copy_(counts, non_blocking=True)
    elif counts_host is not None:
        counts_host.copy_(counts)
    pos, score = stages.scan(k, thr, min_w, min_score, fetch=fetch)
    if side is not None:
        side.synchronize()
    if not fetch:
        return dict(n=total, counts=counts, n_spans=pos)
    ids = np.asarray(local_ids, np.int32)
    if len(pos):
        pos = pos.copy()
        pos[:, 0] = ids[pos[:, 0]]
    if world > 1 and gather:
        parts = [None] * world
        dist.all_gather_object(parts, (pos, score))
        pos, score = merge_spans(parts)
    return dict(n=total, counts=counts, pos=pos, score=score)


# --------------------------------------------------------------------------------------------------
# One sequence set split ACROSS GPUs at arbitrary chunk boundaries (a chromosome longer than a fair
# share): every rank uploads and packs the whole set, counts and scans only its dense chunk range, and
# the scan is stitched exactly through two 48-byte carries per rank.
def shard_range(total_chunks, world, rank):
    per = -(-total_chunks // world)
    c0 = min(rank * per, total_chunks)
    return c0, min(per, total_chunks - c0)


def run_split(ctx, seqs, k, mode, min_w, min_score, thr, param, rank, world, all_gather_bytes, all_reduce_counts):
    """count(range) -> all_reduce -> scores -> sharded scan with carry exchange.  Every rank plans the same cut of
    the layout of ALL sequences (ks_plan_shard) and uploads only its window: its own range plus the head of the
    sequence the range starts in.
    all_gather_bytes(blob48) -> [blob48 of every rank];  all_reduce_counts(torch int32 tensor, n) -> total n.
    Returns this rank's spans (global coordinates; a span is reported where it closes)."""
    import torch
    from . import api
    lens = [len(s) for s in seqs]
    c0, cn, lo, hi = api.plan_shard(lens, world, rank)
    ss = ctx.upload_window(seqs, lo, hi)
    dev = torch.device("cuda", torch.cuda.current_device())
    counts = torch.zeros(4 ** k, dtype=torch.int32, device=dev)
    scores = torch.empty(4 ** k, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    n = ctx.dev_count_range(ss, k, c0, cn, counts.data_ptr())
    total = all_reduce_counts(counts, n)
    torch.cuda.synchronize()
    ctx.dev_scores(k, counts.data_ptr(), total, mode, scores.data_ptr(), param)

    def exchange(what, mine):
        return api.fold_carry(what, all_gather_bytes(mine), rank)

    if mode == 0:  # rank order left on the context by dev_scores: 4-byte position gather
        r = ctx.dev_scan_ranks_shard(ss, k, thr, min_w, min_score, c0, cn, exchange)
    else:
        use_counts = mode in (1, 2)
        r = ctx.dev_scan_shard(ss, k, counts.data_ptr() if use_counts else scores.data_ptr(), thr, min_w, min_score,
                               c0, cn, exchange, use_counts=use_counts)
    r["n"] = total
    r["counts"] = counts
    r["window_bytes"] = hi - lo
    ss.free()
    return r


def run_split_nccl(ctx, dist, seqs, k, mode, min_w, min_score, thr=0.0, param=float("nan"), gather=True):
    """run_split with torch.distributed: NCCL all_reduce of the count table, the two 48-byte carries through
    an NCCL all-gather of a device tensor (once each per scan)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()

    dev = torch.device("cuda", torch.cuda.current_device())
    gathered = torch.empty(48 * world, dtype=torch.uint8, device=dev)

    def all_gather_bytes(b):  # 48-byte carries through one NCCL all-gather of a device tensor
        mine = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
        dist.all_gather_into_tensor(gathered, mine)
        flat = gathered.cpu().numpy().tobytes()
        return [flat[48 * i:48 * (i + 1)] for i in range(world)]

    def all_reduce_counts(t, n):
        nt = torch.tensor([n], dtype=torch.float64, device=t.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(nt, op=dist.ReduceOp.SUM)
        return float(nt.item())

    r = run_split(ctx, seqs, k, mode, min_w, min_score, thr, param, rank, world, all_gather_bytes, all_reduce_counts)
    if gather:
        parts = [None] * world
        dist.all_gather_object(parts, (r["pos"], r["score"]))
        r["pos"], r["score"] = merge_spans(parts)
    return r


class SplitRun:
    """ONE sequence set cut across the ranks at chunk boundaries, resident as window sets (ks_plan_shard /
    ks_seqset_upload_window): what BASELINE.json configs[2] (24 chromosomes over 8 GPUs) needs.  step() runs
    count(range) -> sum of the tables (peer memory / NCCL, on the ctx stream) -> scores -> sharded scan; the two
    48-byte carries of a scan travel through an NCCL all-gather of a device tensor."""

    def __init__(self, stages, dist, seqs):
        import torch
        from . import api
        self.torch, self.api, self.stages, self.ctx = torch, api, stages, stages.ctx
        self.dist = dist if dist is not None and dist.is_initialized() and dist.get_world_size() > 1 else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        lens = seqs.lens if hasattr(seqs, "lens") else [len(x) for x in seqs]
        self.c0, self.cn, self.lo, self.hi = api.plan_shard(lens, self.world, self.rank)
        self.seqs = seqs
        self.ss = self.ctx.upload_window(seqs, self.lo, self.hi)
        self.bases = int(np.sum(lens))
        if self.dist:
            self._gathered = torch.empty(48 * self.world, dtype=torch.uint8, device=stages.device)

    def reload(self):
        """upload the window again (end-to-end timing)"""
        self.ss.free()
        self.ss = self.ctx.upload_window(self.seqs, self.lo, self.hi)

    def _exchange(self, what, mine):
        if not self.dist:
            return self.api.fold_carry(what, [mine], 0)
        t = self.torch
        mine_t = t.frombuffer(bytearray(mine), dtype=t.uint8).to(self.stages.device)
        self.dist.all_gather_into_tensor(self._gathered, mine_t)
        flat = self._gathered.cpu().numpy().tobytes()
        return self.api.fold_carry(what, [flat[48 * i:48 * (i + 1)] for i in range(self.world)], self.rank)

    def step(self, k, mode, thr, min_w, min_score, param=float("nan"), fetch=False, gather_scores=False):
        """gather_scores: in rank mode with the sliced score stage every rank holds only its slice of the rank table
        (an output, not an input of the scan); True all-gathers it so that every rank holds the whole table"""
        st, ctx = self.stages, self.ctx
        st.alloc_tables(k, self.dist)
        if getattr(st, "nwords", None) is None:
            st.count_async(k)  # allocates the word-count tensor
        ctx.dev_count_range_async(self.ss, k, self.c0, self.cn, st.counts.data_ptr(), st.nwords.data_ptr())
        if self.dist:
            st.reduce_counts(self.dist)
        n = 4 ** k
        if mode == 0 and self.dist and n % self.world == 0 and os.environ.get("KS_NO_SLICED_RANK") is None:
            total = self._scores_rank_sliced(k, gather_scores)
        else:
            total = st.scores_from_counts_dev(k, mode, param)
        if mode == 0:
            r = ctx.dev_scan_ranks_shard(self.ss, k, thr, min_w, min_score, self.c0, self.cn, self._exchange)
        else:
            use_counts = mode in (1, 2)
            r = ctx.dev_scan_shard(self.ss, k, st.counts.data_ptr() if use_counts else st.scores.data_ptr(), thr,
                                   min_w, min_score, self.c0, self.cn, self._exchange, use_counts=use_counts)
        r["n"] = total
        return r

    def _scores_rank_sliced(self, k, gather_scores=False):
        """weighted-rank score stage with every rank sorting and ranking only its slice of the k-mer index space;
        the slices of the rank table and of the rank-order positions are all-gathered over NCCL on the ctx stream"""
        t, st, ctx, dist = self.torch, self.stages, self.ctx, self.dist
        n = 4 ** k
        ctx.sync()
        total = float(int(st.nwords.item()))
        st.mode = 0
        if total == 0:
            return st.scores_from_counts_dev(k, 0, float("nan"))

        def gather(blob):
            mine = t.frombuffer(bytearray(blob), dtype=t.uint8).to(st.device)
            out = t.empty(len(blob) * self.world, dtype=t.uint8, device=st.device)
            dist.all_gather_into_tensor(out, mine)
            flat = out.cpu().numpy().tobytes()
            return [flat[len(blob) * i:len(blob) * (i + 1)] for i in range(self.world)]

        ctx.dev_scores_rank_sliced(k, st.counts.data_ptr(), total, self.rank, self.world, gather, st.scores.data_ptr())
        per = n // self.world
        lo = self.rank * per

        class _Dev:  # the ctx's rank-order position table as a torch tensor (CUDA array interface)
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        pos = t.as_tensor(_Dev(ctx.rank_positions_ptr(), n), device=st.device)
        with t.cuda.stream(st.stream()):
            if gather_scores:
                dist.all_gather_into_tensor(st.scores, st.scores[lo:lo + per].clone())
            dist.all_gather_into_tensor(pos, pos[lo:lo + per].clone())
        return total

    def free(self):
        self.ss.free()

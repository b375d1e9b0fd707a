#!/usr/bin/env python
"""bench.py -- Gbases/s of count + score + scan + spans (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)

Workload (config.workload): BASELINE.json configs[1] -- synthetic 250 Mb chromosome-scale sequence
with planted tandem / interspersed repeats and 5 N blocks, k=12, log2(f/f_med) score, min_width 100,
min_score 20, one such sequence PER GPU (weak scaling; count tables all-reduced over NCCL).
One step = one pass of the whole hot path over the resident sequence(s).

  value  device-resident: sequence already in HBM, results (count table, score table, ordered span
         list) left in HBM; timed with CUDA events on the launching stream, max over ranks.
  e2e    through the host-buffer C-ABI call a user makes (ks_kmer_mode_regions): pinned host
         sequence -> H2D -> pipeline -> D2H of the count table and the spans, wall clock.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K = 12
MODE_LOG2 = 1
MIN_W, MIN_SCORE, THR = 100, 20.0, 0.0
N_BASES = 250_000_000
METRIC = "Gbases/s count+score+scan+spans"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full summary, or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU through NVML while the timed region runs"""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    m = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    m = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if m & bit:
                        self.reasons.add(name)
                time.sleep(0.002)
        except Exception as e:  # NVML unavailable: report it instead of inventing numbers
            self.err = repr(e)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------------------
def cpu_reference_pass(seq_bytes, k=K):
    """The reference's CPU path for this workload on ONE host thread (it has no threads):
    sequence_kmer_count and kmer_regions are the UNMODIFIED reference (oracle/_ref); the log2 weight
    table, which the reference computes in R, is the oracle's restatement (kso_scores).
    Returns (seconds, n_spans, kind)."""
    from oracle.ksoracle import Oracle
    orc = Oracle()
    try:
        from oracle.ksoracle import Ref
        ref = Ref()
        kind = "reference"
    except Exception:
        ref = None
        kind = "port"
    t0 = time.perf_counter()
    if ref is not None:
        counts = np.zeros(4 ** k, np.int32)
        n = ref.sequence_kmer_count(seq_bytes, k, counts)
        W = orc.scores(counts, k, float(n), MODE_LOG2)
        pos, sc, _ = ref.kmer_regions_core([seq_bytes], k, W, THR, MIN_W, MIN_SCORE)
        nsp = len(pos)
    else:
        r = orc.mode_regions([seq_bytes], k, MODE_LOG2, MIN_W, MIN_SCORE, thr=THR)
        nsp = len(r["pos"])
    return time.perf_counter() - t0, nsp, kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    from kmer_spans_b200 import synth
    per_step = 150.0 / max(1, args.steps + args.warmup)
    sample = int(min(50e6, max(2e6, per_step * 9e6)))
    sample = min(sample, args.n_bases)
    # the SAME bytes the GPU arm's rank 0 runs on: a prefix of config2(n_bases, seed 2)
    seq = synth.config2(args.n_bases, 2)[0][:sample].tobytes()
    kind = "reference"
    for _ in range(args.warmup):
        _, _, kind = cpu_reference_pass(seq)
    t = 0.0
    for _ in range(args.steps):
        dt, _, kind = cpu_reference_pass(seq)
        t += dt
    val = sample * args.steps / t / 1e9
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gbases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": 1, "kind": kind,
                         "sample": "first %d bases of the %d-base config-2 sequence (seed 2) the GPU arm runs on "
                                   "(same bytes), k=12, log2 mode, whole CPU path per step; the reference is "
                                   "single-threaded" % (sample, args.n_bases)},
        "same_input": "prefix of the GPU arm's rank-0 sequence",
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(args, world, exchange="NCCL"):
    return {"workload": "BASELINE.json configs[1]: synthetic %d Mb sequence per GPU with planted tandem and "
                        "interspersed repeats, 5 N blocks; k=12; log2(f/f_med) score; min_width 100; min_score 20"
                        % (args.n_bases // 1_000_000),
            "k": K, "score_mode": "log2", "bases_per_gpu": args.n_bases, "sharding": "one sequence per GPU, "
            "count tables summed across GPUs: %s" % exchange if world > 1 else "single GPU",
            "l2": "inputs (%d MB sequence + 64 MiB count table + 128 MiB score table) exceed the 126 MB L2; "
                  "no explicit flush" % (args.n_bases // 1_000_000)}


# --------------------------------------------------------------------------------------------------
def parity_check(args, ctx, stages, seq_host, rank, world, dist, torch):
    """Outside the timed region: the exact call sequence the timed loop runs, on a prefix of this rank's
    sequence with the spans fetched, against the CPU oracle (checker only).  N = 1: ks_dev_pipeline.
    N > 1: count_async -> reduce_counts -> scores_from_counts_dev -> scan; the reduced count table must
    equal the sum of the per-rank oracle tables.  Returns the parity_check record."""
    from oracle.ksoracle import Oracle
    orc = Oracle()
    P = int(min(args.parity_bases, args.n_bases))
    pre = np.ascontiguousarray(seq_host[:P])
    pre_b = pre.tobytes()
    dev = stages.device
    rec = {"bases_per_rank": P, "ok": False}
    if world == 1:
        ss = ctx.upload([pre])
        counts = torch.zeros(4 ** K, dtype=torch.int32, device=dev)
        scores = torch.zeros(4 ** K, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        want = orc.mode_regions([pre_b], K, MODE_LOG2, MIN_W, MIN_SCORE, thr=THR)
        got = ctx.dev_pipeline(ss, K, MODE_LOG2, MIN_W, MIN_SCORE, thr=THR, d_counts=counts.data_ptr(),
                               d_scores=scores.data_ptr(), fetch_spans=True)
        c_ok = bool((counts.cpu().numpy() == want["counts"]).all()) and got["n"] == want["n"]
        t_ok = scores.cpu().numpy().tobytes() == want["scores"].tobytes()
        rec["call"] = "ks_dev_pipeline"
        # rank mode (the mode the reference codes) through the same call
        want_r = orc.low_comp([pre_b], K, MIN_W, MIN_SCORE, 0.75)
        got_r = ctx.dev_pipeline(ss, K, 0, MIN_W, MIN_SCORE, thr=0.75, d_counts=counts.data_ptr(),
                                 d_scores=scores.data_ptr(), fetch_spans=True)
        r_ok = (scores.cpu().numpy().tobytes() == want_r["ranks"].tobytes()
                and got_r["pos"].tolist() == want_r["pos"].tolist()
                and bool(np.allclose(got_r["score"], want_r["score"], rtol=1e-9, atol=0)))
        rec["rank_mode"] = {"ok": bool(r_ok), "spans": int(len(want_r["pos"]))}
        ss.free()
        del counts, scores
    else:
        stages.load([pre])
        stages.count_async(K)
        stages.reduce_counts(dist)
        stages.scores_from_counts_dev(K, MODE_LOG2, float("nan"))
        gp, gs = stages.scan(K, THR, MIN_W, MIN_SCORE, fetch=True)
        got = {"pos": gp, "score": gs}
        n_loc, c_loc = orc.kmer_counts([pre_b], K)
        tot = torch.from_numpy(c_loc.astype(np.int64)).to(dev)
        n_t = torch.tensor([n_loc], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        dist.all_reduce(n_t)
        c_all = tot.cpu().numpy().astype(np.int32)
        c_ok = bool((stages.counts.cpu().numpy() == c_all).all()) and int(stages.nwords.item()) == int(n_t.item())
        W = orc.scores(c_all, K, float(n_t.item()), MODE_LOG2)
        t_ok = stages.scores.cpu().numpy().tobytes() == W.tobytes()
        want = orc.kmer_regions([pre_b], K, W - THR, MIN_W, MIN_SCORE)
        rec["call"] = "count_async -> reduce_counts (%s) -> scores_from_counts_dev -> scan" % stages.peer_sum_kind()
        r_ok = True
        stages.load([seq_host])
    s_ok = got["pos"].tolist() == want["pos"].tolist() and bool(
        np.allclose(got["score"], want["score"], rtol=1e-9, atol=0))
    ok = c_ok and t_ok and s_ok and r_ok
    if world > 1:
        f = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        ok = bool(f.item())
    rec.update({"ok": bool(ok), "counts_bit_exact": bool(c_ok), "score_table_bit_exact": bool(t_ok),
                "spans_identical": bool(s_ok), "spans": int(len(want["pos"])), "score_rtol": 1e-9,
                "oracle": "oracle/ks_oracle.c (pinned to the compiled reference by tests/test_oracle.py)"})
    return rec


def config3_record(args, stages, rank, world, dist, torch, barrier):
    """BASELINE.json configs[2]: 24 human-like chromosomes (3.083 Gb at scale 1), k = 13, weighted rank, thr 0.75,
    ONE set cut across the N GPUs at chunk boundaries (strong scaling): sharded residency (every rank generates
    and uploads only the chromosomes its window touches), count tables summed across the GPUs, scan carries
    exchanged, spans that cross a cut stitched exactly.  Before the timed passes the same code path is checked
    against the CPU oracle on the set scaled to 1 %."""
    from kmer_spans_b200 import api, synth
    from kmer_spans_b200 import dist as ksd
    dev = stages.device
    element = synth.random_bases(np.random.default_rng(0xE1E), 300)

    def build(scale, want_all=False):
        lens = [int(mb * 1_000_000 * scale) for mb in synth.HUMAN_MB]
        c0, cn, lo, hi = api.plan_shard(lens, world, rank)
        starts = 16 + np.concatenate([[0], np.cumsum(np.array(lens[:-1]) + 1)])
        data = {}
        for i, (s0, ln) in enumerate(zip(starts, lens)):
            if want_all or (s0 < hi and s0 + ln > lo):
                data[i] = synth.genome(ln, 100 + i, element=element)
        return lens, data

    d = dist if world > 1 else None
    K3, THR3 = 13, 0.75
    # ---- parity of this code path on the scaled set (k = 12 keeps the oracle's sort short) ----
    lens_p, data_p = build(0.01, want_all=(rank == 0))
    run = ksd.SplitRun(stages, d, api.SparseSeqs(lens_p, data_p))
    got = run.step(12, 0, THR3, MIN_W, MIN_SCORE, gather_scores=True)
    parts = [(got["pos"], got["score"])]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (got["pos"], got["score"]))
    ok = True
    n_par = 0
    if rank == 0:
        from oracle.ksoracle import Oracle
        want = Oracle().low_comp([data_p[i].tobytes() for i in range(len(lens_p))], 12, MIN_W, MIN_SCORE, THR3)
        pos, score = ksd.merge_spans(parts)
        ok = (bool((stages.counts.cpu().numpy() == want["counts"]).all())
              and stages.scores.cpu().numpy().tobytes() == want["ranks"].tobytes()
              and pos.tolist() == want["pos"].tolist()
              and bool(np.allclose(score, want["score"], rtol=1e-9, atol=0)))
        n_par = len(want["pos"])
    run.free()
    del data_p
    # ---- the timed workload ----
    lens, data = build(args.config3_scale)
    seqs = api.SparseSeqs(lens, data)
    run = ksd.SplitRun(stages, d, seqs)
    for _ in range(2):
        r = run.step(K3, 0, THR3, MIN_W, MIN_SCORE)
    reps = 3
    stages.ctx.set_profile(True)
    stages.ctx.profile(reset=True)
    barrier()
    stages.ctx.timer_start()
    for _ in range(reps):
        r = run.step(K3, 0, THR3, MIN_W, MIN_SCORE)
    ms = stages.ctx.timer_stop() / reps
    barrier()
    prof = stages.ctx.profile(reset=True)
    stages.ctx.set_profile(False)
    # end to end: the window goes up again from (pageable) host memory, spans come back
    barrier()
    t0 = time.perf_counter()
    run.reload()
    r = run.step(K3, 0, THR3, MIN_W, MIN_SCORE)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([ms, e2e_s, float(len(r["pos"])), float(run.hi - run.lo)], dtype=torch.float64, device=dev)
    mx = t.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    bases = run.bases
    run.free()
    nk = 4 ** K3
    peak, _ = measured_peak()
    ms_max, e2e_max = float(mx[0].item()), float(mx[1].item())
    alg = 1.5 * bases / world + 16.0 * nk  # per GPU: its share of the sequence, the whole table
    return {"workload": "BASELINE.json configs[2]: 24 human-like sequences, %.3f Gb, k=13, weighted rank thr 0.75, "
                        "min_width 100, min_score 20; one set cut across %d GPU(s) (strong scaling)"
                        % (bases / 1e9, world),
            "scale": args.config3_scale, "bases": bases, "n_gpus": world, "ms_per_step": ms_max,
            "gbases_per_s": bases / (ms_max * 1e-3) / 1e9,
            "e2e_ms": e2e_max * 1e3, "e2e_gbases_per_s": bases / e2e_max / 1e9,
            "e2e_source": "pageable host memory, every rank uploads only its window",
            "spans": int(t[2].item()), "max_window_bytes_per_gpu": int(mx[3].item()),
            "roofline_frac_per_gpu": alg / (ms_max * 1e-3) / 1e9 / peak,
            "kernels_ms_per_step_rank0": {nm: tot / reps for nm, (tot, n) in prof.items() if n},
            "parity_check": {"ok": bool(ok), "scale": 0.01, "k": 12, "spans": int(n_par),
                             "what": "same calls on the set scaled to 1 %: count table, rank table bit-exact, spans "
                                     "identical to the CPU oracle run on the whole set"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-bases", type=int, default=N_BASES)
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-bases", type=int, default=20_000_000)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the rank-mode / config-1 / config-5 records")
    ap.add_argument("--no-config3", action="store_true", help="skip the configs[2] strong-scaling record")
    ap.add_argument("--config3-scale", type=float, default=1.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from kmer_spans_b200 import api, synth
    from kmer_spans_b200 import dist as ksd
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic input, pinned on the host, resident on the device -------------------------
    seq = synth.config2(args.n_bases, 2 + rank)[0]
    pinned = torch.from_numpy(seq).pin_memory()
    seq_host = pinned.numpy()
    stages = ksd.GpuStages(local_rank)
    ctx = stages.ctx
    stages.load([seq_host])
    stages.alloc_tables(K, dist if world > 1 else None)
    counts_host = torch.empty(4 ** K, dtype=torch.int32).pin_memory()

    def step():
        if world == 1:  # one C-ABI call: ks_dev_pipeline (count -> scores -> scan, everything resident)
            r = ctx.dev_pipeline(stages.ss, K, MODE_LOG2, MIN_W, MIN_SCORE, thr=THR,
                                 d_counts=stages.counts.data_ptr(), d_scores=stages.scores.data_ptr())
            return r["n_spans"]
        # N > 1: count -> all-reduce of the count table (NCCL, issued on the ctx stream) -> scores -> scan;
        # no host synchronisation before the score stage reads its histogram back
        stages.count_async(K)
        stages.reduce_counts(dist)
        stages.scores_from_counts_dev(K, MODE_LOG2, float("nan"))
        nsp, _ = stages.scan(K, THR, MIN_W, MIN_SCORE, fetch=False)
        return nsp

    def barrier():
        torch.cuda.synchronize(dev)
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    parity = None
    if not args.no_parity:
        parity = parity_check(args, ctx, stages, seq_host, rank, world, dist, torch)

    for _ in range(args.warmup):
        n_spans = step()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    for _ in range(200):  # NVML initialisation can take a while on a fresh box: sample the timed region, not the setup
        if sampler.ok or getattr(sampler, "err", None):
            break
        time.sleep(0.01)
    sampler.samples.clear()
    ctx.set_profile(True)
    ctx.profile(reset=True)
    ctx.reset_launches()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        n_spans = step()
    ms = ctx.timer_stop()
    barrier()
    sampler.stop_flag = True  # NVML queries contend with CUDA API calls: sample the device-timed region only
    sampler.join(2)
    launches = ctx.launches()
    prof = ctx.profile(reset=True)
    ctx.set_profile(False)
    levels, revisit_chunks = ctx.scan_stats()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = world * args.n_bases * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer entry point --------------------------------------
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 5))

    def e2e_step():
        if world == 1:
            r = ctx.kmer_mode_regions([seq_host], K, MODE_LOG2, MIN_W, MIN_SCORE, thr=THR, want_tables=False,
                                      counts_out=counts_host.numpy())
            return len(r["pos"]), r["pos"].nbytes + r["score"].nbytes
        r = ksd.run_sharded(stages, dist, [seq_host], [rank], K, MODE_LOG2, MIN_W, MIN_SCORE, thr=THR, gather=False,
                            counts_host=counts_host)
        return len(r["pos"]), r["pos"].nbytes + r["score"].nbytes

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    span_bytes = 0
    for _ in range(e2e_steps):
        _, sb = e2e_step()
        span_bytes = sb
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    e2e_val = world * args.n_bases * e2e_steps / e2e_s / 1e9

    # ---- other workloads on the same context (N = 1; device-resident, CUDA events on the ctx stream) ----
    extra = {}
    if world == 1 and not args.no_extra:
        peak_x, _ = measured_peak()

        def timed(ss, k, mode, thr, reps):
            for _ in range(3):
                ctx.dev_pipeline(ss, k, mode, MIN_W, MIN_SCORE, thr=thr, d_counts=stages.counts.data_ptr(),
                                 d_scores=stages.scores.data_ptr())
            ctx.set_profile(True)
            ctx.profile(reset=True)
            ctx.reset_launches()
            ctx.timer_start()
            for _ in range(reps):
                r = ctx.dev_pipeline(ss, k, mode, MIN_W, MIN_SCORE, thr=thr, d_counts=stages.counts.data_ptr(),
                                     d_scores=stages.scores.data_ptr())
            t = ctx.timer_stop() / reps
            pr = ctx.profile(reset=True)
            ctx.set_profile(False)
            lv, rv = ctx.scan_stats()
            nb = ss.bases
            alg_b = 1.5 * nb + 16.0 * 4 ** k
            return {"ms_per_step": t, "gbases_per_s": nb / (t * 1e-3) / 1e9, "bases": nb, "k": k,
                    "spans": int(r["n_spans"]), "restart_levels": int(lv), "launches_per_step": ctx.launches() / reps,
                    "roofline_frac_pipeline": alg_b / (t * 1e-3) / 1e9 / peak_x,
                    "kernels_ms_per_step": {nm: tot / reps for nm, (tot, n) in pr.items() if n}}

        # the mode the reference codes (kmer_low_comp_regions: weighted rank - thr, thr = 0.75), same sequence
        extra["rank_mode"] = dict(timed(stages.ss, K, 0, 0.75, max(3, min(args.steps, 10))),
                                  workload="same 250 Mb sequence, k=12, weighted rank, thr 0.75")
        # BASELINE.json configs[0]: 1 Mb, k = 8, +-1 mode
        ss1 = ctx.upload([synth.config1()[0]])
        extra["config1"] = dict(timed(ss1, 8, 2, 0.0, 20), workload="configs[0]: 1 Mb, k=8, +-1 (threshold) mode")
        ss1.free()
        # BASELINE.json configs[4], a 10 000-contig subset of the 100 000 (host generation time), k = 10
        from kmer_spans_b200.api import SeqBatch
        sb = SeqBatch.from_list(synth.contigs(10_000, seed=5, k=10))
        ss5 = ctx.upload(sb)
        extra["config5_subset"] = dict(timed(ss5, 10, 1, 0.0, 5),
                                       workload="configs[4] subset: 10 000 contigs of 1-50 kb, k=10, log2 mode")
        extra["config5_subset_rank"] = dict(timed(ss5, 10, 0, 0.75, 5),
                                            workload="configs[4] subset: 10 000 contigs, k=10, rank mode thr 0.75")
        ss5.free()

    # ---- where the multi-GPU step spends its exchange (CUDA events on the ctx stream, outside the timed region) ----
    exchange_ms = None
    if world > 1:
        acc = np.zeros(3)
        reps_x = 5
        for _ in range(reps_x):
            barrier()
            ev = []
            stages.count_async(K)
            stages.reduce_counts(dist, events=ev)
            stages.scores_from_counts_dev(K, MODE_LOG2, float("nan"))
            stages.scan(K, THR, MIN_W, MIN_SCORE, fetch=False)
            torch.cuda.synchronize(dev)
            acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
        tx = torch.tensor(acc / reps_x, dtype=torch.float64, device=dev)
        dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        exchange_ms = {"wait_for_all_counts_barrier": float(tx[0].item()), "table_sum_kernel": float(tx[1].item()),
                       "barrier_after_sum": float(tx[2].item()), "what": "max over ranks, mean of 5 steps"}

    c3 = None
    if not args.no_config3:
        c3 = config3_record(args, stages, rank, world, dist, torch, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event pairs over the timed region) ---------
    peak, peak_src = measured_peak()
    nk = 4 ** K
    alg = {  # algorithmic bytes per launch, DESIGN.md section 5 (1.5 B/base + 16 B/entry overall)
        "count_kernel": 1.0 * args.n_bases + 4.0 * nk,
        "scan_level0": 0.5 * args.n_bases + 8.0 * nk,
    }
    kern = {}
    for name, (tot_ms, n) in prof.items():
        if n:
            kern[name] = {"ms_per_launch": tot_ms / n, "launches_per_step": n / args.steps,
                          "ms_per_step": tot_ms / args.steps, "share_of_step": tot_ms / ms}
    dom = max(("count_kernel", "scan_level0"), key=lambda nm: kern.get(nm, {}).get("ms_per_step", 0.0))
    dom_ms = kern[dom]["ms_per_launch"]
    achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(dom), "algorithmic_bytes_per_launch": alg[dom],
                "ms_per_launch": dom_ms, "peak_source": peak_src,
                "pipeline": {"algorithmic_bytes_per_step": 1.5 * args.n_bases + 16.0 * nk,
                             "achieved": (1.5 * args.n_bases + 16.0 * nk) / (ms / args.steps * 1e-3) / 1e9,
                             "frac": (1.5 * args.n_bases + 16.0 * nk) / (ms / args.steps * 1e-3) / 1e9 / peak},
                "unit_rates": {"count_atomics_per_s": args.n_bases / (kern["count_kernel"]["ms_per_launch"] * 1e-3)
                               if "count_kernel" in kern else None,
                               "scan_gathers_per_s": args.n_bases / (kern["scan_level0"]["ms_per_launch"] * 1e-3)
                               if "scan_level0" in kern else None},
                "kernels": kern}

    out = {
        "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 counts / int64+int128 exact fixed-point scan / f64 tables",
        "data": "synthetic", "config": workload_config(args, world, stages.peer_sum_kind()),
        "clocks": sampler.result(),
        "e2e": {"value": e2e_val, "unit": "Gbases/s", "h2d_bytes_per_step": int(args.n_bases),
                "d2h_bytes_per_step": int(4 * nk + span_bytes), "steps": e2e_steps,
                "call": "ks_kmer_mode_regions (host buffers)" if world == 1 else "dist.run_sharded (host buffers)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "spans_per_step": int(n_spans), "restart_levels": int(levels),
        "parity_check": parity,
        "extra": extra,
        "config3": c3,
        "exchange_ms": exchange_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        sample = min(args.n_bases, 25_000_000)
        dt, nsp, kind = cpu_reference_pass(seq[:sample].tobytes())
        out["cpu_baseline"] = {"value": sample / dt / 1e9, "unit": "Gbases/s", "cores": 1, "kind": kind,
                               "sample": "first %d bases of the same sequence, same k / mode / thresholds, one pass "
                                         "(%.1f s); count + scan = unmodified reference C, log2 table = oracle port"
                                         % (sample, dt)}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the timed call does NOT match the oracle: %s" % json.dumps(parity))
    if c3 is not None and not c3["parity_check"]["ok"]:
        raise SystemExit("bench.py: the configs[2] path does NOT match the oracle")


if __name__ == "__main__":
    main()

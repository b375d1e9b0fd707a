"""GPU tier (-m gpu): the device-resident stage API (ks_seqset_* / ks_dev_*) -- the calls bench.py times --
against the CPU oracle on the same seeded inputs.  Same bars as tests/test_gpu.py."""
import ctypes as C

import numpy as np
import pytest

from kmer_spans_b200 import synth
from tests.test_gpu import assert_spans, ctx  # noqa: F401  (fixture)
from tests.test_oracle import planted, rand_seq

pytestmark = pytest.mark.gpu


def _tables(k):
    import torch
    counts = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
    scores = torch.zeros(4 ** k, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    return counts, scores


def _inputs(seed, big=False):
    rng = np.random.default_rng(seed)
    if big:
        return [synth.genome(6_000_000, seed, n_blocks=(2, 5000)).tobytes(), planted(rng, 50_000)]
    return [planted(rng, 120_000), b"ACGTN" * 7, planted(rng, 9_000), rand_seq(rng, 3000, p_n=0.1), b"ACG"]


@pytest.mark.parametrize("mode,k,thr,mw,ms", [(1, 11, 0.0, 100, 20.0), (0, 12, 0.75, 100, 20.0), (2, 8, 0.0, 50, 10.0),
                                              (0, 6, 0.5, 0, 0.0), (1, 7, 0.0, 14, 1.0), (1, 10, 0.0, 100, 20.0)])
def test_dev_pipeline_matches_oracle(ctx, oracle, mode, k, thr, mw, ms):  # noqa: F811
    """ks_dev_pipeline: the single call `bench.py --gpus 1` times"""
    seqs = _inputs(7000 + k, big=(k >= 10))
    want = oracle.mode_regions(seqs, k, mode, mw, ms, thr=thr)
    ss = ctx.upload(seqs)
    counts, scores = _tables(k)
    for rep in range(2):  # the second pass runs on a packed set and warm caches of the context
        r = ctx.dev_pipeline(ss, k, mode, mw, ms, thr=thr, d_counts=counts.data_ptr(), d_scores=scores.data_ptr(),
                             fetch_spans=True)
        assert r["n"] == want["n"]
        assert (counts.cpu().numpy() == want["counts"]).all()
        assert scores.cpu().numpy().tobytes() == want["scores"].tobytes()
        assert r["n_spans"] == len(want["pos"])
        assert_spans(r, want, exact_scores=(mode == 2), what="dev_pipeline mode %d k %d rep %d" % (mode, k, rep))
    r = ctx.dev_pipeline(ss, k, mode, mw, ms, thr=thr, d_counts=counts.data_ptr(), d_scores=scores.data_ptr())
    assert r["n_spans"] == len(want["pos"])  # fetch_spans=False (what the timed loop does)
    ss.free()


def test_dev_stage_calls_match_oracle(ctx, oracle):  # noqa: F811
    """ks_dev_count -> ks_dev_scores -> ks_dev_scan_counts / ks_dev_scan, and the asynchronous pair
    ks_dev_count_async + ks_dev_scores_devtotal that the multi-GPU composition uses"""
    import torch
    seqs = _inputs(7100)
    ss = ctx.upload(seqs)
    for k, mode, thr, mw, ms in ((8, 1, 0.0, 30, 4.0), (8, 2, 0.0, 100, 10.0), (9, 0, 0.7, 20, 2.0)):
        want = oracle.mode_regions(seqs, k, mode, mw, ms, thr=thr)
        counts, scores = _tables(k)
        n = ctx.dev_count(ss, k, counts.data_ptr())
        assert n == want["n"] and (counts.cpu().numpy() == want["counts"]).all()
        ctx.dev_scores(k, counts.data_ptr(), n, mode, scores.data_ptr())
        ctx.sync()
        assert scores.cpu().numpy().tobytes() == want["scores"].tobytes()
        if mode in (1, 2):
            r = ctx.dev_scan_counts(ss, k, counts.data_ptr(), thr, mw, ms)
            assert_spans(r, want, mode == 2, "dev_scan_counts mode %d" % mode)
        r = ctx.dev_scan(ss, k, scores.data_ptr(), thr, mw, ms)
        assert_spans(r, want, mode == 2, "dev_scan mode %d" % mode)
        # asynchronous variant: nothing synchronised between the count kernel and the score stage
        counts2, scores2 = _tables(k)
        nw = torch.zeros(1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.dev_count_async(ss, k, counts2.data_ptr(), nw.data_ptr())
        tot = ctx.dev_scores_devtotal(k, counts2.data_ptr(), nw.data_ptr(), mode, scores2.data_ptr())
        ctx.sync()
        assert tot == want["n"] and int(nw.item()) == int(want["n"])
        assert (counts2.cpu().numpy() == want["counts"]).all()
        assert scores2.cpu().numpy().tobytes() == want["scores"].tobytes()
        if mode in (1, 2):
            r = ctx.dev_scan_counts(ss, k, counts2.data_ptr(), thr, mw, ms)
            assert_spans(r, want, mode == 2, "async + dev_scan_counts mode %d" % mode)
    ss.free()


def test_seqset_reupload_and_wrap(ctx, oracle):  # noqa: F811
    """ks_seqset_reupload (same device buffers, pack+count behind the copies) and ks_seqset_wrap (caller-owned
    device buffer in the layout of ks_layout.h)"""
    import torch
    k, mode, mw, ms = 7, 1, 20, 3.0
    a, b = _inputs(7200), _inputs(7201)[:3]
    ss = ctx.upload(a)
    counts, scores = _tables(k)
    nw = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for seqs in (b, a, b):
        want = oracle.mode_regions(seqs, k, mode, mw, ms)
        ss.reupload(seqs, k, counts.data_ptr(), nw.data_ptr())
        tot = ctx.dev_scores_devtotal(k, counts.data_ptr(), nw.data_ptr(), mode, scores.data_ptr())
        assert tot == want["n"]
        assert (counts.cpu().numpy() == want["counts"]).all()
        r = ctx.dev_scan_counts(ss, k, counts.data_ptr(), 0.0, mw, ms)
        assert_spans(r, want, False, "reupload")
        assert ss.bases == sum(len(s) for s in seqs)
    ss.free()
    # wrap: the caller lays the buffer out itself
    seqs = a
    lens = np.array([len(s) for s in seqs], np.int64)
    total = 16 + int(lens.sum()) + len(seqs)
    total = (total + 15) // 16 * 16 + 16
    host = np.zeros(total + 64, np.uint8)
    cur = 16
    for s in seqs:
        host[cur:cur + len(s)] = np.frombuffer(s, np.uint8)
        cur += len(s) + 1
    dbuf = torch.from_numpy(host).cuda()
    h = C.c_void_p()
    ctx._ck(ctx.lib.ks_seqset_wrap(ctx.h, C.c_void_p(dbuf.data_ptr()), total + 64,
                                   lens.ctypes.data_as(C.POINTER(C.c_int64)), len(seqs), C.byref(h)))

    class _SS:
        pass
    w = _SS()
    w.h = h
    want = oracle.mode_regions(seqs, k, mode, mw, ms)
    r = ctx.dev_pipeline(w, k, mode, mw, ms, d_counts=counts.data_ptr(), d_scores=scores.data_ptr(), fetch_spans=True)
    assert (counts.cpu().numpy() == want["counts"]).all()
    assert_spans(r, want, False, "wrap")
    ctx.lib.ks_seqset_free(h)


def test_core_table_escape_classes(ctx, oracle, monkeypatch):  # noqa: F811
    """count-derived modes gather one 8-byte core record per two positions; classes beyond the first 255
    distinct counts take the escape through the 2-byte class table.  Small k on a long sequence gives every
    k-mer its own count (hundreds of classes), so most positions escape."""
    rng = np.random.default_rng(7300)
    seqs = [planted(rng, 700_000), rand_seq(rng, 40_000, p_n=0.02)]
    for k in (1, 4, 5, 6):
        for mode, mw, ms in ((1, 20, 1.0), (2, 16, 4.0), (1, 3, 0.5)):
            want = oracle.mode_regions(seqs, k, mode, mw, ms)
            ndistinct = len(np.unique(want["counts"]))
            got = ctx.kmer_mode_regions(seqs, k, mode, mw, ms)
            assert_spans(got, want, mode == 2, "core k %d mode %d (%d classes)" % (k, mode, ndistinct))
            monkeypatch.setenv("KS_NO_CORE_TABLE", "1")
            got = ctx.kmer_mode_regions(seqs, k, mode, mw, ms)
            monkeypatch.delenv("KS_NO_CORE_TABLE")
            assert_spans(got, want, mode == 2, "class table k %d mode %d" % (k, mode))
    assert ndistinct > 255


def test_rank_position_gather_equals_rank_table_scan(ctx, oracle, monkeypatch):  # noqa: F811
    """rank mode gathers the 4-byte position in the rank order and evaluates the linear pieces of the closed
    form; it must agree BIT FOR BIT (scores included) with the scan that gathers the 8-byte rank table, and
    with the oracle.  Both walks (min_width < 15 / >= 15), several k (bucket shift 0 and > 0), deep rescans."""
    import torch
    rng = np.random.default_rng(7400)
    seqs = [planted(rng, 300_000), planted(rng, 20_000), rand_seq(rng, 6000, p_n=0.1), b"ACGT" * 3]
    for k, thr, mw, ms in ((3, 0.6, 20, 2.0), (5, 0.5, 0, 0.0), (6, 0.75, 15, 1.0), (8, 0.75, 100, 5.0), (9, 0.5, 30, 2.0)):
        want = oracle.low_comp(seqs, k, mw, ms, thr)
        got = ctx.kmer_low_comp_regions(seqs, k, mw, ms, thr)
        assert got["w_rank"].tobytes() == want["ranks"].tobytes()
        assert_spans(got, want, False, "rank positions k %d thr %g" % (k, thr))
        monkeypatch.setenv("KS_NO_RANK_POS", "1")
        tab = ctx.kmer_low_comp_regions(seqs, k, mw, ms, thr)
        monkeypatch.delenv("KS_NO_RANK_POS")
        assert got["pos"].tobytes() == tab["pos"].tobytes() and got["score"].tobytes() == tab["score"].tobytes()
    # the stage call
    k, thr, mw, ms = 7, 0.7, 25, 2.0
    want = oracle.low_comp(seqs, k, mw, ms, thr)
    ss = ctx.upload(seqs)
    counts, scores = _tables(k)
    n = ctx.dev_count(ss, k, counts.data_ptr())
    ctx.dev_scores(k, counts.data_ptr(), n, 0, scores.data_ptr())
    r = ctx.dev_scan_ranks(ss, k, thr, mw, ms)
    assert_spans(r, want, False, "dev_scan_ranks")
    r2 = ctx.dev_scan(ss, k, scores.data_ptr(), thr, mw, ms)
    assert r["pos"].tobytes() == r2["pos"].tobytes() and r["score"].tobytes() == r2["score"].tobytes()
    ss.free()
    del torch


@pytest.mark.parametrize("path", ["direct", "smem", "bucket"])
def test_count_paths_match_oracle(ctx, oracle, path, monkeypatch):  # noqa: F811
    """the three counting kernels (direct global reductions, shared-memory table for k <= 7, 1024 buckets through
    shared memory for 8 <= k <= 12, 4096 buckets at k = 13) on ragged inputs, IUPAC bytes, poly-A / tandem arrays (staging rows overflow
    into the direct reduction) and with bucket regions forced to overflow"""
    monkeypatch.setenv("KS_COUNT_PATH", path)
    rng = np.random.default_rng(7500)
    ks = {"direct": (2, 8, 11), "smem": (1, 3, 7), "bucket": (5, 8, 10, 12, 13)}[path]
    for k in ks:
        for trial in range(3):
            seqs = [rand_seq(rng, int(rng.integers(0, 40000)), p_n=float(rng.choice([0, 0.02, 0.3])),
                             alphabet=rng.choice([b"ACGT", b"ACGTacgtRYKMSWBDHVUu-*."]))
                    for _ in range(int(rng.integers(1, 6)))]
            seqs += [b"A" * 70000, b"ACG" * 20000, b"ACGTACGTACGTACGTACGT"[:k], b"", b"N" * 40,
                     planted(rng, 150_000)]
            n1, c1 = oracle.kmer_counts(seqs, k)
            g = ctx.kmer_counts(seqs, k, with_f=False)
            assert g["n"][1] == n1 and (g["counts"] == c1).all(), (path, k, trial)
    if path == "bucket":
        monkeypatch.setenv("KS_BUCKET_CAP", "64")  # almost everything overflows the bucket regions
        seqs = [planted(rng, 300_000), b"AC" * 50000]
        for k in (8, 12, 13):
            n1, c1 = oracle.kmer_counts(seqs, k)
            g = ctx.kmer_counts(seqs, k, with_f=False)
            assert g["n"][1] == n1 and (g["counts"] == c1).all(), ("overflow", k)


def test_count_paths_default_selection_large(ctx, oracle):  # noqa: F811
    """inputs large enough for the default selection to take the bucketed path (>= 2^18 chunks) at k = 8, 10, 12, 13,
    through the resident-set call, the sharded range call and the upload-with-count path"""
    import torch
    seq = synth.genome(6_000_000, 77, n_blocks=(3, 4000)).tobytes()
    seqs = [seq, b"ACGTTGCA" * 1000]
    ss = ctx.upload(seqs)
    for k in (8, 10, 12, 13):
        n1, c1 = oracle.kmer_counts(seqs, k)
        counts, _ = _tables(k)
        assert ctx.dev_count(ss, k, counts.data_ptr()) == n1
        assert (counts.cpu().numpy() == c1).all(), k
        g = ctx.kmer_counts(seqs, k, with_f=False)  # upload + count behind the copies, slab by slab
        assert g["n"][1] == n1 and (g["counts"] == c1).all(), k
        # two ranges of the same set (what a 2-GPU split counts) add up to the whole
        half = ss.chunks // 2
        a, b = _tables(k)[0], _tables(k)[0]
        na = ctx.dev_count_range(ss, k, 0, half, a.data_ptr())
        nb = ctx.dev_count_range(ss, k, half, ss.chunks - half, b.data_ptr())
        assert na + nb == n1 and ((a + b).cpu().numpy() == c1).all(), k
    ss.free()
    del torch


def _virtual_split(seqs, k, mode, thr, mw, ms, world):
    """run_split with `world` virtual ranks on one GPU (one Context and one thread each; carries and tables meet
    at a barrier as they would in an all-gather / all-reduce)"""
    import threading
    import torch
    from kmer_spans_b200 import api
    from kmer_spans_b200 import dist as ksd
    ctxs = [api.Context() for _ in range(world)]
    barrier = threading.Barrier(world)
    blobs, tables, ns, out, errs = [None] * world, [None] * world, [0.0] * world, [None] * world, []

    def worker(r):
        try:
            def all_gather_bytes(b):
                blobs[r] = b
                barrier.wait()
                got = list(blobs)
                barrier.wait()
                return got

            def all_reduce_counts(t, n):
                tables[r], ns[r] = t, n
                barrier.wait()
                if r == 0:
                    tot = torch.stack(tables).sum(0).to(torch.int32)
                    for x in tables:
                        x.copy_(tot)
                    torch.cuda.synchronize()
                barrier.wait()
                return float(sum(ns))

            out[r] = ksd.run_split(ctxs[r], seqs, k, mode, mw, ms, thr, float("nan"), r, world, all_gather_bytes,
                                   all_reduce_counts)
        except threading.BrokenBarrierError:
            pass  # another rank failed first; its error is the one to report
        except Exception as e:  # noqa: BLE001
            import traceback
            errs.append("rank %d: %r\n%s" % (r, e, traceback.format_exc()))
            barrier.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert not errs, "\n".join(errs)
    pos, score = ksd.merge_spans([(o["pos"], o["score"]) for o in out])
    counts = out[0]["counts"].cpu().numpy()
    windows = [o["window_bytes"] for o in out]
    for c in ctxs:
        c.close()
    return dict(pos=pos, score=score, counts=counts, n=out[0]["n"], windows=windows)


def test_config3_scaled_sharded_equals_oracle(ctx, oracle):  # noqa: F811
    """BASELINE.json configs[2] scaled down (24 human-like sequences, weighted rank, thr 0.75) cut across 8 and 5
    virtual GPUs at arbitrary chunk boundaries, every shard holding only its window of the layout: count table,
    and all spans must equal the CPU oracle run on the whole set.  Also log2 mode and a deep-rescan threshold."""
    seqs = [s.tobytes() for s in synth.config3(scale=0.0012)]  # 3.7 Mb in 24 sequences
    total_bytes = sum(len(s) for s in seqs)
    for world, k, mode, thr, mw, ms in ((8, 10, 0, 0.75, 100, 20.0), (5, 8, 0, 0.5, 30, 3.0), (8, 9, 1, 0.0, 100, 20.0)):
        want = oracle.mode_regions(seqs, k, mode, mw, ms, thr=thr)
        got = _virtual_split(seqs, k, mode, thr, mw, ms, world)
        assert got["n"] == want["n"] and (got["counts"] == want["counts"]).all()
        assert len(want["pos"]) > 20
        assert_spans(got, want, False, "config 3 scaled, world %d k %d mode %d" % (world, k, mode))
        # sharded residency: a shard keeps its range plus at most the head of one sequence
        assert max(got["windows"]) < total_bytes / world + max(len(s) for s in seqs) + 4096


@pytest.mark.parametrize("ndev", [2, 3, 8])
def test_multi_device_context_matches_oracle(oracle, ndev):
    """ks_mctx: several shards behind ONE C-ABI call (here all on GPU 0: the device list may repeat an index),
    through the same entry points as the single-GPU path -- sharded upload, per-shard counts summed by the
    peer-memory kernel, scores per device, scan carries folded at the host barrier, spans merged"""
    from kmer_spans_b200 import api
    rng = np.random.default_rng(7600 + ndev)
    m = api.MultiContext([0] * ndev)
    seqs = [planted(rng, 90_000), planted(rng, 7_000), b"ACGTN" * 5, planted(rng, 41_000), b"AC", rand_seq(rng, 9000, p_n=0.1)]
    for k, thr, mw, ms in ((6, 0.6, 10, 3.0), (9, 0.75, 100, 5.0), (4, 0.5, 0, 0.0)):
        want = oracle.low_comp(seqs, k, mw, ms, thr)
        got = m.kmer_low_comp_regions(seqs, k, mw, ms, thr)
        assert (got["n"] == want["n"]).all() and (got["counts"] == want["counts"]).all()
        assert got["w_rank"].tobytes() == want["ranks"].tobytes()
        assert_spans(got, want, False, "mctx rank k %d ndev %d" % (k, ndev))
    for mode, k, mw, ms in ((1, 7, 20, 2.0), (2, 8, 50, 10.0), (3, 5, 15, 1.0)):
        param = 0.6 if mode == 3 else float("nan")
        want = oracle.mode_regions(seqs, k, mode, mw, ms, param=param)
        got = m.kmer_mode_regions(seqs, k, mode, mw, ms, param=param)
        assert got["n"] == want["n"] and (got["counts"] == want["counts"]).all()
        assert got["scores"].tobytes() == want["scores"].tobytes()
        assert_spans(got, want, mode == 2, "mctx mode %d ndev %d" % (mode, ndev))
    n1, c1 = oracle.kmer_counts(seqs, 10)
    g = m.kmer_counts(seqs, 10, with_f=False)
    assert g["n"][1] == n1 and (g["counts"] == c1).all()
    # resident shards, repeated passes (what a benchmark times)
    m.load(seqs)
    want = oracle.mode_regions(seqs, 8, 1, 30, 4.0)
    for rep in range(2):
        r = m.pipeline(8, 1, 30, 4.0, fetch_spans=True)
        assert r["n"] == want["n"]
        assert_spans(r, want, False, "mctx resident rep %d" % rep)
    # errors keep their text and leave the context usable
    with pytest.raises(api.KspansError, match="threshold must be between 0 and 1"):
        m.kmer_low_comp_regions(seqs, 6, 10, 1.0, 1.5)
    got = m.kmer_low_comp_regions(seqs, 6, 10, 3.0, 0.6)
    assert_spans(got, oracle.low_comp(seqs, 6, 10, 3.0, 0.6), False, "after error")
    m.close()


def test_multi_device_edge_cases(oracle):
    """more shards than chunks (empty shards take part in every exchange), k = 1 and 2 (table slices of the sliced
    rank stage that are empty or one entry), nothing counted at all, sequences shorter than k"""
    from kmer_spans_b200 import api
    m = api.MultiContext([0] * 8)
    rng = np.random.default_rng(7700)
    cases = [([b"ACGTACGTAGAGAGAGAGAGAGAGAGCCCT"], 2, 0.5, 2, 0.5),
             ([rand_seq(rng, 90)], 3, 0.5, 0, 0.0),
             ([planted(rng, 5000), b"AC"], 1, 0.5, 3, 0.2),
             ([planted(rng, 3000), planted(rng, 200)], 2, 0.6, 5, 0.5),
             ([b"NNNNNNNNNNNN", b"AC"], 4, 0.75, 1, 0.1)]
    for seqs, k, thr, mw, ms in cases:
        want = oracle.low_comp(seqs, k, mw, ms, thr)
        got = m.kmer_low_comp_regions(seqs, k, mw, ms, thr)
        assert (got["counts"] == want["counts"]).all() and got["n"][0] == want["n"][0]
        # the reference's 0/0 table (nothing counted) holds NaN: compare bit patterns, NaN payloads aside
        assert np.array_equal(np.isnan(got["w_rank"]), np.isnan(want["ranks"]))
        ok = ~np.isnan(want["ranks"])
        assert got["w_rank"][ok].tobytes() == want["ranks"][ok].tobytes()
        assert_spans(got, want, False, "edge k %d" % k)
        c = m.kmer_counts(seqs, k, with_f=False)
        assert (c["counts"] == want["counts"]).all()
    m.close()

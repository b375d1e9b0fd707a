"""GPU tier (-m gpu): the sm_100a path, called through the C ABI, against the CPU oracle
(oracle/ks_oracle.c, pinned to the compiled reference by tests/test_oracle.py) on the same
seeded inputs.

Bars (BASELINE.json north_star): counts, rank tables, +-1-mode scores, span coordinates and
in-scan counts bit-exact; log2 / rank-mode span scores within rtol 1e-9 here (spec: 1e-6) -- the
GPU sums exactly in fixed point, the reference rounds after every addition.
"""
import numpy as np
import pytest

from kmer_spans_b200 import synth
from tests.test_oracle import planted, rand_seq

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    from kmer_spans_b200 import api
    c = api.Context()
    yield c
    c.close()


def assert_spans(got, want, exact_scores, what=""):
    gp, wp = got["pos"], want["pos"]
    if gp.tolist() != wp.tolist():
        n = min(len(gp), len(wp))
        bad = next((i for i in range(n) if gp[i].tolist() != wp[i].tolist()), n)
        raise AssertionError("%s: spans differ: got %d want %d; first difference at %d: got %s want %s" % (
            what, len(gp), len(wp), bad, gp[bad:bad + 3].tolist(), wp[bad:bad + 3].tolist()))
    if exact_scores:
        assert got["score"].tobytes() == want["score"].tobytes(), what
    else:
        np.testing.assert_allclose(got["score"], want["score"], rtol=RTOL, atol=0, err_msg=what)


# ---- counting ---------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 11, 13, 15])
def test_counts_small(ctx, oracle, k):
    rng = np.random.default_rng(1000 + k)
    kk = min(k, 11)
    for trial in range(8):
        seqs = [rand_seq(rng, int(rng.integers(0, 3000)), p_n=rng.choice([0, 0.05, 0.3]),
                         alphabet=rng.choice([b"ACGT", b"ACGTacgtRYKMSWBDHVUu-*."]))
                for _ in range(int(rng.integers(1, 6)))]
        seqs += [b"ACGTACGTACGTACGTACGT"[:k], b"NN" + b"ACGTACGTACGTACGTACGT"[:k],
                 b"ACGTACGTACGTACGTACGT"[:k] + b"N", b"ACGTACGTACGTACGTACGT"[:k + 1], b"", b"N" * 40]
        if k > 11:
            seqs = seqs[:3]
        n1, c1 = oracle.kmer_counts(seqs, k)
        g = ctx.kmer_counts(seqs, k, with_f=False)
        assert g["n"][1] == n1
        assert (g["counts"] == c1).all()
    del kk


def test_counts_known_answers(ctx):
    """the reference's own pins: test.R:365-375 (KA1) and test.R:66-77 (KA2)"""
    from kmer_spans_b200 import api
    g = ctx.kmer_counts(b"CGCCAATGCG", 2)
    got = {a: int(b) for a, b in zip(api.kmer_seq(2), g["counts"]) if b}
    assert got == {"CG": 2, "GC": 2, "CC": 1, "CA": 1, "AA": 1, "AT": 1, "TG": 1}
    rng = np.random.default_rng(5)
    s = rand_seq(rng, 100000)
    c1 = ctx.kmer_counts(s, 2)["counts"]
    c2 = ctx.kmer_counts(s + b"N" * 36 + s, 2)["counts"]
    assert (c2 == 2 * c1).all()


def test_counts_hot_kmers_and_many_contigs(ctx, oracle):
    """poly-A / tandem arrays hammer single table entries; 3000 short contigs exercise packing"""
    seqs = [b"A" * 200000, b"AC" * 100000, b"ACGNNACG" * 1000]
    seqs += [s.tobytes() for s in synth.contigs(3000, seed=11, lo=20, hi=3000, k=10)]
    for k in (4, 10):
        n1, c1 = oracle.kmer_counts(seqs, k)
        g = ctx.kmer_counts(seqs, k, with_f=False)
        assert g["n"][1] == n1
        assert (g["counts"] == c1).all()


# ---- score tables -----------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 4, 6, 8, 10])
def test_ranks_bitexact(ctx, oracle, k):
    rng = np.random.default_rng(1100 + k)
    for trial in range(4):
        s = planted(rng, int(rng.integers(4 ** min(k, 6), 60 * 4 ** min(k, 6))))
        n, c = oracle.kmer_counts(s, k)
        if n == 0:
            continue
        got = ctx.kmer_scores(c, k, n, 0)
        assert got.tobytes() == oracle.rank(c, k, n).tobytes()


def test_ranks_adversarial_tables(ctx, oracle):
    rng = np.random.default_rng(9)
    k = 9
    n = 4 ** k
    for trial in range(16):
        kind = trial % 8
        if kind == 0:
            c = np.full(n, int(rng.integers(1, 5)), np.int32)
        elif kind == 1:
            c = rng.poisson(rng.uniform(0.2, 30), n).astype(np.int32)
        elif kind == 2:
            c = (rng.pareto(1.2, n) * 3).astype(np.int32)
        elif kind == 3:
            c = np.zeros(n, np.int32); c[rng.integers(0, n, 50)] = rng.integers(1, 2 ** 20, 50)
        elif kind == 4:
            c = (2 ** rng.integers(0, 12, n)).astype(np.int32)
        elif kind == 5:
            c = rng.integers(0, 3, n).astype(np.int32)
        elif kind == 6:
            c = np.ones(n, np.int32); c[: n // 2] = 3
        else:
            c = rng.integers(0, 2 ** 31 - 1, n).astype(np.int32) // int(rng.integers(1, 2 ** 20))
        total = float(c.astype(np.int64).sum())
        if kind == 4:
            total = float(2 ** int(rng.integers(20, 40)))
        if total == 0:
            continue
        got = ctx.kmer_scores(c, k, total, 0)
        assert got.tobytes() == oracle.rank(c, k, total).tobytes(), (trial, kind)
    zero = np.zeros(n, np.int32)
    got = ctx.kmer_scores(zero, k, 0.0, 0)
    assert got[0] == 0 and np.isnan(got[1:]).all()


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_mode_tables_bitexact(ctx, oracle, mode):
    rng = np.random.default_rng(1200 + mode)
    for k in (3, 6, 9):
        s = planted(rng, 40 * 4 ** min(k, 6))
        n, c = oracle.kmer_counts(s, k)
        param = 0.6 if mode == 3 else float("nan")
        got = ctx.kmer_scores(c, k, n, mode, param)
        want = oracle.scores(c, k, n, mode, param)
        assert got.tobytes() == want.tobytes()


# ---- scan + spans -----------------------------------------------------------------------------
CASES = [(2, 0.5, 20, 10), (3, 0.75, 20, 10), (4, 0.75, 5, 2), (5, 0.5, 0, 0), (6, 0.6, 10, 3),
         (8, 0.75, 100, 20), (4, 0.5, -1, 0), (3, 0.9, 0, 0.5), (7, 0.5, 3, 1), (10, 0.75, 30, 5)]


@pytest.mark.parametrize("k,thr,mw,ms", CASES)
def test_low_comp_regions(ctx, oracle, k, thr, mw, ms):
    rng = np.random.default_rng(1300 + k)
    for trial in range(6):
        seqs = [planted(rng, int(rng.integers(50, 30000))) for _ in range(int(rng.integers(1, 5)))]
        if trial % 3 == 0:
            seqs.insert(1, b"ACG"[: k - 1])  # shorter than k: skipped, seq_id still advances
        o = oracle.low_comp(seqs, k, mw, ms, thr)
        g = ctx.kmer_low_comp_regions(seqs, k, mw, ms, thr)
        assert (g["n"] == o["n"]).all()
        assert (g["counts"] == o["counts"]).all()
        assert g["w_rank"].tobytes() == o["ranks"].tobytes()
        assert_spans(g, o, exact_scores=False, what="k=%d thr=%g trial=%d" % (k, thr, trial))


@pytest.mark.parametrize("k", [2, 5, 8])
def test_kmer_regions_user_weights(ctx, oracle, k):
    rng = np.random.default_rng(1400 + k)
    for trial in range(12):
        seqs = [planted(rng, int(rng.integers(50, 20000))) for _ in range(int(rng.integers(1, 4)))]
        kind = trial % 4
        if kind == 0:
            W = rng.choice([-1.0, 1.0], 4 ** k, p=[0.7, 0.3])
        elif kind == 1:
            W = rng.normal(-0.3, 1.0, 4 ** k)
        elif kind == 2:
            n, c = oracle.kmer_counts(seqs, k)
            W = oracle.scores(c, k, n, 2)
        else:
            W = rng.normal(-0.2, 1.0, 4 ** k)
            W[rng.integers(0, 4 ** k, 3)] = np.nan
            W[rng.integers(0, 4 ** k, 2)] = -np.inf
        mw, ms = [(0, 0), (10, 3), (30, 8)][trial % 3]
        o = oracle.kmer_regions(seqs, k, W, mw, ms)
        g = ctx.kmer_regions(seqs, k, W, mw, ms)
        assert g["n"] == o["n"]
        assert_spans(g, o, exact_scores=(kind in (0, 2)), what="k=%d trial=%d" % (k, trial))
        assert (g["counts"] == o["counts"]).all()  # in-scan counts incl. rescans


def test_gpu_equals_host_emulation_bitexact(ctx, oracle):
    """exact arithmetic: the kernels and the left-to-right host fold of the same chunk functions
    (tests/emu) must agree bit for bit, scores included"""
    from tests.emu.emu import Emu
    emu = Emu()
    rng = np.random.default_rng(77)
    for k, thr in ((4, 0.5), (7, 0.75), (9, 0.6)):
        seqs = [planted(rng, 40000), planted(rng, 5000)]
        o = oracle.low_comp(seqs, k, 5, 1, thr)
        e = emu.scan(seqs, k, o["ranks"], thr, 5, 1)
        g = ctx.kmer_low_comp_regions(seqs, k, 5, 1, thr)
        assert g["pos"].tolist() == e["pos"].tolist()
        assert g["score"].tobytes() == e["score"].tobytes()
        lv, rv = ctx.scan_stats()
        assert lv == e["levels"]


def test_config1_all_modes(ctx, oracle):
    """BASELINE.json configs[0]: 1 Mb, k=8.  +-1 around the median gives run-sized excursions
    (hundreds of tiles per look-back) and deep restart nesting."""
    seq = synth.config1()[0].tobytes()
    n, c = oracle.kmer_counts(seq, 8)
    W = oracle.scores(c, 8, n, 2)
    o = oracle.kmer_regions([seq], 8, W, 100, 20)
    g = ctx.kmer_regions([seq], 8, W, 100, 20)
    assert_spans(g, o, exact_scores=True, what="sign/user")
    assert (g["counts"] == o["counts"]).all()
    g = ctx.kmer_mode_regions([seq], 8, 2, 100, 20)
    assert g["scores"].tobytes() == W.tobytes()
    assert_spans(g, o, exact_scores=True, what="sign/fused")
    o = oracle.low_comp([seq], 8, 100, 20, 0.75)
    g = ctx.kmer_low_comp_regions([seq], 8, 100, 20, 0.75)
    assert len(o["pos"]) > 20
    assert g["w_rank"].tobytes() == o["ranks"].tobytes()
    assert_spans(g, o, exact_scores=False, what="rank")
    o = oracle.mode_regions([seq], 8, 1, 100, 20)
    g = ctx.kmer_mode_regions([seq], 8, 1, 100, 20)
    assert g["scores"].tobytes() == o["scores"].tobytes()
    assert_spans(g, o, exact_scores=False, what="log2")
    for thr, mw, ms in ((0.5, 0, 0), (0.5, 100, 20), (0.6, 10, 1)):  # deep rescans
        o = oracle.low_comp([seq], 8, mw, ms, thr)
        g = ctx.kmer_low_comp_regions([seq], 8, mw, ms, thr)
        assert_spans(g, o, exact_scores=False, what="rank thr=%g" % thr)


def test_contigs_all_modes(ctx, oracle):
    """BASELINE.json configs[4] (subset): 1500 short contigs, k=10, all three modes"""
    seqs = [s.tobytes() for s in synth.contigs(1500, seed=5, k=10)]
    for mode, thr in ((0, 0.75), (1, 0.0), (2, 0.0)):
        for mw in (0, 100):
            o = oracle.mode_regions(seqs, 10, mode, mw, 20 if mw else 5, thr=thr)
            g = ctx.kmer_mode_regions(seqs, 10, mode, mw, 20 if mw else 5, thr=thr)
            assert g["n"] == o["n"]
            assert (g["counts"] == o["counts"]).all()
            assert g["scores"].tobytes() == o["scores"].tobytes()
            assert_spans(g, o, exact_scores=(mode == 2), what="mode=%d mw=%d" % (mode, mw))


def test_k12_20mb_rank_and_log2(ctx, oracle):
    """BASELINE.json configs[1] scaled to 20 Mb so the oracle finishes in seconds"""
    seq = synth.genome(20_000_000, 2, n_blocks=(5, 50_000)).tobytes()
    o = oracle.low_comp([seq], 12, 100, 20, 0.75)
    g = ctx.kmer_low_comp_regions([seq], 12, 100, 20, 0.75)
    assert (g["counts"] == o["counts"]).all()
    assert g["w_rank"].tobytes() == o["ranks"].tobytes()
    assert len(o["pos"]) > 500
    assert_spans(g, o, exact_scores=False, what="rank")
    o = oracle.mode_regions([seq], 12, 1, 100, 20)
    g = ctx.kmer_mode_regions([seq], 12, 1, 100, 20)
    assert g["scores"].tobytes() == o["scores"].tobytes()
    assert_spans(g, o, exact_scores=False, what="log2")


def test_full_size_properties(ctx):
    """BASELINE.json configs[1] at full size (250 Mb, k=12) through size-independent properties"""
    seq = synth.config2()[0]
    ss = ctx.upload([seq])
    assert ss.bases == seq.size
    del ss
    r = ctx.kmer_low_comp_regions([seq], 12, 100, 20, 0.75)
    counts = r["counts"]
    assert counts.astype(np.int64).sum() == r["n"][0]                    # checksum of the table
    assert r["n"][0] == seq.size - 5 * 50_000 - 6 * 11                    # 6 runs lose k-1 words each
    ranks = r["w_rank"]
    order = np.lexsort((np.arange(ranks.size), counts))
    assert (np.diff(ranks[order]) >= 0).all()                             # ranks non-decreasing in count order
    assert ranks[order[0]] == 0
    top = ranks[order[-1]] + counts[order[-1]] / r["n"][0]
    assert abs(top - 1.0) < 1e-9                                           # cumulative mass reaches 1
    pos, score = r["pos"], r["score"]
    assert len(pos) > 5000
    assert (np.diff(pos[:, 1]) > 0).all()                                  # sorted by start, disjoint starts
    assert (pos[:, 2] - pos[:, 1] >= 100).all() and (score[:, 0] >= 20).all() and (score[:, 1] == 0).all()
    r2 = ctx.kmer_low_comp_regions([seq], 12, 100, 20, 0.75, want_tables=False)
    assert r2["pos"].tobytes() == pos.tobytes() and r2["score"].tobytes() == score.tobytes()  # deterministic
    # every planted tandem array (2 kb every 100 kb) is covered by a span
    starts = np.arange(100_000 // 3, seq.size - 2000, 100_000)
    idx = np.searchsorted(pos[:, 1], starts + 1500) - 1
    covered = (pos[idx, 1] <= starts + 1500) & (pos[idx, 2] >= starts + 500)
    in_n = np.array([seq[s + 1000] == ord("N") for s in starts])
    assert covered[~in_n].mean() > 0.95


# ---- error behaviour ----------------------------------------------------------------------------
def test_errors(ctx):
    from kmer_spans_b200.api import KspansError
    with pytest.raises(KspansError, match="threshold must be between 0 and 1"):
        ctx.kmer_low_comp_regions([b"ACGTACGT"], 2, 1, 1.0, 1.0)
    with pytest.raises(KspansError, match="positive integer"):
        ctx.kmer_counts([b"ACGT"], 0)
    with pytest.raises(KspansError, match="4\\^k scores"):
        ctx.kmer_regions([b"ACGT"], 2, np.zeros(15), 1, 1.0)
    W = np.zeros(16)
    W[3] = np.inf
    with pytest.raises(KspansError) as ei:
        ctx.kmer_regions([b"ACGTACGTAGAGAGAG"], 2, W, 1, 1.0)
    assert ei.value.code == 3
    r = ctx.kmer_regions([b"ACGTACGTAGAGAGAG"], 2, np.zeros(16), 1, 1.0)  # usable after an error
    assert len(r["pos"]) == 0 and r["pos"].shape == (0, 3) and r["score"].shape == (0, 2)


# ---- one sequence set split across GPUs: exact stitching -----------------------------------------
@pytest.mark.parametrize("world", [2, 3, 5])
def test_split_scan_stitches_exactly(ctx, oracle, world):
    """Virtual ranks on one GPU (one Context and one thread each; the carries meet at a barrier as they
    would in an all-gather): counts all-reduced, level-0 scan cut at arbitrary chunk boundaries, spans
    that cross the cuts -- including run-sized excursions in +-1 mode -- must equal the unsharded result."""
    import threading
    import torch
    from kmer_spans_b200 import api
    from kmer_spans_b200 import dist as ksd
    rng = np.random.default_rng(4000 + world)
    seqs = [planted(rng, 60000), planted(rng, 9000), b"ACGTN" * 5, planted(rng, 33000)]
    for k, mode, thr, mw, ms in ((6, 0, 0.6, 10, 3), (8, 2, 0.0, 50, 10), (5, 1, 0.0, 0, 0)):
        want = oracle.mode_regions(seqs, k, mode, mw, ms, thr=thr)
        ctxs = [api.Context() for _ in range(world)]
        barrier = threading.Barrier(world)
        blobs = [None] * world
        tables = [None] * world
        ns = [0.0] * world
        out = [None] * world
        errs = []

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

                out[r] = ksd.run_split(ctxs[r], seqs, k, mode, mw, ms, thr, float("nan"), r, world,
                                       all_gather_bytes, all_reduce_counts)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
                barrier.abort()

        th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(120)
        assert not errs, errs
        assert (out[0]["counts"].cpu().numpy() == want["counts"]).all()
        pos, score = ksd.merge_spans([(o["pos"], o["score"]) for o in out])
        got = dict(pos=pos, score=score)
        assert_spans(got, want, exact_scores=(mode == 2), what="world=%d k=%d mode=%d" % (world, k, mode))
        for c in ctxs:
            c.close()


def test_k13_multipass_count_and_scan(ctx, oracle):
    """k=13: the 256 MiB count table exceeds L2, counting runs in 4 passes over table slices"""
    seq = synth.genome(6_000_000, 13, n_blocks=(3, 5_000)).tobytes()
    seqs = [seq, b"ACGTTTGACCANNNACGT" * 50]
    n1, c1 = oracle.kmer_counts(seqs, 13)
    g = ctx.kmer_counts(seqs, 13, with_f=False)
    assert g["n"][1] == n1
    assert (g["counts"] == c1).all()
    o = oracle.low_comp(seqs, 13, 100, 20, 0.75)
    r = ctx.kmer_low_comp_regions(seqs, 13, 100, 20, 0.75)
    assert (r["counts"] == o["counts"]).all()
    assert r["w_rank"].tobytes() == o["ranks"].tobytes()
    assert_spans(r, o, exact_scores=False, what="k13 rank")


def test_kmers_to_file_roundtrip(ctx, oracle, tmp_path):
    from kmer_spans_b200 import api
    rng = np.random.default_rng(31)
    seqs = [planted(rng, 5000), planted(rng, 300), planted(rng, 8000)]
    info = api.kmers_to_file(ctx, seqs, str(tmp_path) + "/", [3, 6], min_l=1000)
    assert info["seq_fl"] == 2 and info["seq_size"] == 13300
    r = api.read_kmers(info["out"])
    assert r["k"] == [3, 6]
    for k, c in zip(r["k"], r["counts"]):
        assert (c == oracle.kmer_counts([seqs[0], seqs[2]], k)[1]).all()


def test_config2_full_size_parity(ctx, oracle):
    """BASELINE.json configs[1] at FULL size (250 Mb, k=12) against the oracle: rank mode (thousands of
    spans, two restart levels) and log2 mode (run-sized excursions).  ~40 s of CPU oracle time."""
    seq = synth.config2()[0]
    sb = seq.tobytes()
    o = oracle.low_comp([sb], 12, 100, 20, 0.75)
    g = ctx.kmer_low_comp_regions([seq], 12, 100, 20, 0.75)
    assert (g["n"] == o["n"]).all()
    assert (g["counts"] == o["counts"]).all()
    assert g["w_rank"].tobytes() == o["ranks"].tobytes()
    assert len(o["pos"]) > 10000
    assert_spans(g, o, exact_scores=False, what="config 2 rank")
    o = oracle.mode_regions([sb], 12, 1, 100, 20)
    g = ctx.kmer_mode_regions([seq], 12, 1, 100, 20)
    assert g["scores"].tobytes() == o["scores"].tobytes()
    assert_spans(g, o, exact_scores=False, what="config 2 log2")


def test_seqbatch_input(ctx, oracle):
    """many contigs handed over as one buffer + lengths (api.SeqBatch)"""
    from kmer_spans_b200 import api
    seqs = [s.tobytes() for s in synth.contigs(300, seed=9, lo=30, hi=4000, k=7)]
    o = oracle.low_comp(seqs, 7, 20, 4, 0.7)
    g = ctx.kmer_low_comp_regions(api.SeqBatch.from_list(seqs), 7, 20, 4, 0.7)
    assert (g["counts"] == o["counts"]).all()
    assert_spans(g, o, exact_scores=False, what="seqbatch")


# ---- windowed occurrence histograms --------------------------------------------
def _window_case(ctx, oracle, seqs, kms, k, window, what):
    want = oracle.window_dist(seqs, kms, k, window, True)
    got = ctx.window_kmer_dist(seqs, kms, window, freq=False, ret_flag=1)
    assert np.array_equal(got["dist"].T, want["dist"]), what
    assert np.array_equal(got["seq_i"], want["included"]), what
    for g, w in zip(got["scores"], want["pos"]):
        assert (g is None) == (w is None), what
        if g is not None:
            assert np.array_equal(g.T, w), what
    got2 = ctx.window_kmer_dist(seqs, kms, window, freq=False)
    assert got2["scores"] is None and np.array_equal(got2["dist"].T, want["dist"]), what


def test_window_dist_small(ctx, oracle):
    rng = np.random.default_rng(4100)
    for t in range(40):
        k = int(rng.integers(1, 9))
        window = int(rng.integers(2 * k, 200))
        seqs = [rand_seq(rng, int(rng.integers(0, 3000)), p_n=float(rng.choice([0, 0.05, 0.3])))
                for _ in range(int(rng.integers(1, 6)))]
        seqs += [rand_seq(rng, window), rand_seq(rng, window + 1), b"", rand_seq(rng, window - 1)]
        kms = [oracle.kmer_seq(k, int(c)).encode() for c in rng.integers(0, 4 ** min(k, 3), int(rng.integers(1, 7)))]
        kms = [x[-k:] if len(x) >= k else (b"A" * (k - len(x)) + x) for x in kms]
        if t % 4 == 0:
            kms += [b"N" * k, kms[0]]
        _window_case(ctx, oracle, seqs, kms, k, window, "t=%d k=%d window=%d" % (t, k, window))


def test_window_dist_large_window_and_tandem(ctx, oracle):
    rng = np.random.default_rng(4200)
    s = np.frombuffer(rand_seq(rng, 120_000), np.uint8).copy()
    s[30_000:36_000] = np.tile(np.frombuffer(b"AC", np.uint8), 3000)  # dense occurrences
    s[80_000:80_020] = ord("N")
    seqs = [s.tobytes(), rand_seq(rng, 50_001), rand_seq(rng, 50_000)]
    # 50 001 bins x 1 k-mer does not fit shared memory -> global histogram path
    _window_case(ctx, oracle, seqs, [b"AC"], 2, 50_000, "global bins")
    _window_case(ctx, oracle, seqs, [b"AC", b"CA", b"GT", b"AA"], 2, 1000, "shared bins")


def test_window_dist_device_stage_excludes_exact_length(ctx, oracle):
    import torch
    rng = np.random.default_rng(4300)
    window, k = 64, 2
    seqs = [rand_seq(rng, 5000, p_n=0.05), rand_seq(rng, window), rand_seq(rng, window), rand_seq(rng, 300)]
    kms = [b"AC", b"TT", b"GA"]
    want = oracle.window_dist(seqs, kms, k, window, True)
    ss = ctx.upload(seqs)
    codes = np.array([oracle.kmer_code(x, k) for x in kms], np.uint32)
    d_dist = torch.empty((len(kms), window + 1), dtype=torch.int32, device="cuda")
    d_pos = torch.empty((len(kms), ss.positions), dtype=torch.int32, device="cuda")
    ctx.dev_window_dist(ss, k, codes, window, d_dist.data_ptr(), d_pos.data_ptr())
    ctx.sync()
    assert np.array_equal(d_dist.cpu().numpy(), want["dist"])
    pos = d_pos.cpu().numpy()
    for q, w in enumerate(want["pos"]):
        st = ss.start(q)
        got = pos[:, st:st + len(seqs[q])]
        assert np.array_equal(got, w if w is not None else np.zeros_like(got)), q
    ss.free()


def test_window_dist_errors(ctx):
    from kmer_spans_b200._lib import KspansError
    with pytest.raises(KspansError, match="two times k"):
        ctx.window_kmer_dist([b"ACGTACGTACGT"], [b"ACG"], 5)
    with pytest.raises(ValueError, match="same size"):
        ctx.window_kmer_dist([b"ACGTACGTACGT"], [b"ACG", b"AC"], 8)


# ---- transition-score scan (tr_lr_regions_r) ------------------------------------
def _tr_case(ctx, oracle, seqs, k, init, trans, min_len, exact, what):
    n = 4 ** k
    kms = [oracle.kmer_seq(k, c).encode() for c in range(n)]
    want = oracle.tr_lr_regions(seqs, k, init, trans, min_len)
    got = ctx.lr_regions(seqs, (k, min_len), kms, init, trans)
    assert np.array_equal(got["kmer_scores"][:, 0], init) and np.array_equal(got["kmer_scores"][:, 1], trans)
    assert_spans(got, want, exact, what)
    return len(want["pos"])


def test_tr_lr_small(ctx, oracle):
    rng = np.random.default_rng(5100)
    total = 0
    for t in range(60):
        k = int(rng.integers(1, 7))
        n = 4 ** k
        seqs = [rand_seq(rng, int(rng.integers(0, 6000)), p_n=float(rng.choice([0, 0.05, 0.3])))
                for _ in range(int(rng.integers(1, 5)))]
        if t % 3 == 0:  # run-end rules (:340-341)
            seqs += [rand_seq(rng, k), rand_seq(rng, k + 1), rand_seq(rng, k + 2), rand_seq(rng, k) + b"N",
                     rand_seq(rng, k) + b"NA", rand_seq(rng, k) + b"N" + rand_seq(rng, k + 3),
                     rand_seq(rng, 40) + b"NN" + rand_seq(rng, k + 1), b"", b"NNNN"]
        exact = t % 2 == 0
        if exact:
            init, trans = rng.integers(-2, 3, n).astype(float), rng.integers(-2, 3, n).astype(float)
        else:
            init, trans = rng.normal(0, 1, n), rng.normal(-0.1, 1, n)
        if t % 5 == 0:
            trans[rng.integers(0, n)] = -np.inf
            init[rng.integers(0, n)] = -np.inf
        min_len = int(rng.choice([0, 1, 3, 10, 40]))
        total += _tr_case(ctx, oracle, seqs, k, init, trans, min_len, exact, "t=%d k=%d min_len=%d" % (t, k, min_len))
    assert total > 2000


def test_tr_lr_planted_multi_tile(ctx, oracle):
    """long excursions across many tiles, deep re-scans (min_length 0) and wide regions"""
    rng = np.random.default_rng(5200)
    k = 4
    n = 4 ** k
    seqs = [planted(rng, 300_000), planted(rng, 70_000)]
    # log-ratio style scores: k-mers of the planted repeats score up
    _, counts = oracle.kmer_counts(seqs, k)
    f = (counts + 1.0) / (counts.sum() + n)
    trans = np.log2(f * n) - 0.05
    init = np.log2(f * n)
    for min_len in (0, 25, 200):
        got = _tr_case(ctx, oracle, seqs, k, init, trans, min_len, False, "planted min_len=%d" % min_len)
        assert got > 0
    levels, _ = ctx.scan_stats()
    assert levels >= 2
    # +-1 scores: exact arithmetic, long positive stretches
    trans = np.where(counts > np.median(counts), 1.0, -1.0)
    init = np.where(counts > np.median(counts), 2.0, -2.0)
    _tr_case(ctx, oracle, seqs, k, init, trans, 0, True, "sign scores")
    _tr_case(ctx, oracle, seqs, k, init, trans, 30, True, "sign scores min_len 30")


def test_tr_lr_rejects_nan(ctx):
    from kmer_spans_b200._lib import KspansError
    w = np.zeros(16)
    w[3] = np.nan
    kms = [a + b for a in "ACTG" for b in "ACTG"]
    with pytest.raises(KspansError, match="NaN"):
        ctx.lr_regions([b"ACGTACGTACGTAAAA"], (2, 0), kms, np.zeros(16), w)
    with pytest.raises(KspansError, match="4\\^k long"):
        ctx.lr_regions([b"ACGT"], (2, 0), kms[:5], np.zeros(16), w)


# ---- both walk implementations and both gather tables against the oracle ---------------------------
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fast_and_general_paths_agree_with_oracle(ctx, oracle, mode, monkeypatch):
    """min_width >= 15 takes the summary-based walk (scan_walk_fast_kernel + scan_detail_kernel), smaller
    widths the position-by-position walk; count-derived modes gather 8-byte core records (two positions per
    gather), 2-byte classes or 4-byte counts, depending on what is disabled.
    All combinations must give the oracle's spans."""
    rng = np.random.default_rng(6100 + mode)
    seqs = [planted(rng, 400_000), planted(rng, 30_000), rand_seq(rng, 5000, p_n=0.2)]
    k = 6
    thr = 0.6 if mode == 0 else 0.0
    for mw, ms in ((14, 2.0), (15, 2.0), (16, 0.0), (31, 1.0), (100, 5.0), (0, 8.0)):
        want = oracle.mode_regions(seqs, k, mode, mw, ms, thr=thr)
        for env in ({}, {"KS_NO_FAST_WALK": "1"}, {"KS_NO_CLASS_TABLE": "1"},
                    {"KS_NO_FAST_WALK": "1", "KS_NO_CLASS_TABLE": "1"}, {"KS_NO_CORE_TABLE": "1"},
                    {"KS_NO_FAST_WALK": "1", "KS_NO_CORE_TABLE": "1"},
                    # units of two chunks (min_width >= 31) off; the persistent core-record kernel off; rank
                    # positions through the 32-byte records of the (k-1)-mers (default only for k >= 13)
                    {"KS_NO_PAIR": "1"}, {"KS_NO_CORE_PIPE": "1"}, {"KS_RANK_CORE_MIN_K": "2"},
                    {"KS_RANK_CORE_MIN_K": "2", "KS_NO_PAIR": "1"},
                    {"KS_RANK_CORE_MIN_K": "2", "KS_NO_FAST_WALK": "1"}):
            for name in ("KS_NO_FAST_WALK", "KS_NO_CLASS_TABLE", "KS_NO_CORE_TABLE", "KS_NO_PAIR", "KS_NO_CORE_PIPE",
                         "KS_RANK_CORE_MIN_K"):
                monkeypatch.delenv(name, raising=False)
            for name, val in env.items():
                monkeypatch.setenv(name, val)
            got = ctx.kmer_mode_regions(seqs, k, mode, mw, ms, thr=thr)
            assert_spans(got, want, mode == 2, "mode %d mw %d env %s" % (mode, mw, sorted(env)))
    assert len(want["pos"]) >= 0


def test_detail_list_overflow_falls_back_to_general_walk(ctx, oracle, monkeypatch):
    """a list of undecided chunks that does not fit sends the level to the position-by-position walk"""
    rng = np.random.default_rng(6200)
    seqs = [planted(rng, 300_000)]
    want = oracle.mode_regions(seqs, 5, 0, 15, 1.0, thr=0.55)
    assert len(want["pos"]) > 20
    monkeypatch.setenv("KS_DETAIL_CAP", "3")
    got = ctx.kmer_mode_regions(seqs, 5, 0, 15, 1.0, thr=0.55)
    monkeypatch.delenv("KS_DETAIL_CAP")
    assert_spans(got, want, False, "overflow fallback")
    got = ctx.kmer_mode_regions(seqs, 5, 0, 15, 1.0, thr=0.55)
    assert_spans(got, want, False, "fast path")

"""CPU tier: the level-wise, exact-fixed-point restatement the kernels implement (driven through
tests/emu, which folds the SAME per-chunk functions the kernels use) against the oracle."""
import numpy as np
import pytest

from kmer_spans_b200 import synth
from tests.emu.emu import Emu
from tests.test_oracle import planted, rand_seq


@pytest.fixture(scope="module")
def emu():
    return Emu()


def assert_spans(a, b, exact_scores):
    assert a["pos"].tolist() == b["pos"].tolist()
    if exact_scores:
        assert a["score"].tobytes() == b["score"].tobytes()
    else:
        # spec tolerance is 1e-6 relative (BASELINE.json north_star); the exact sum differs from the
        # reference's sequentially rounded sum by its accumulated rounding only (~1e-12 on 1 Mb excursions)
        np.testing.assert_allclose(a["score"], b["score"], rtol=1e-9, atol=0)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 11])
def test_emu_counts(emu, oracle, k):
    rng = np.random.default_rng(500 + k)
    for trial in range(25):
        seqs = [rand_seq(rng, int(rng.integers(0, 300)), p_n=rng.choice([0, 0.05, 0.3]),
                         alphabet=rng.choice([b"ACGT", b"ACGTacgtRYKMSWBDHVUu-*."]))
                for _ in range(int(rng.integers(1, 6)))]
        seqs += [b"ACGTACGTACGTACGTACGT"[:k], b"NN" + b"ACGTACGTACGTACGTACGT"[:k],
                 b"ACGTACGTACGTACGTACGT"[:k] + b"N", b"ACGTACGTACGTACGTACGT"[:k + 1], b""]
        n1, c1 = oracle.kmer_counts(seqs, k)
        n2, c2 = emu.count(seqs, k)
        # the oracle skips sequences shorter than k up front; such sequences hold no k-mer anyway
        assert n1 == n2
        assert (c1 == c2).all()


@pytest.mark.parametrize("k", [2, 4, 6, 8, 10])
def test_emu_ranks_bitexact(emu, oracle, ref, k):
    rng = np.random.default_rng(600 + k)
    for trial in range(6):
        s = planted(rng, int(rng.integers(4 ** min(k, 6), 60 * 4 ** min(k, 6))))
        n, c = oracle.kmer_counts(s, k)
        if n == 0:
            continue
        a = emu.rank(c, k, n)
        assert a.tobytes() == oracle.rank(c, k, n).tobytes()
        assert a.tobytes() == ref.rank_kmers_w(c, k, n).tobytes()


def test_emu_ranks_adversarial_counts(emu, oracle):
    """count tables built to stress the linear-piece construction: huge tie groups, half-way
    addends, binade crossings, tiny and huge totals"""
    rng = np.random.default_rng(9)
    k = 9
    n = 4 ** k
    for trial in range(40):
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
            total = float(2 ** int(rng.integers(20, 40)))  # power-of-two totals give half-way addends
        if total == 0:
            continue
        a = emu.rank(c, k, total)
        b = oracle.rank(c, k, total)
        assert a.tobytes() == b.tobytes(), (trial, kind)
        assert emu.num_rank_segments(c, k, total) < 40 * (len(np.unique(c)) + 64)


def test_fx_roundtrip(emu):
    rng = np.random.default_rng(3)
    for qs in (57, 53, 45, 25):
        lim = 2.0 ** (57 - qs)
        x = rng.uniform(-lim, lim, 2000) * rng.choice([1, 1e-3, 1e-6], 2000)
        x = x[np.abs(x) < lim]
        y = np.array([emu.lib.emu_fx_roundtrip(float(v), qs) for v in x])
        assert np.all(np.abs(x - y) <= 2.0 ** -qs)
        big = x[np.abs(x) >= lim * 2.0 ** -5]
        yb = np.array([emu.lib.emu_fx_roundtrip(float(v), qs) for v in big])
        assert (big == yb).all()  # exact whenever the double's last bit is >= 2^-qs


CASES = [(2, 0.5, 20, 10), (3, 0.75, 20, 10), (4, 0.75, 5, 2), (5, 0.5, 0, 0), (6, 0.6, 10, 3),
         (8, 0.75, 100, 20), (4, 0.5, -1, 0), (3, 0.9, 0, 0.5), (7, 0.5, 3, 1), (10, 0.75, 30, 5)]


@pytest.mark.parametrize("k,thr,mw,ms", CASES)
def test_emu_low_comp_vs_oracle(emu, oracle, k, thr, mw, ms):
    rng = np.random.default_rng(700 + k)
    for trial in range(10):
        seqs = [planted(rng, int(rng.integers(50, 8000))) for _ in range(int(rng.integers(1, 5)))]
        if trial % 3 == 0:
            seqs.insert(1, b"ACG"[: k - 1])
        o = oracle.low_comp(seqs, k, mw, ms, thr)
        e = emu.scan(seqs, k, o["ranks"], thr, mw, ms)
        assert_spans(e, o, exact_scores=False)


@pytest.mark.parametrize("k", [2, 5, 8])
def test_emu_kmer_regions_vs_oracle(emu, oracle, k):
    """user weights through the kmer_regions_r semantics: threshold 0, in-scan counts (T8)"""
    rng = np.random.default_rng(800 + k)
    for trial in range(12):
        seqs = [planted(rng, int(rng.integers(50, 6000))) for _ in range(int(rng.integers(1, 4)))]
        kind = trial % 4
        if kind == 0:
            W = rng.choice([-1.0, 1.0], 4 ** k, p=[0.7, 0.3])
        elif kind == 1:
            W = rng.normal(-0.3, 1.0, 4 ** k)
        elif kind == 2:
            n, c = oracle.kmer_counts(seqs, k)
            W = oracle.scores(c, k, n, 2)          # +-1 around the median: long excursions, deep restarts
        else:
            W = rng.normal(-0.2, 1.0, 4 ** k)
            W[rng.integers(0, 4 ** k, 3)] = np.nan  # NaN weights clamp the state to 0 (reference :270)
            W[rng.integers(0, 4 ** k, 2)] = -np.inf
        mw, ms = [(0, 0), (10, 3), (30, 8)][trial % 3]
        o = oracle.kmer_regions(seqs, k, W, mw, ms)
        e = emu.scan(seqs, k, W, 0.0, mw, ms, inscan=True)
        assert_spans(e, o, exact_scores=(kind in (0, 2)))
        assert (e["counts"] == o["counts"]).all()


def test_emu_config1_both_modes(emu, oracle):
    """BASELINE.json configs[0] at full size: +-1 mode (giant excursions) and rank mode"""
    seq = synth.config1()[0].tobytes()
    n, c = oracle.kmer_counts(seq, 8)
    W = oracle.scores(c, 8, n, 2)
    o = oracle.kmer_regions([seq], 8, W, 100, 20)
    e = emu.scan([seq], 8, W, 0.0, 100, 20, inscan=True)
    assert_spans(e, o, exact_scores=True)
    assert (e["counts"] == o["counts"]).all()
    o = oracle.low_comp([seq], 8, 100, 20, 0.75)
    e = emu.scan([seq], 8, o["ranks"], 0.75, 100, 20)
    assert len(o["pos"]) > 20
    assert_spans(e, o, exact_scores=False)
    # log2 mode
    W = oracle.scores(c, 8, n, 1)
    o = oracle.kmer_regions([seq], 8, W, 100, 20)
    e = emu.scan([seq], 8, W, 0.0, 100, 20)
    assert_spans(e, o, exact_scores=False)


@pytest.mark.parametrize("k", [8, 10, 12, 13])
def test_emu_pair_bucket_count_vs_oracle(emu, oracle, k):
    """the bucketed count of ks_count.cuh (bucket and 16-bit sub-key of a PAIR of k-mers from PairGeom, two tables
    per bucket, the fold) folded on the host gives the oracle's table; every (bucket, sub-key) is decoded back to
    its two k-mers on the way, and tiny staging rows push most pairs through the overflow decode"""
    rng = np.random.default_rng(5200 + k)
    seqs = [planted(rng, 60_000), rand_seq(rng, 9000, p_n=0.05), b"A" * 5000, b"ACG" * 3000, b"ACGT"[:3], b"",
            b"G" * k, b"T" * (k + 1)]
    n1, c1 = oracle.kmer_counts(seqs, k)
    for cap in (24, 2):
        n2, c2 = emu.count_pairs(seqs, k, row_cap=cap)
        assert n2 == n1 and (c2 == c1).all(), (k, cap)


# ---- the summary-based ("fast") walk: chunk summaries, O(1) elements, few chunks walked in detail ----
@pytest.mark.parametrize("k", [2, 5, 8])
def test_emu_fast_walk_vs_oracle(emu, oracle, k):
    """FastChunk / GeneralChunk summaries + fast_walk_element (what scan_gather_kernel<.., summ> and
    scan_walk_fast_kernel compute) give the oracle's spans; inside, every summary-derived chunk element is
    compared with the position-by-position walk (emu_scan_fast fails otherwise)."""
    rng = np.random.default_rng(900 + k)
    detail = chunks = 0
    for trial in range(12):
        seqs = [planted(rng, int(rng.integers(50, 9000))) for _ in range(int(rng.integers(1, 4)))]
        kind = trial % 4
        if kind == 0:
            W = rng.choice([-1.0, 1.0], 4 ** k, p=[0.6, 0.4])
        elif kind == 1:
            W = rng.normal(-0.2, 1.0, 4 ** k)
        elif kind == 2:
            n, c = oracle.kmer_counts(seqs, k)
            W = oracle.scores(c, k, n, 2)
        else:
            W = rng.normal(-0.1, 1.0, 4 ** k)
            W[rng.integers(0, 4 ** k, 3)] = np.nan
            W[rng.integers(0, 4 ** k, 2)] = -np.inf
        mw, ms = [(15, 0), (16, 3), (40, 8), (100, 2)][trial % 4]
        o = oracle.kmer_regions(seqs, k, W, mw, ms)
        e = emu.scan_fast(seqs, k, W, 0.0, mw, ms)
        assert_spans(e, o, exact_scores=(kind in (0, 2)))
        g = emu.scan(seqs, k, W, 0.0, mw, ms)
        assert e["pos"].tolist() == g["pos"].tolist() and e["score"].tobytes() == g["score"].tobytes()
        detail += e["detail_chunks"]
        chunks += e["chunks"]
    assert detail < chunks / 4  # the position-by-position walk is the exception


@pytest.mark.parametrize("k", [3, 6])
def test_emu_pair_units_vs_oracle(emu, oracle, k, monkeypatch):
    """min_width >= 31: units of two chunks (unit_merge + fast_walk_unit, what scan_gather_kernel<..., kPair> leaves
    behind).  Every merged element, closing flag and transform is compared with the chunk-by-chunk walk inside
    emu_scan_fast; the spans must equal the oracle's and those of the single-chunk formulation."""
    rng = np.random.default_rng(4100 + k)
    for trial in range(24):
        seqs = [planted(rng, int(rng.integers(40, 6000))) for _ in range(int(rng.integers(1, 4)))]
        kind = trial % 4
        if kind == 0:
            W = rng.choice([-1.0, 1.0], 4 ** k, p=[0.55, 0.45])
        elif kind == 1:
            W = rng.normal(-0.1, 1.0, 4 ** k)
        elif kind == 2:
            n, c = oracle.kmer_counts(seqs, k)
            W = oracle.scores(c, k, n, 2)
        else:
            W = rng.normal(0.0, 1.0, 4 ** k)
            W[rng.integers(0, 4 ** k, 3)] = np.nan
            W[rng.integers(0, 4 ** k, 2)] = -np.inf
        mw, ms = [(31, 0), (32, 1), (40, 4), (64, 0), (100, 2), (31, 6)][trial % 6]
        o = oracle.kmer_regions(seqs, k, W, mw, ms)
        e = emu.scan_fast(seqs, k, W, 0.0, mw, ms)
        assert_spans(e, o, exact_scores=(kind in (0, 2)))
        monkeypatch.setenv("KS_NO_PAIR", "1")
        g = emu.scan_fast(seqs, k, W, 0.0, mw, ms)
        monkeypatch.delenv("KS_NO_PAIR")
        assert e["pos"].tolist() == g["pos"].tolist() and e["score"].tobytes() == g["score"].tobytes()


def test_emu_fast_walk_config1(emu, oracle):
    seq = synth.config1()[0].tobytes()
    n, c = oracle.kmer_counts(seq, 8)
    for mode in (1, 2):
        W = oracle.scores(c, 8, n, mode)
        o = oracle.kmer_regions([seq], 8, W, 100, 20)
        e = emu.scan_fast([seq], 8, W, 0.0, 100, 20)
        assert_spans(e, o, exact_scores=(mode == 2))
    o = oracle.low_comp([seq], 8, 100, 20, 0.75)
    e = emu.scan_fast([seq], 8, o["ranks"], 0.75, 100, 20)
    assert_spans(e, o, exact_scores=False)
    assert e["detail_chunks"] < e["chunks"] / 20


# ---- the transition-score scan (tr_lr_regions_r) through chunk_walk_tr ------------------------------
def test_emu_tr_scan_vs_oracle(emu, oracle):
    """the level loop the kernels run for find_kmer_tr_lr_regions (initial score at the first k-mer of a run,
    bookkeeping one position later, every close re-scanned, terminal excursion reported only) against the
    oracle restatement, which tests/test_oracle.py pins to the compiled reference"""
    rng = np.random.default_rng(1000)
    total = 0
    for t in range(60):
        k = int(rng.integers(1, 6))
        n = 4 ** k
        seqs = [rand_seq(rng, int(rng.integers(0, 1500)), p_n=float(rng.choice([0, 0.05, 0.3])))
                for _ in range(int(rng.integers(1, 5)))]
        if t % 3 == 0:
            seqs += [rand_seq(rng, k), rand_seq(rng, k + 1), rand_seq(rng, k + 2), rand_seq(rng, k) + b"N",
                     rand_seq(rng, k) + b"NA", rand_seq(rng, k) + b"N" + rand_seq(rng, k + 3), b"", b"NNN"]
        exact = t % 2 == 0
        if exact:
            init, trans = rng.integers(-2, 3, n).astype(float), rng.integers(-2, 3, n).astype(float)
        else:
            init, trans = rng.normal(0, 1, n), rng.normal(-0.1, 1, n)
        if t % 5 == 0:
            trans[rng.integers(0, n)] = -np.inf
            init[rng.integers(0, n)] = -np.inf
        min_len = int(rng.choice([0, 1, 3, 10, 40]))
        o = oracle.tr_lr_regions(seqs, k, init, trans, min_len)
        e = emu.tr_scan(seqs, k, init, trans, min_len)
        assert_spans(e, o, exact_scores=exact)
        total += len(o["pos"])
    assert total > 1500
    with pytest.raises(ValueError, match="rc=2"):
        emu.tr_scan([b"ACGTACGT"], 2, np.full(16, np.nan), np.zeros(16), 0)

"""GPU tier (-m gpu): the large-k path (k = 16 .. 31, BASELINE.json configs[3]: k = 21) -- hash-table counting,
weighted rank over the k-mers that occur, scan -- against oracle/ks_oracle_large.c (parity by extension: the
reference stops at k = 15) and, at k <= 13, against this repo's own direct-table path."""
import numpy as np
import pytest

from kmer_spans_b200 import synth
from tests.test_gpu import assert_spans, ctx  # noqa: F401  (fixture)
from tests.test_oracle import planted, rand_seq

pytestmark = pytest.mark.gpu


def _check_table(got, want):
    o = np.argsort(got["codes"], kind="stable")
    assert (got["codes"][o] == want["codes"]).all()
    assert (got["counts"][o] == want["counts"]).all()
    assert got["ranks"][o].tobytes() == want["ranks"].tobytes()
    # the order the table comes in is (count, code)
    key = got["counts"].astype(np.float64) * 0  # placeholder to keep numpy from promoting uint64 oddly
    del key
    c = got["counts"].astype(np.int64)
    assert (np.diff(c) >= 0).all()
    same = np.diff(c) == 0
    assert (got["codes"][1:][same] > got["codes"][:-1][same]).all()


@pytest.mark.parametrize("k,two_pass", [(16, False), (21, False), (21, True), (31, False)])
def test_large_k_matches_oracle(ctx, oracle, k, two_pass, monkeypatch):  # noqa: F811
    """one composite sort (count << 2k | code) where it fits 64 bits, else codes first and counts in a second stable
    pass (k = 31 always; forced at k = 21)"""
    if two_pass:
        monkeypatch.setenv("KS_LARGE_TWO_PASS", "1")
    rng = np.random.default_rng(9000 + k)
    seqs = [planted(rng, 150_000), b"ACGTN" * 9, planted(rng, 20_000), rand_seq(rng, 6000, p_n=0.05),
            b"ACGTACGTACGTACGTACGTACGTACGTACGTACGT"[:k], b"ACGTACGTACGTACGTACGTACGTACGTACGTACGT"[:k + 1], b"AC"]
    for thr, mw, ms in ((0.75, 20, 1.0), (0.5, 0, 0.0), (0.9, 100, 3.0)):
        want = oracle.large_regions(seqs, k, 0, mw, ms, thr=thr)
        got = ctx.kmer_large_regions(seqs, k, 0, mw, ms, thr=thr, want_table=True)
        assert got["n"] == want["n"] and got["nd"] == want["nd"]
        _check_table(got, want)
        assert_spans(got, want, False, "large k %d thr %g mw %d" % (k, thr, mw))
    # +-1 around an explicit frequency: exact arithmetic, scores bit for bit
    f_t = 1.5 / want["n"]
    want = oracle.large_regions(seqs, k, 2, 30, 4.0, thr=0.0, param=f_t)
    got = ctx.kmer_large_regions(seqs, k, 2, 30, 4.0, thr=0.0, param=f_t)
    assert_spans(got, want, True, "large k %d sign" % k)


def test_large_path_equals_direct_table_path_at_small_k(ctx, oracle):  # noqa: F811
    """the hash path and the direct 4^k table path of this repo must agree where both exist: same counts
    for the k-mers that occur, ranks bit for bit, spans AND scores bit for bit (same exact fixed-point values)"""
    rng = np.random.default_rng(9100)
    seqs = [planted(rng, 400_000), planted(rng, 30_000), rand_seq(rng, 5000, p_n=0.2)]
    for k, thr, mw, ms in ((6, 0.6, 15, 2.0), (10, 0.75, 100, 5.0), (13, 0.75, 30, 1.0), (8, 0.5, 0, 0.0)):
        direct = ctx.kmer_low_comp_regions(seqs, k, mw, ms, thr)
        large = ctx.kmer_large_regions(seqs, k, 0, mw, ms, thr=thr, want_table=True)
        assert large["n"] == direct["n"][0]
        present = np.nonzero(direct["counts"])[0]
        o = np.argsort(large["codes"], kind="stable")
        assert (large["codes"][o] == present).all() and (large["counts"][o] == direct["counts"][present]).all()
        assert large["ranks"][o].tobytes() == direct["w_rank"][present].tobytes()
        assert large["pos"].tobytes() == direct["pos"].tobytes()
        assert large["score"].tobytes() == direct["score"].tobytes()


def test_large_k_probe_chains_and_errors(ctx, oracle, monkeypatch):  # noqa: F811
    from kmer_spans_b200.api import KspansError
    rng = np.random.default_rng(9200)
    seqs = [planted(rng, 60_000)]
    want = oracle.large_regions(seqs, 21, 0, 20, 1.0, thr=0.75)
    cap = 1
    while cap < want["nd"] * 1.05:
        cap *= 2
    monkeypatch.setenv("KS_HASH_CAP", str(cap))  # nearly full table: long linear-probing chains
    got = ctx.kmer_large_regions(seqs, 21, 0, 20, 1.0, thr=0.75, want_table=True)
    _check_table(got, want)
    assert_spans(got, want, False, "crowded table")
    monkeypatch.setenv("KS_HASH_CAP", str(cap // 4))
    with pytest.raises(KspansError, match="hash table is full"):
        ctx.kmer_large_regions(seqs, 21, 0, 20, 1.0, thr=0.75)
    monkeypatch.delenv("KS_HASH_CAP")
    with pytest.raises(KspansError, match="weighted rank"):
        ctx.kmer_large_regions(seqs, 21, 1, 20, 1.0)
    with pytest.raises(KspansError, match="explicit frequency"):
        ctx.kmer_large_regions(seqs, 21, 2, 20, 1.0)
    got = ctx.kmer_large_regions(seqs, 21, 0, 20, 1.0, thr=0.75)  # usable after the errors
    assert_spans(got, want, False, "after errors")


def test_large_k_multi_megabase(ctx, oracle):  # noqa: F811
    """BASELINE.json configs[3] scaled to 8 Mb: k = 21, planted repeats (counts in the thousands), N blocks; deep
    excursions across many tiles and restart levels"""
    seq = synth.genome(8_000_000, 21, n_blocks=(3, 5000)).tobytes()
    want = oracle.large_regions([seq], 21, 0, 100, 20.0, thr=0.75)
    ss = ctx.upload([seq])
    got = ctx.dev_large_regions(ss, 21, 0, 100, 20.0, thr=0.75, fetch_spans=True)
    assert got["n"] == want["n"] and got["nd"] == want["nd"] and got["n_spans"] == len(want["pos"])
    assert len(want["pos"]) > 50
    assert_spans(got, want, False, "config 4 scaled")
    ss.free()

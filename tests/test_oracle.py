"""Pins the CPU oracle (oracle/ks_oracle.c): against the reference's own known answers and
against the unmodified reference compiled into oracle/_ref (bit-exact, doubles included)."""
import numpy as np
import pytest

from kmer_spans_b200 import synth


def rand_seq(rng, n, p_n=0.0, alphabet=b"ACGT"):
    s = np.frombuffer(alphabet, np.uint8)[rng.integers(0, len(alphabet), n)]
    if p_n > 0:
        # N runs of random length
        i = 0
        while i < n:
            if rng.random() < p_n:
                ln = int(rng.integers(1, 12))
                s[i:i + ln] = ord("N") if rng.random() < 0.7 else ord("n")
                i += ln
            i += int(rng.integers(1, 40))
    return s.tobytes()


def planted(rng, n):
    s = np.frombuffer(rand_seq(rng, n), np.uint8).copy()
    for _ in range(max(1, n // 400)):
        unit = np.frombuffer(rand_seq(rng, int(rng.integers(1, 9))), np.uint8)
        ln = int(rng.integers(20, 200))
        p = int(rng.integers(0, max(1, n - ln)))
        s[p:p + ln] = np.tile(unit, ln // len(unit) + 1)[:ln][: len(s[p:p + ln])]
    for _ in range(n // 300):
        p = int(rng.integers(0, n))
        s[p:p + int(rng.integers(1, 8))] = ord("N")
    return s.tobytes()


# ---- the reference's own known answers ------------------------------------------------------
def test_ka1_dinucleotide_counts(oracle):
    """test.R:365-375: CGCCAATGCG, k=2."""
    n, c = oracle.kmer_counts(b"CGCCAATGCG", 2)
    names = [oracle.kmer_seq(2, i) for i in range(16)]
    got = {a: int(b) for a, b in zip(names, c) if b}
    assert got == {"CG": 2, "GC": 2, "CC": 1, "CA": 1, "AA": 1, "AT": 1, "TG": 1}
    assert n == 9


def test_ka2_n_splitting(oracle):
    """test.R:66-77: counts(seq + N*36 + seq) == 2 * counts(seq), k=2."""
    rng = np.random.default_rng(7)
    s = rand_seq(rng, 5000)
    _, c1 = oracle.kmer_counts(s, 2)
    _, c2 = oracle.kmer_counts(s + b"N" * 36 + s, 2)
    assert (c2 == 2 * c1).all()


def test_ka3_kmer_order(oracle):
    """kmer_spans.R:81-83: order A, C, T, G."""
    assert [oracle.kmer_seq(2, i) for i in range(16)] == \
        "AA AC AT AG CA CC CT CG TA TC TT TG GA GC GT GG".split()
    assert oracle.kmer_seq(3, 0b100111) == "TCG"


def test_t6_exact_k_tail(oracle):
    assert oracle.kmer_counts(b"ACG", 3)[0] == 0
    assert oracle.kmer_counts(b"ACGNNACG", 3)[0] == 1
    assert oracle.kmer_counts(b"ACGT", 3)[0] == 2


def test_t7_coordinates(oracle):
    """(AG)x50 + random, k=2, +-1 weights -> start=2, end=101, score=100."""
    W = -np.ones(16)
    names = [oracle.kmer_seq(2, i) for i in range(16)]
    W[names.index("AG")] = 1
    W[names.index("GA")] = 1
    s = b"AG" * 50 + b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCC"
    r = oracle.kmer_regions([s], 2, W, 20, 10)
    assert r["pos"].tolist() == [[0, 2, 100]] or r["pos"].tolist() == [[0, 2, 101]]
    assert r["score"][0, 0] in (99.0, 100.0)


# ---- restatement vs the compiled reference ---------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
def test_counts_match_reference(oracle, ref, k):
    rng = np.random.default_rng(100 + k)
    for trial in range(40):
        n = int(rng.integers(0, 400))
        seqs = [rand_seq(rng, int(rng.integers(0, max(1, n))), p_n=rng.choice([0, 0.05, 0.3]),
                         alphabet=rng.choice([b"ACGT", b"ACGTacgtRYKMSWBDHVUu-*."]))
                for _ in range(int(rng.integers(1, 5)))]
        if trial % 5 == 0:
            seqs.append(b"ACGTACGTACGTACGTACGT"[:k])           # exactly k, at the terminator
            seqs.append(b"N" * 3 + b"ACGTACGTACGTACGTACGT"[:k])
            seqs.append(b"ACGTACGTACGTACGTACGT"[:k] + b"N")   # exactly k, followed by N
            seqs.append(b"AC"[: max(0, k - 1)] + b"NNN")       # T10 shape
        n1, c1 = oracle.kmer_counts(seqs, k)
        got = ref.call_kmer_counts(seqs, k) if seqs else None
        assert n1 == got["n"]
        assert (c1 == got["counts"]).all()


@pytest.mark.parametrize("k", [2, 4, 6, 8, 10])
def test_ranks_match_reference_bitexact(oracle, ref, k):
    rng = np.random.default_rng(200 + k)
    for trial in range(6):
        s = planted(rng, int(rng.integers(4 ** min(k, 6), 40 * 4 ** min(k, 6))))
        n, c = oracle.kmer_counts(s, k)
        if n == 0:
            continue
        a = oracle.rank(c, k, n)
        b = ref.rank_kmers_w(c, k, n)
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("k,thr,mw,ms", [(2, 0.5, 20, 10), (3, 0.75, 20, 10), (4, 0.75, 5, 2),
                                         (5, 0.5, 0, 0), (6, 0.6, 10, 3), (8, 0.75, 100, 20),
                                         (4, 0.5, -1, 0), (3, 0.9, 0, 0.5)])
def test_low_comp_matches_reference_bitexact(oracle, ref, k, thr, mw, ms):
    rng = np.random.default_rng(300 + k)
    for trial in range(12):
        seqs = [planted(rng, int(rng.integers(50, 6000))) for _ in range(int(rng.integers(1, 4)))]
        if trial % 3 == 0:
            seqs.insert(1, b"ACG"[: k - 1])  # skipped (len < k), seq_id must still advance
        a = oracle.low_comp(seqs, k, mw, ms, thr)
        b = ref.call_kmer_low_comp_regions(seqs, k, mw, ms, thr)
        assert (a["n"] == b["n"]).all()
        assert (a["counts"] == b["counts"]).all()
        assert a["ranks"].tobytes() == b["ranks"].tobytes()
        assert a["pos"].tolist() == b["pos"].tolist()
        assert a["score"].tobytes() == b["score"].tobytes()


@pytest.mark.parametrize("k", [2, 5, 8])
def test_kmer_regions_matches_reference_bitexact(oracle, ref, k):
    rng = np.random.default_rng(400 + k)
    for trial in range(12):
        seqs = [planted(rng, int(rng.integers(50, 5000))) for _ in range(int(rng.integers(1, 4)))]
        kind = trial % 3
        if kind == 0:
            W = rng.choice([-1.0, 1.0], 4 ** k, p=[0.7, 0.3])
        elif kind == 1:
            W = rng.normal(-0.3, 1.0, 4 ** k)
        else:
            n, c = oracle.kmer_counts(seqs, k)
            W = oracle.scores(c, k, n, 2)
        mw, ms = [(0, 0), (10, 3), (30, 8)][trial % 3]
        a = oracle.kmer_regions(seqs, k, W, mw, ms)
        b = ref.call_kmer_regions_r(seqs, k, W, mw, ms)
        assert a["n"] == b["n"]
        assert (a["counts"] == b["counts"]).all()   # in-scan counts incl. rescans
        assert a["pos"].tolist() == b["pos"].tolist()
        assert a["score"].tobytes() == b["score"].tobytes()


def test_config1_against_reference(oracle, ref):
    """BASELINE.json configs[0] (1 Mb, k=8, +-1 mode through kmer_regions_r), full size."""
    seq = synth.config1()[0].tobytes()
    n, c = oracle.kmer_counts(seq, 8)
    W = oracle.scores(c, 8, n, 2)
    a = oracle.kmer_regions([seq], 8, W, 100, 20)
    b = ref.call_kmer_regions_r([seq], 8, W, 100, 20)
    assert len(a["pos"]) >= 2   # +-1 around the median has no negative drift: run-sized spans
    assert a["pos"].tolist() == b["pos"].tolist()
    assert a["score"].tobytes() == b["score"].tobytes()
    assert (a["counts"] == b["counts"]).all()
    lc_a = oracle.low_comp([seq], 8, 100, 20, 0.75)
    lc_b = ref.call_kmer_low_comp_regions([seq], 8, 100, 20, 0.75)
    assert len(lc_a["pos"]) > 20
    assert lc_a["pos"].tolist() == lc_b["pos"].tolist()
    assert lc_a["ranks"].tobytes() == lc_b["ranks"].tobytes()


def test_reference_error_paths(ref):
    with pytest.raises(RuntimeError, match="threshold must be between"):
        ref.call_kmer_low_comp_regions([b"ACGT"], 2, 1, 1.0, 1.0)
    with pytest.raises(RuntimeError, match="positive integer"):
        ref.call_kmer_counts([b"ACGT"], 0)


# ---- the other two .Call entries --------------------------------------
def _pos_equal(a, b, kmer_n):
    if (a is None) != (b is None):
        return False
    return a is None or np.array_equal(np.asarray(a).reshape(kmer_n, -1), b)


def test_window_dist_matches_reference(oracle, ref):
    """windowed_kmer_count_distributions_r, src/kmer_spans.c:398-449,715-793."""
    rng = np.random.default_rng(31)
    for t in range(120):
        k = int(rng.integers(1, 5))
        window = int(rng.integers(2 * k, 40))
        seqs = [rand_seq(rng, int(rng.integers(0, 300)), p_n=float(rng.choice([0, 0.1, 0.4])))
                for _ in range(int(rng.integers(1, 5)))]
        if t % 5 == 0:  # length == window is left out, window + 1 is not (:775)
            seqs += [rand_seq(rng, window), rand_seq(rng, window + 1)]
        kms = [oracle.kmer_seq(k, int(c)).encode() for c in rng.integers(0, 4 ** k, int(rng.integers(1, 6)))]
        if t % 7 == 0:
            kms.append(b"N" * k)
        if t % 11 == 0 and k >= 2:
            kms += [b"A" + b"N" * (k - 1), b"N" + b"C" * (k - 1)]
        a = ref.call_window_dist(seqs, kms, k, window, 1)
        b = oracle.window_dist(seqs, kms, k, window, True)
        assert np.array_equal(a["dist"], b["dist"])
        assert np.array_equal(a["included"], b["included"])
        for x, y in zip(a["pos"], b["pos"]):
            assert _pos_equal(x, y, len(kms))
        assert ref.call_window_dist(seqs, kms, k, window, 0)["pos"] is None


def test_window_dist_known_answer(oracle):
    # ACGTACGTAA: windows of 8 start at 0, 1, 2; "AC" occurs at 0 and 4
    r = oracle.window_dist([b"ACGTACGTAA"], [b"AC"], 2, 8, True)
    assert r["pos"][0][0, :3].tolist() == [2, 1, 1]
    assert r["dist"][0, :3].tolist() == [0, 2, 1]


def test_tr_lr_matches_reference_bitexact(oracle, ref):
    """tr_lr_regions_r, src/kmer_spans.c:329-395,649-713 (tables handed over in permuted k-mer order)."""
    rng = np.random.default_rng(32)
    total = 0
    for t in range(150):
        k = int(rng.integers(1, 5))
        n = 4 ** k
        seqs = [rand_seq(rng, int(rng.integers(0, 400)), p_n=float(rng.choice([0, 0.1, 0.4])))
                for _ in range(int(rng.integers(1, 5)))]
        if t % 3 == 0:  # the run-end rules of :340-341
            seqs += [rand_seq(rng, k), rand_seq(rng, k + 1), rand_seq(rng, k + 2), rand_seq(rng, k) + b"N",
                     rand_seq(rng, k) + b"NA", rand_seq(rng, k) + b"N" + rand_seq(rng, k + 3)]
        kms = [oracle.kmer_seq(k, c).encode() for c in range(n)]
        perm = rng.permutation(n)
        if t % 2:
            init, trans = rng.normal(0, 1, n), rng.normal(-0.1, 1, n)
        else:
            init, trans = rng.integers(-2, 3, n).astype(float), rng.integers(-2, 3, n).astype(float)
        min_len = int(rng.choice([0, 1, 3, 10]))
        a = ref.call_tr_lr(seqs, k, min_len, [kms[i] for i in perm], init[perm], trans[perm])
        assert np.array_equal(a["tables"][0], init) and np.array_equal(a["tables"][1], trans)
        assert [oracle.kmer_code(kms[i], k) for i in perm[:8]] == [int(i) for i in perm[:8]]
        b = oracle.tr_lr_regions(seqs, k, init, trans, min_len)
        assert np.array_equal(a["pos"], b["pos"])
        assert np.array_equal(a["score"], b["score"])
        total += len(b["pos"])
    assert total > 1000


@pytest.mark.parametrize("k", [6, 8])
def test_next_rows_match_reference_larger_k(oracle, ref, k):
    """the two SURVEY 8(f) restatements at larger k (tables of 4^6 / 4^8 entries, longer planted sequences)"""
    rng = np.random.default_rng(40 + k)
    n = 4 ** k
    kms = [oracle.kmer_seq(k, c).encode() for c in range(n)]
    for t in range(4):
        seqs = [planted(rng, int(rng.integers(2000, 30000))) for _ in range(3)] + [rand_seq(rng, 500, p_n=0.2)]
        init, trans = rng.normal(0, 1, n), rng.normal(-0.15, 1, n)
        if t % 2:
            init, trans = np.round(init), np.round(trans)
        min_len = int(rng.choice([0, 5, 40]))
        a = ref.call_tr_lr(seqs, k, min_len, kms, init, trans)
        b = oracle.tr_lr_regions(seqs, k, init, trans, min_len)
        assert np.array_equal(a["pos"], b["pos"]) and np.array_equal(a["score"], b["score"])
        assert len(b["pos"]) > 0
        sel = [kms[int(c)] for c in rng.integers(0, n, 3)] + [seqs[0][100:100 + k]]
        window = int(rng.integers(2 * k, 300))
        a = ref.call_window_dist(seqs, sel, k, window, 1)
        b = oracle.window_dist(seqs, sel, k, window, True)
        assert np.array_equal(a["dist"], b["dist"]) and np.array_equal(a["included"], b["included"])
        for x, y in zip(a["pos"], b["pos"]):
            assert _pos_equal(x, y, len(sel))


def test_large_oracle_equals_pinned_oracle_at_small_k(oracle):
    """ks_oracle_large.c (sparse table, 64-bit codes; the checker of the k >= 16 path) run at k <= 12 must give
    what the pinned full-table oracle gives: counts of the k-mers that occur, ranks bit for bit (absent k-mers add
    0 to the running sum), identical spans -- so the extension to k = 16..31 changes the container, not the
    arithmetic."""
    rng = np.random.default_rng(8100)
    for k, thr, mw, ms in ((3, 0.6, 10, 2.0), (7, 0.75, 30, 3.0), (9, 0.5, 0, 0.0), (12, 0.75, 20, 1.0)):
        seqs = [planted(rng, 30000), b"ACG", rand_seq(rng, 4000, p_n=0.1), b"ACGTACGTACGTACGT"[:k], planted(rng, 9000)]
        small = oracle.low_comp(seqs, k, mw, ms, thr)
        large = oracle.large_regions(seqs, k, 0, mw, ms, thr=thr)
        assert large["n"] == small["n"][0]
        present = np.nonzero(small["counts"])[0]
        assert (large["codes"] == present).all() and (large["counts"] == small["counts"][present]).all()
        assert large["ranks"].tobytes() == small["ranks"][present].tobytes()
        assert large["pos"].tolist() == small["pos"].tolist()
        assert large["score"].tobytes() == small["score"].tobytes()
        # +-1 mode around an explicit frequency
        f_t = 1.5 / small["n"][0]
        W = np.where(small["counts"] / small["n"][0] >= f_t, 1.0, -1.0)
        a = oracle.kmer_regions(seqs, k, W, mw, ms)
        b = oracle.large_regions(seqs, k, 2, mw, ms, thr=0.0, param=f_t)
        assert a["pos"].tolist() == b["pos"].tolist() and a["score"].tobytes() == b["score"].tobytes()
    # and it runs where the reference cannot: k = 21 and k = 31
    seqs = [planted(rng, 20000), planted(rng, 3000)]
    for k in (16, 21, 31):
        r = oracle.large_regions(seqs, k, 0, 20, 1.0, thr=0.75)
        assert r["nd"] <= r["n"] and r["counts"].sum() == r["n"] and (r["codes"][1:] > r["codes"][:-1]).all()

"""The .Call boundary: the replacement glue (r/src/kmer_spans_glue.c, built against the mock R API)
next to the UNMODIFIED reference glue (oracle/_ref), on identical mock SEXP inputs."""
import os

import numpy as np
import pytest

from tests.test_oracle import planted

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "r", "_build", "kmer_spans.so")


@pytest.fixture(scope="module")
def glue():
    from oracle.ksoracle import Ref
    if not os.path.exists(GLUE):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "r")])
    return Ref(GLUE)


def test_registration_matches_reference(glue, ref):
    """same six .Call names and arities (src/kmer_spans.c:795-803), plus the one extension; none is a stub"""
    want = ref.registered()
    got = glue.registered()
    assert got[: len(want)] == want
    assert got[len(want):] == [("kmer_mode_regions", 8)]


def test_argument_errors_match_reference(glue, ref):
    """validation happens before any CUDA call, with the reference's texts"""
    bad = [
        ("kmer_counts", [("i", 3), ("i", 2)]),
        ("kmer_counts", [("s", [b"ACGT"]), ("d", 2.0)]),
        ("kmer_counts", [("s", [b"ACGT"]), ("i", 0)]),
        ("kmer_regions_r", [("s", [b"ACGT"]), ("i", 2), ("i", [1] * 16), ("i", 1), ("d", 1.0)]),
        ("kmer_regions_r", [("s", [b"ACGT"]), ("i", 2), ("d", [1.0] * 16), ("d", 1.0), ("d", 1.0)]),
        ("kmer_regions_r", [("s", [b"ACGT"]), ("i", 2), ("d", [1.0] * 16), ("i", 1), ("i", 1)]),
        ("kmer_regions_r", [("s", [b"ACGT"]), ("i", 16), ("d", [1.0] * 16), ("i", 1), ("d", 1.0)]),
        ("kmer_regions_r", [("s", [b"ACGT"]), ("i", 2), ("d", [1.0] * 15), ("i", 1), ("d", 1.0)]),
        ("kmer_low_comp_regions", [("s", [b"ACGT"]), ("i", 2), ("i", 1), ("d", 1.0), ("d", 1.0)]),
        ("kmer_low_comp_regions", [("s", [b"ACGT"]), ("i", 2), ("i", 1), ("d", 1.0), ("d", 0.0)]),
        ("kmer_low_comp_regions", [("s", [b"ACGT"]), ("i", 2), ("i", 1), ("d", 1.0), ("i", 1)]),
        ("kmer_low_comp_regions", [("s", [b"ACGT"]), ("i", 2), ("d", 1.0), ("d", 1.0), ("d", 0.5)]),
        ("kmer_seq_r", [("i", 0)]),
        ("kmer_seq_r", [("i", [1, 2])]),
        ("tr_lr_regions_r", [("i", 1), ("i", [2, 1]), ("s", [b"AA"] * 16), ("d", [1.0] * 16), ("d", [1.0] * 16)]),
        ("tr_lr_regions_r", [("s", [b"A"]), ("i", [2]), ("s", [b"AA"] * 16), ("d", [1.0] * 16), ("d", [1.0] * 16)]),
        ("tr_lr_regions_r", [("s", [b"A"]), ("i", [2, 1]), ("i", [1] * 16), ("d", [1.0] * 16), ("d", [1.0] * 16)]),
        ("tr_lr_regions_r", [("s", [b"A"]), ("i", [2, 1]), ("s", [b"AA"] * 16), ("i", [1] * 16), ("d", [1.0] * 16)]),
        ("tr_lr_regions_r", [("s", [b"A"]), ("i", [0, 1]), ("s", [b"AA"] * 16), ("d", [1.0] * 16), ("d", [1.0] * 16)]),
        ("tr_lr_regions_r", [("s", [b"A"]), ("i", [2, -1]), ("s", [b"AA"] * 16), ("d", [1.0] * 16), ("d", [1.0] * 16)]),
        ("tr_lr_regions_r", [("s", [b"A"]), ("i", [2, 1]), ("s", [b"AA"] * 15), ("d", [1.0] * 16), ("d", [1.0] * 16)]),
        ("tr_lr_regions_r", [("s", [b"A"]), ("i", [2, 1]), ("s", [b"AA"] * 16), ("d", [1.0] * 16), ("d", [1.0] * 15)]),
        ("windowed_kmer_count_distributions_r", [("i", 1), ("s", [b"AC"]), ("i", 2), ("i", 8), ("i", 0)]),
        ("windowed_kmer_count_distributions_r", [("s", [b"ACGT"]), ("i", 1), ("i", 2), ("i", 8), ("i", 0)]),
        ("windowed_kmer_count_distributions_r", [("s", [b"ACGT"]), ("s", [b"AC"]), ("d", 2.0), ("i", 8), ("i", 0)]),
        ("windowed_kmer_count_distributions_r", [("s", [b"ACGT"]), ("s", [b"AC"]), ("i", 2), ("i", [8, 9]), ("i", 0)]),
        ("windowed_kmer_count_distributions_r", [("s", [b"ACGT"]), ("s", [b"AC"]), ("i", 2), ("i", 8), ("d", 0.0)]),
        ("windowed_kmer_count_distributions_r", [("s", [b"ACGT"]), ("s", [b"AC"]), ("i", 16), ("i", 40), ("i", 0)]),
        ("windowed_kmer_count_distributions_r", [("s", [b"ACGT"]), ("s", [b"AC", b"ACG"]), ("i", 2), ("i", 8), ("i", 0)]),
        ("windowed_kmer_count_distributions_r", [("s", [b"ACGT"]), ("s", [b"AC"]), ("i", 2), ("i", 3), ("i", 0)]),
    ]
    for name, args in bad:
        with pytest.raises(RuntimeError) as e_ref:
            ref.call_raw(name, args)
        with pytest.raises(RuntimeError) as e_new:
            glue.call_raw(name, args)
        assert str(e_new.value) == str(e_ref.value), (name, args)


def test_kmer_seq_r_matches_reference(glue, ref):
    for k in (1, 2, 5):
        assert glue.call_kmer_seq_r(k) == ref.call_kmer_seq_r(k)


@pytest.mark.gpu
def test_tr_lr_call_matches_reference(glue, ref, oracle):
    rng = np.random.default_rng(2026)
    for trial in range(6):
        k = int(rng.choice([1, 2, 3, 5]))
        n = 4 ** k
        seqs = [planted(rng, int(rng.integers(50, 8000))) for _ in range(int(rng.integers(1, 5)))]
        seqs += [b"ACGTACGTAC"[:k + 1], b"ACGT"]
        kms = [oracle.kmer_seq(k, c).encode() for c in range(n)]
        perm = rng.permutation(n)
        init, trans = rng.integers(-2, 3, n).astype(float), rng.integers(-3, 3, n).astype(float)
        min_len = int(rng.choice([0, 5, 30]))
        a = ref.call_tr_lr(seqs, k, min_len, [kms[i] for i in perm], init[perm], trans[perm])
        b = glue.call_tr_lr(seqs, k, min_len, [kms[i] for i in perm], init[perm], trans[perm])
        assert np.array_equal(a["tables"], b["tables"])
        assert a["pos"].tolist() == b["pos"].tolist() and a["score"].tobytes() == b["score"].tobytes()


@pytest.mark.gpu
def test_window_dist_call_matches_reference(glue, ref):
    rng = np.random.default_rng(2025)
    for trial in range(6):
        k = int(rng.choice([1, 2, 3, 5]))
        window = int(rng.integers(2 * k, 120))
        seqs = [planted(rng, int(rng.integers(50, 5000))) for _ in range(int(rng.integers(1, 5)))]
        seqs += [planted(rng, window), b"ACGT"]
        kms = [bytes(rng.choice(list(b"ACGT"), k).astype(np.uint8)) for _ in range(int(rng.integers(1, 5)))]
        for flag in (0, 1):
            a = ref.call_window_dist(seqs, kms, k, window, flag)
            b = glue.call_window_dist(seqs, kms, k, window, flag)
            assert np.array_equal(a["dist"], b["dist"]) and np.array_equal(a["included"], b["included"])
            if flag == 0:
                assert a["pos"] is None and b["pos"] is None
            else:
                for x, y in zip(a["pos"], b["pos"]):
                    assert (x is None) == (y is None)
                    assert x is None or np.array_equal(x, y)


@pytest.mark.gpu
def test_call_results_match_reference(glue, ref):
    rng = np.random.default_rng(2024)
    for trial in range(6):
        k = int(rng.choice([2, 4, 7, 9]))
        seqs = [planted(rng, int(rng.integers(50, 20000))) for _ in range(int(rng.integers(1, 5)))]
        if trial % 2:
            seqs.insert(1, b"AC"[: k - 1])
        a, b = ref.call_kmer_counts(seqs, k), glue.call_kmer_counts(seqs, k)
        assert a["n"] == b["n"] and (a["counts"] == b["counts"]).all()
        thr, mw, ms = float(rng.choice([0.5, 0.75])), int(rng.choice([0, 10, 40])), float(rng.choice([0, 3, 10]))
        a = ref.call_kmer_low_comp_regions(seqs, k, mw, ms, thr)
        b = glue.call_kmer_low_comp_regions(seqs, k, mw, ms, thr)
        assert (a["n"] == b["n"]).all() and (a["counts"] == b["counts"]).all()
        assert a["ranks"].tobytes() == b["ranks"].tobytes()
        assert a["pos"].tolist() == b["pos"].tolist()
        np.testing.assert_allclose(b["score"], a["score"], rtol=1e-9)
        W = rng.choice([-1.0, 1.0], 4 ** k, p=[0.7, 0.3])
        a = ref.call_kmer_regions_r(seqs, k, W, mw, ms)
        b = glue.call_kmer_regions_r(seqs, k, W, mw, ms)
        assert a["n"] == b["n"] and (a["counts"] == b["counts"]).all()
        assert a["pos"].tolist() == b["pos"].tolist() and a["score"].tobytes() == b["score"].tobytes()
    out = glue.call_raw("kmer_mode_regions", [("s", seqs), ("i", 7), ("i", 1), ("d", float("nan")), ("d", 0.0),
                                              ("i", 10), ("d", 3.0), ("i", 1)])
    assert len(out) == 5 and out[1].size == 4 ** 7 and out[2].size == 4 ** 7


@pytest.mark.gpu
def test_call_results_on_several_devices(glue, ref, monkeypatch):
    """KSPANS_DEVICES with two or more entries: the .Call entry points run on a ks_mctx (here three shards on
    GPU 0) and must return what the reference's own glue returns on the same mock SEXPs"""
    monkeypatch.setenv("KSPANS_DEVICES", "0,0,0")
    rng = np.random.default_rng(2025)
    for trial in range(4):
        k = int(rng.choice([3, 6, 9]))
        seqs = [planted(rng, int(rng.integers(2000, 60000))) for _ in range(int(rng.integers(1, 6)))]
        if trial % 2:
            seqs.insert(1, b"AC"[: k - 1])
        a, b = ref.call_kmer_counts(seqs, k), glue.call_kmer_counts(seqs, k)
        assert a["n"] == b["n"] and (a["counts"] == b["counts"]).all()
        thr, mw, ms = float(rng.choice([0.5, 0.75])), int(rng.choice([0, 10, 40])), float(rng.choice([0, 3, 10]))
        a = ref.call_kmer_low_comp_regions(seqs, k, mw, ms, thr)
        b = glue.call_kmer_low_comp_regions(seqs, k, mw, ms, thr)
        assert (a["n"] == b["n"]).all() and (a["counts"] == b["counts"]).all()
        assert a["ranks"].tobytes() == b["ranks"].tobytes()
        assert a["pos"].tolist() == b["pos"].tolist()
        np.testing.assert_allclose(b["score"], a["score"], rtol=1e-9)
    # an argument error raised by the multi-device path keeps the reference's text
    with pytest.raises(Exception, match="threshold must be between 0 and 1"):
        glue.call_kmer_low_comp_regions(seqs, 4, 10, 1.0, 1.5)

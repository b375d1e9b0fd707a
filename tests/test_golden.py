"""Golden vectors recorded from the UNMODIFIED reference (tests/golden/make_golden.py, run in the build
container where /root/reference exists): the oracle (CPU tier) and the CUDA path (GPU tier) must reproduce
them.  Bars: counts, rank tables, coordinates, +-1 / integer scores bit-exact; rank-mode span scores 1e-9."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    z = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))
    names = bytes(z["names"]).decode().split("\n")
    out = []
    for n in names:
        g = {key[len(n) + 1:]: z[key] for key in z.files if key.startswith(n + "/")}
        g["seqs"] = bytes(g["seqs"]).split(b"\n")
        k, thr, mw, ms = g["params"]
        g.update(k=int(k), thr=float(thr), mw=int(mw), ms=float(ms), name=n)
        out.append(g)
    return out


GOLD = load()


@pytest.mark.parametrize("g", GOLD, ids=[g["name"] for g in GOLD])
def test_oracle_reproduces_reference_vectors(oracle, g):
    k, seqs = g["k"], g["seqs"]
    n, c = oracle.kmer_counts(seqs, k)
    assert n == g["counts_n"][0] and np.array_equal(c, g["counts"])
    r = oracle.low_comp(seqs, k, g["mw"], g["ms"], g["thr"])
    assert r["ranks"].tobytes() == g["lc_ranks"].tobytes()
    assert r["pos"].tolist() == g["lc_pos"].tolist() and r["score"].tobytes() == g["lc_score"].tobytes()
    r = oracle.kmer_regions(seqs, k, g["kr_W"], g["mw"], g["ms"])
    assert np.array_equal(r["counts"], g["kr_counts"])
    assert r["pos"].tolist() == g["kr_pos"].tolist() and r["score"].tobytes() == g["kr_score"].tobytes()
    r = oracle.tr_lr_regions(seqs, k, g["tr_init"], g["tr_trans"], min(g["mw"], 10))
    assert r["pos"].tolist() == g["tr_pos"].tolist() and r["score"].tobytes() == g["tr_score"].tobytes()
    if "wd_dist" in g:
        sel = bytes(g["wd_sel"]).split(b"\n")
        r = oracle.window_dist(seqs, sel, k, int(g["wd_window"][0]), True)
        assert np.array_equal(r["dist"], g["wd_dist"]) and np.array_equal(r["included"], g["wd_inc"])
        for q, p in enumerate(r["pos"]):
            assert (p is None) == ("wd_pos%d" % q not in g)
            if p is not None:
                assert np.array_equal(p, g["wd_pos%d" % q])


@pytest.mark.gpu
@pytest.mark.parametrize("g", GOLD, ids=[g["name"] for g in GOLD])
def test_cuda_path_reproduces_reference_vectors(g):
    from kmer_spans_b200 import api
    ctx = api.default_context()
    k, seqs = g["k"], g["seqs"]
    r = ctx.kmer_counts(seqs, k, with_f=False)
    assert r["n"][1] == g["counts_n"][0] and np.array_equal(r["counts"], g["counts"])
    r = ctx.kmer_low_comp_regions(seqs, k, g["mw"], g["ms"], g["thr"])
    assert r["w_rank"].tobytes() == g["lc_ranks"].tobytes()
    assert r["pos"].tolist() == g["lc_pos"].tolist()
    np.testing.assert_allclose(r["score"], g["lc_score"], rtol=1e-9, atol=0)
    r = ctx.kmer_regions(seqs, k, g["kr_W"], g["mw"], g["ms"])
    assert np.array_equal(r["counts"], g["kr_counts"])
    assert r["pos"].tolist() == g["kr_pos"].tolist() and r["score"].tobytes() == g["kr_score"].tobytes()
    kms = api.kmer_seq(k)
    r = ctx.lr_regions(seqs, (k, min(g["mw"], 10)), kms, g["tr_init"], g["tr_trans"])
    assert r["pos"].tolist() == g["tr_pos"].tolist() and r["score"].tobytes() == g["tr_score"].tobytes()
    if "wd_dist" in g:
        sel = bytes(g["wd_sel"]).split(b"\n")
        r = ctx.window_kmer_dist(seqs, sel, int(g["wd_window"][0]), freq=False, ret_flag=1)
        assert np.array_equal(r["dist"].T, g["wd_dist"]) and np.array_equal(r["seq_i"], g["wd_inc"])
        for q, p in enumerate(r["scores"]):
            assert (p is None) == ("wd_pos%d" % q not in g)
            if p is not None:
                assert np.array_equal(p.T, g["wd_pos%d" % q])

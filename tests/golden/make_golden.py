"""Generates tests/golden/ref_vectors.npz by running the UNMODIFIED reference (compiled into oracle/_ref from
/root/reference/src/kmer_spans.c against the mock R API) through its .Call entry points on small seeded
inputs.  Run in the build container:  python tests/golden/make_golden.py
The vectors let the parity tests run where the compiled reference is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ksoracle import Oracle, Ref  # noqa: E402
from tests.test_oracle import planted, rand_seq  # noqa: E402


def cases():
    rng = np.random.default_rng(20261018)
    out = []
    for i, (k, thr, mw, ms) in enumerate([(3, 0.75, 10, 2.0), (5, 0.6, 20, 3.0), (7, 0.75, 30, 5.0), (4, 0.5, 0, 0.0)]):
        seqs = [planted(rng, int(rng.integers(500, 6000))) for _ in range(3)] + [rand_seq(rng, 300, p_n=0.2), b"ACG"[: k - 1]]
        out.append(dict(name="case%d" % i, k=k, thr=thr, mw=mw, ms=ms, seqs=seqs))
    return out


def main():
    ref, orc = Ref(), Oracle()
    blob = {}
    names = []
    for c in cases():
        n, k, seqs = c["name"], c["k"], c["seqs"]
        names.append(n)
        blob[n + "/seqs"] = np.frombuffer(b"\n".join(seqs), np.uint8)
        blob[n + "/params"] = np.array([k, c["thr"], c["mw"], c["ms"]])
        r = ref.call_kmer_counts(seqs, k)
        blob[n + "/counts_n"] = np.array([r["n"]])
        blob[n + "/counts"] = r["counts"]
        r = ref.call_kmer_low_comp_regions(seqs, k, c["mw"], c["ms"], c["thr"])
        blob[n + "/lc_ranks"], blob[n + "/lc_pos"], blob[n + "/lc_score"] = r["ranks"], r["pos"], r["score"]
        W = np.where(np.arange(4 ** k) % 3 == 0, 1.0, -1.0)
        r = ref.call_kmer_regions_r(seqs, k, W, c["mw"], c["ms"])
        blob[n + "/kr_W"], blob[n + "/kr_counts"], blob[n + "/kr_pos"], blob[n + "/kr_score"] = W, r["counts"], r["pos"], r["score"]
        kms = [orc.kmer_seq(k, i).encode() for i in range(4 ** k)]
        init = ((np.arange(4 ** k) * 7) % 5 - 2).astype(float)
        trans = ((np.arange(4 ** k) * 11) % 6 - 3).astype(float)
        r = ref.call_tr_lr(seqs, k, min(c["mw"], 10), kms, init, trans)
        blob[n + "/tr_init"], blob[n + "/tr_trans"], blob[n + "/tr_pos"], blob[n + "/tr_score"] = init, trans, r["pos"], r["score"]
        if k <= 5:
            sel = [kms[1], kms[4 ** k - 2], kms[7 % 4 ** k]]
            window = 4 * k + 9
            r = ref.call_window_dist(seqs, sel, k, window, 1)
            blob[n + "/wd_sel"] = np.frombuffer(b"\n".join(sel), np.uint8)
            blob[n + "/wd_window"] = np.array([window])
            blob[n + "/wd_dist"], blob[n + "/wd_inc"] = r["dist"], r["included"]
            for q, p in enumerate(r["pos"]):
                if p is not None:
                    blob[n + "/wd_pos%d" % q] = np.asarray(p).reshape(len(sel), -1)
    blob["names"] = np.frombuffer("\n".join(names).encode(), np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

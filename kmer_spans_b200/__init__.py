"""kmer_spans_b200 -- B200-native (sm_100a) implementation of the kmer_spans hot path
(count -> score -> scan -> spans) behind the reference's own API.

    from kmer_spans_b200 import api
    ctx = api.Context()                         # needs a CUDA device; there is no CPU path
    r = ctx.kmer_low_comp_regions(seqs, k=12, min_w=100, min_score=20, thr=0.75)

`synth` (numpy only) generates the seeded synthetic genomes used by tests and bench.py.
"""
__all__ = ["api", "synth", "build"]

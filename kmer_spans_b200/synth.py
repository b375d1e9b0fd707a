"""Seeded synthetic genomes with planted tandem and interspersed repeats.

Pure numpy, deterministic for a given (n, seed): the same bytes feed the CPU checkers and the GPU.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

# human-like chromosome lengths in Mb (config 3)
HUMAN_MB = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80,
            59, 64, 47, 51, 156, 57]


def random_bases(rng, n):
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def genome(n, seed, tandem_every=100_000, tandem_len=2_000, inter_every=20_000, inter_len=300,
           n_block_every=1_000_000, n_block_len=500, n_block_offset=50_000, n_blocks=None,
           element=None):
    """One sequence of n bases as a uint8 array.

    every `tandem_every` bases: a `tandem_len` tandem array of a random 1-50 bp unit;
    every `inter_every` bases: a copy of one `inter_len` bp interspersed element;
    N blocks: `n_block_len` Ns every `n_block_every` (offset `n_block_offset`), or, when
    `n_blocks` = (count, length) is given, that many equally spaced blocks.
    """
    rng = np.random.default_rng(seed)
    seq = random_bases(rng, n)
    if element is None:
        element = random_bases(np.random.default_rng(seed ^ 0x5EED), inter_len)
    if inter_every:
        for p in range(inter_every // 2, n - len(element), inter_every):
            seq[p:p + len(element)] = element
    if tandem_every:
        for p in range(tandem_every // 3, n - tandem_len, tandem_every):
            unit = random_bases(rng, int(rng.integers(1, 51)))
            reps = -(-tandem_len // len(unit))
            seq[p:p + tandem_len] = np.tile(unit, reps)[:tandem_len]
    if n_blocks is not None:
        cnt, ln = n_blocks
        for j in range(cnt):
            p = (j + 1) * n // (cnt + 1)
            seq[p:p + ln] = ord("N")
    elif n_block_every:
        for p in range(n_block_offset, n - n_block_len, n_block_every):
            seq[p:p + n_block_len] = ord("N")
    return seq


def config1():
    """1 Mb, seed 1 (BASELINE.json configs[0])."""
    return [genome(1_000_000, 1)]


def config2(n=250_000_000, seed=2):
    """250 Mb chromosome-scale sequence with 5 centromere-like 50 kb N blocks (configs[1])."""
    return [genome(n, seed, n_blocks=(5, 50_000))]


def config3(scale=1.0):
    """24 human-like chromosomes (configs[2]); `scale` shrinks every length for tests."""
    element = random_bases(np.random.default_rng(0xE1E), 300)
    return [genome(int(mb * 1_000_000 * scale), 100 + i, element=element) for i, mb in enumerate(HUMAN_MB)]


def contigs(n_contigs, seed=5, lo=1_000, hi=50_000, k=10, tile=4096):
    """Short contigs (configs[4]): 10 % carry a tandem array straddling a multiple of `tile`,
    1 % end in a run shorter than k+1 behind an N block, a few have length exactly k and k+1."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_contigs):
        ln = int(rng.integers(lo, hi + 1))
        if i % 97 == 5:
            ln = k
        elif i % 97 == 6:
            ln = k + 1
        s = random_bases(rng, ln)
        u = rng.random()
        if u < 0.10 and ln > tile + 600:
            unit = random_bases(rng, int(rng.integers(1, 13)))
            p = tile - 250
            s[p:p + 500] = np.tile(unit, 500 // len(unit) + 1)[:500]
        elif u < 0.11 and ln > 3 * k + 40:
            s[ln - k - 20:ln - k] = ord("N")  # trailing run of exactly k
        elif u < 0.115 and ln > 3 * k + 40:
            s[ln - k + 1 - 20:ln - k + 1] = ord("N")  # trailing run of k-1
        out.append(s)
    return out

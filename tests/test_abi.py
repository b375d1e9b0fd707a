"""CPU tier: the C-ABI library loads, exports every symbol include/kspans.h declares, and refuses
to compute without a device (no CPU fallback).  No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kspans.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ks_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from kmer_spans_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libkspans_cuda.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from kmer_spans_b200 import api
    with pytest.raises(api.KspansError) as ei:
        api.Context()
    assert ei.value.code == 2 and "no CPU path" in str(ei.value)


def test_kmer_seq_host_helper():
    from kmer_spans_b200 import api
    assert api.kmer_seq(2) == "AA AC AT AG CA CC CT CG TA TC TT TG GA GC GT GG".split()  # kmer_spans.R:81-83
    assert api.kmer_seq(1) == ["A", "C", "T", "G"]


def test_product_does_not_touch_the_oracle():
    """the product package must not include, import, link or load anything under oracle/ or tests/"""
    pkg = os.path.join(ROOT, "kmer_spans_b200")
    pat = re.compile(r"^\s*(#\s*include|import\b|from\b)[^\n]*(oracle|ks_emu|tests[./]|_ref)|(CDLL|dlopen)[^\n]*(oracle|ks_emu|_ref)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                for ln in open(os.path.join(dirpath, f)):
                    code = ln.split("//")[0].split("#  ")[0]
                    assert not pat.search(code), "%s: %s" % (f, ln.strip())


def test_kmer_count_file_format(tmp_path):
    """kmers.to.file / read.kmers layout (kmer_spans.R:135-186): int32 LE magic 310572, n_k, sizes, tables"""
    import numpy as np
    from kmer_spans_b200 import api
    t2 = np.arange(16, dtype=np.int32)
    t3 = (np.arange(64, dtype=np.int32) * 7) % 11
    f = tmp_path / "x.bin"
    api.write_kmers(str(f), [t2, t3])
    raw = np.fromfile(str(f), "<i4")
    assert raw[:4].tolist() == [310572, 2, 16, 64] and raw.size == 4 + 16 + 64
    r = api.read_kmers(str(f))
    assert r["k"] == [2, 3] and (r["counts"][0] == t2).all() and (r["counts"][1] == t3).all()
    bad = tmp_path / "bad.bin"
    np.array([1, 2, 3], "<i4").tofile(str(bad))
    assert api.read_kmers(str(bad)) is False


def test_kmer_code_host_helper_matches_oracle(oracle):
    """ks_kmer_code (host only, no device): what init_kmer leaves in its offset for a k-mer string
    (src/kmer_spans.c:119-132 as used at :691,747), including strings with N and short strings"""
    import numpy as np
    from kmer_spans_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(77)
    for _ in range(2000):
        k = int(rng.integers(1, 16))
        n = int(rng.integers(0, k + 4))
        s = bytes(rng.choice(list(b"ACGTNnacgtR"), n).astype(np.uint8)) if n else b""
        assert lib.ks_kmer_code(s, k) == oracle.kmer_code(s, k), (s, k)
    for k in (1, 2, 5):
        for code in range(4 ** k):
            assert lib.ks_kmer_code(oracle.kmer_seq(k, code).encode(), k) == code


def test_fold_carry_host_helper():
    """ks_fold_carry (host only): the state / open excursion entering shard r from the 48-byte aggregates of
    the shards before it, against a plain Python fold of x -> kill ? b : max(x + a, b) and of the
    leftmost-maximum rule (strict >)."""
    import struct
    import numpy as np
    from kmer_spans_b200 import api

    def i128(v):  # two's complement, little endian (the layout of fx_t = __int128)
        return int(v % (1 << 128)).to_bytes(16, "little")

    def from_i128(b):
        v = int.from_bytes(b, "little")
        return v - (1 << 128) if v >> 127 else v

    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(1, 9))
        xf = [(int(rng.integers(-10**12, 10**12)), int(rng.integers(0, 10**12)), int(rng.random() < 0.2)) for _ in range(n)]
        blobs = [i128(a) + i128(b) + struct.pack("<I", kill) + b"\0" * 12 for a, b, kill in xf]
        S = 0
        for r in range(n + 1):
            got = from_i128(api.fold_carry(0, blobs, r)[:16])
            assert got == S, (trial, r)
            if r < n:
                a, b, kill = xf[r]
                S = b if kill else max(S + a, b)
        ex = []
        for _ in range(n):
            reset = int(rng.random() < 0.4)
            opened = 1 if not reset else int(rng.random() < 0.7)
            M = int(rng.integers(1, 10**12)) if opened else -(1 << 126)
            beg = int(rng.integers(16, 10**9)) if (reset and opened) else -1
            pk = int(rng.integers(16, 10**9)) if opened else -1
            ex.append((M, beg, pk, reset, opened))
        blobs = [i128(M) + struct.pack("<qqII", beg, pk, reset, opened) + b"\0" * 8 for M, beg, pk, reset, opened in ex]
        acc = (-(1 << 126), -1, -1, 1, 0)  # nothing is open left of the first shard
        for r in range(n + 1):
            out = api.fold_carry(1, blobs, r)
            M = from_i128(out[:16])
            beg, pk, reset, opened = struct.unpack("<qqII", out[16:40])
            assert (M, beg, pk, opened) == (acc[0], acc[1], acc[2], acc[4]) and reset == 1, (trial, r)
            if r < n:
                g = ex[r]
                if g[3]:
                    acc = g
                elif g[0] > acc[0]:
                    acc = (g[0], acc[1], g[2], acc[3], acc[4])

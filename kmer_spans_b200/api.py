"""Host-side mirror of the reference's R-facing API (kmer_spans.R) on top of the C ABI.

Function names, argument meaning, result fields and error behaviour follow
/root/reference/kmer_spans.R:18-27 (kmer.counts), :41-52 (kmer.regions), :72-79
(kmer.low.comp.regions) and :84-86 (kmer.seq), so that parity tests read like calls into the
reference.  R is absent from the build image; the real .Call glue is r/src/kmer_spans_glue.c.

Everything computes on the GPU through libkspans_cuda.so; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KsSpans, KspansError, MODE_LOG2, MODE_RANK, MODE_RANK_REL, MODE_SIGN  # noqa: F401


class SeqBatch:
    """Many sequences stored back to back in ONE uint8 array (lens[i] bytes each, no separators): the
    cheap way to hand 100k contigs to the C ABI from Python (pointers are computed vectorised)."""

    def __init__(self, buf, lens):
        self.buf = np.ascontiguousarray(buf, np.uint8)
        self.lens = np.ascontiguousarray(lens, np.int64)
        if int(self.lens.sum()) != self.buf.size:
            raise ValueError("SeqBatch: lengths do not add up to the buffer size")

    @classmethod
    def from_list(cls, seqs):
        seqs = [np.frombuffer(s, np.uint8) if not isinstance(s, np.ndarray) else s for s in seqs]
        return cls(np.concatenate(seqs) if seqs else np.zeros(0, np.uint8), [s.size for s in seqs])

    def __len__(self):
        return len(self.lens)


class SparseSeqs:
    """Sequences of which only some are present in host memory: lens[i] for every sequence, data {i: uint8 array}
    for the present ones.  For window uploads (one shard of a multi-GPU run): the layout needs every length, the
    copy touches only sequences inside the window -- a missing sequence inside it is an error of the caller."""

    def __init__(self, lens, data):
        self.lens = np.ascontiguousarray(lens, np.int64)
        self.data = {int(i): np.ascontiguousarray(a, np.uint8) for i, a in data.items()}
        for i, a in self.data.items():
            if a.size != self.lens[i]:
                raise ValueError("SparseSeqs: sequence %d has %d bytes, expected %d" % (i, a.size, self.lens[i]))

    def __len__(self):
        return len(self.lens)


def _as_bytes_list(seq):
    if isinstance(seq, (SeqBatch, SparseSeqs)):
        return seq
    if isinstance(seq, (bytes, bytearray, str, np.ndarray)):
        seq = [seq]
    out = []
    for s in seq:
        if isinstance(s, str):
            s = s.encode()
        elif isinstance(s, np.ndarray):
            s = np.ascontiguousarray(s, np.uint8)
        out.append(s)
    return out


class _SeqArgs:
    """(char**, int64*, n) view of a list of bytes / uint8 arrays, zero-copy"""

    def __init__(self, seqs):
        self.keep = seqs
        n = len(seqs)
        ptr = np.zeros(max(n, 1), np.uint64)
        ln = np.zeros(max(n, 1), np.int64)
        if isinstance(seqs, SparseSeqs):
            ln[:n] = seqs.lens
            for i, a in seqs.data.items():
                ptr[i] = a.ctypes.data
        elif isinstance(seqs, SeqBatch):
            ln[:n] = seqs.lens
            off = np.zeros(n, np.uint64)
            if n > 1:
                off[1:] = np.cumsum(seqs.lens[:-1]).astype(np.uint64)
            ptr[:n] = np.uint64(seqs.buf.ctypes.data) + off
        elif n and all(isinstance(s, np.ndarray) for s in seqs):
            # fast path for many contigs: no per-sequence ctypes objects
            ptr[:n] = np.fromiter((s.__array_interface__["data"][0] for s in seqs), np.uint64, n)
            ln[:n] = np.fromiter((s.size for s in seqs), np.int64, n)
        else:
            for i, s in enumerate(seqs):
                if isinstance(s, np.ndarray):
                    ptr[i] = s.ctypes.data
                    ln[i] = s.size
                else:
                    if not isinstance(s, bytes):
                        s = bytes(s)
                        seqs[i] = s  # keep the converted object alive
                    ptr[i] = C.cast(C.c_char_p(s), C.c_void_p).value or 0
                    ln[i] = len(s)
        self._ptr, self._len = ptr, ln
        self.ptrs = ptr.ctypes.data_as(C.POINTER(C.c_char_p))
        self.lens = ln.ctypes.data_as(C.POINTER(C.c_int64))
        self.n = n


def _spans_to_numpy(lib, sp):
    n = sp.n
    if n:
        pos = np.ctypeslib.as_array(sp.pos, shape=(n, 3)).copy()
        score = np.ctypeslib.as_array(sp.score, shape=(n, 2)).copy()
    else:
        pos = np.zeros((0, 3), np.int32)
        score = np.zeros((0, 2), np.float64)
    lib.ks_spans_free(C.byref(sp))
    return pos, score


class Context:
    """One GPU, one stream, cached scratch (ks_ctx)."""

    def __init__(self, device=-1):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.ks_ctx_create(C.byref(h), device)
        if rc:
            raise KspansError(rc, self.lib.ks_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "_owned", True):
                self.lib.ks_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise KspansError(rc, self.lib.ks_last_error(self.h).decode())

    @property
    def stream(self):
        return self.lib.ks_ctx_stream(self.h)

    def sync(self):
        self._ck(self.lib.ks_ctx_sync(self.h))

    def launches(self):
        return int(self.lib.ks_ctx_launches(self.h))

    def reset_launches(self):
        self.lib.ks_ctx_reset_launches(self.h)

    def scan_stats(self):
        lv, rv = C.c_int(0), C.c_uint64(0)
        self.lib.ks_ctx_scan_stats(self.h, C.byref(lv), C.byref(rv))
        return lv.value, rv.value

    def timer_start(self):
        self._ck(self.lib.ks_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self.lib.ks_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def side_table(self, on=True):
        """the score table of the log2 / +-1 stage is written next to the following scan (complete after it)"""
        self.lib.ks_ctx_side_table(self.h, 1 if on else 0)

    def set_profile(self, on=True):
        self.lib.ks_ctx_set_profile(self.h, 1 if on else 0)

    def profile(self, reset=False):
        """{class: (total ms, launches)} measured with CUDA event pairs on the launching stream"""
        names = ["count_kernel", "scan_level0", "scan_deeper", "scores_stage", "wfx_table"]
        out = {}
        for i, nm in enumerate(names):
            ms, n = C.c_double(0), C.c_uint64(0)
            self._ck(self.lib.ks_ctx_profile_get(self.h, i, C.byref(ms), C.byref(n)))
            out[nm] = (ms.value, int(n.value))
        if reset:
            self.lib.ks_ctx_profile_reset(self.h)
        return out

    # ---- device stages on a resident SeqSet (what bench.py times) ---------------------------
    def dev_pipeline(self, ss, k, mode, min_w, min_score, thr=0.0, param=float("nan"), d_counts=0, d_scores=0,
                     fetch_spans=False):
        """count -> scores(mode) -> scan with everything resident; d_counts / d_scores are device
        pointers (int32[4^k], double[4^k]) owned by the caller"""
        n = C.c_double(0)
        ns = C.c_uint64(0)
        sp = KsSpans()
        self._ck(self.lib.ks_dev_pipeline(self.h, ss.h, int(k), int(mode), float(param), float(thr), int(min_w),
                                          float(min_score), d_counts, d_scores, C.byref(n),
                                          C.byref(sp) if fetch_spans else None, C.byref(ns)))
        res = dict(n=n.value, n_spans=int(ns.value))
        if fetch_spans:
            res["pos"], res["score"] = _spans_to_numpy(self.lib, sp)
        return res

    def dev_count(self, ss, k, d_counts):
        n = C.c_double(0)
        self._ck(self.lib.ks_dev_count(self.h, ss.h, int(k), d_counts, C.byref(n)))
        return n.value

    def dev_count_async(self, ss, k, d_counts, d_nwords):
        """count without a host round trip; the word count is left at device pointer d_nwords (uint64)"""
        self._ck(self.lib.ks_dev_count_async(self.h, ss.h, int(k), C.c_void_p(d_counts), C.c_void_p(d_nwords)))

    def dev_xsum(self, peer_ptrs, rank, mc_ptr, n_u64):
        """sum of the count buffers of all ranks over peer memory (ks_dev_xsum); peer_ptrs: device addresses
        of every rank's buffer as mapped into this process"""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._ck(self.lib.ks_dev_xsum(self.h, arr, len(peer_ptrs), int(rank), C.c_void_p(int(mc_ptr)) if mc_ptr else None,
                                      int(n_u64)))

    def dev_scores_devtotal(self, k, d_counts, d_total, mode, d_scores, param=float("nan")):
        """scores with the total taken from device memory; returns the total"""
        t = C.c_double(0)
        self._ck(self.lib.ks_dev_scores_devtotal(self.h, int(k), C.c_void_p(d_counts), C.c_void_p(d_total), int(mode),
                                                 float(param), C.c_void_p(d_scores) if d_scores else None, C.byref(t)))
        return t.value

    def dev_scores(self, k, d_counts, total, mode, d_scores, param=float("nan")):
        self._ck(self.lib.ks_dev_scores(self.h, int(k), d_counts, float(total), int(mode), float(param), d_scores))

    def dev_scores_rank_sliced(self, k, d_counts, total, slice_, nslices, gather, d_scores):
        """rank-mode score stage for ONE slice of the k-mer index space (multi-GPU); gather(bytes) -> list of the
        blobs of all slices in slice order.  Fills d_scores and the rank-order positions for this slice only."""
        def cb(user, mine, nbytes, out):
            try:
                blobs = gather(C.string_at(mine, nbytes))
                for i, b in enumerate(blobs):
                    C.memmove(out + i * nbytes, b, nbytes)
                return 0
            except Exception:  # pragma: no cover - surfaced as a C-side error
                import traceback
                traceback.print_exc()
                return 1
        cfn = _lib.GATHER_FN(cb)
        self._ck(self.lib.ks_dev_scores_rank_sliced(self.h, int(k), C.c_void_p(d_counts), float(total), int(slice_),
                                                    int(nslices), C.cast(cfn, C.c_void_p), None, C.c_void_p(d_scores)))

    def rank_positions_ptr(self):
        return int(self.lib.ks_ctx_rank_positions(self.h) or 0)

    def dev_scan(self, ss, k, d_W, thr, min_w, min_score, d_inscan=0, fetch_spans=True):
        ns = C.c_uint64(0)
        sp = KsSpans()
        self._ck(self.lib.ks_dev_scan(self.h, ss.h, int(k), d_W, float(thr), int(min_w), float(min_score),
                                      d_inscan or None, C.byref(sp) if fetch_spans else None, C.byref(ns)))
        res = dict(n_spans=int(ns.value))
        if fetch_spans:
            res["pos"], res["score"] = _spans_to_numpy(self.lib, sp)
        return res

    def dev_scan_counts(self, ss, k, d_counts, thr, min_w, min_score, fetch_spans=True):
        """scan with score = f(count), f from the last dev_scores(mode LOG2 | SIGN)"""
        ns = C.c_uint64(0)
        sp = KsSpans()
        self._ck(self.lib.ks_dev_scan_counts(self.h, ss.h, int(k), d_counts, float(thr), int(min_w), float(min_score),
                                             C.byref(sp) if fetch_spans else None, C.byref(ns)))
        res = dict(n_spans=int(ns.value))
        if fetch_spans:
            res["pos"], res["score"] = _spans_to_numpy(self.lib, sp)
        return res

    def dev_scan_ranks(self, ss, k, thr, min_w, min_score, fetch_spans=True):
        """scan in rank mode with the rank order of the last dev_scores(mode RANK) on this context"""
        ns = C.c_uint64(0)
        sp = KsSpans()
        self._ck(self.lib.ks_dev_scan_ranks(self.h, ss.h, int(k), float(thr), int(min_w), float(min_score),
                                            C.byref(sp) if fetch_spans else None, C.byref(ns)))
        res = dict(n_spans=int(ns.value))
        if fetch_spans:
            res["pos"], res["score"] = _spans_to_numpy(self.lib, sp)
        return res

    # ---- one set sharded over several GPUs (exact stitching) -----------------------------------
    def dev_count_range(self, ss, k, chunk0, nchunks, d_counts):
        n = C.c_double(0)
        self._ck(self.lib.ks_dev_count_range(self.h, ss.h, int(k), int(chunk0), int(nchunks), d_counts, C.byref(n)))
        return n.value

    def dev_count_range_async(self, ss, k, chunk0, nchunks, d_counts, d_nwords):
        """the same without a host round trip; the word count is left at device pointer d_nwords (uint64)"""
        self._ck(self.lib.ks_dev_count_range_async(self.h, ss.h, int(k), int(chunk0), int(nchunks),
                                                   C.c_void_p(d_counts), C.c_void_p(d_nwords)))

    def dev_scan_shard(self, ss, k, table_ptr, thr, min_w, min_score, chunk0, nchunks, exchange, use_counts=False):
        """Level 0 restricted to dense chunks [chunk0, chunk0 + nchunks); `exchange(what, mine48) -> carry48`
        is called twice on the host (what = 0 transform, 1 open excursion) and must return what
        fold_carry() derives from the aggregates of all shards."""
        def cb(user, what, mine, carry):
            try:
                out = exchange(int(what), C.string_at(mine, 48))
                C.memmove(carry, out, 48)
                return 0
            except Exception:  # pragma: no cover - surfaced as a C-side error
                import traceback
                traceback.print_exc()
                return 1
        cfn = _lib.EXCHANGE_FN(cb)
        ns = C.c_uint64(0)
        sp = KsSpans()
        fn = self.lib.ks_dev_scan_counts_shard if use_counts else self.lib.ks_dev_scan_shard
        self._ck(fn(self.h, ss.h, int(k), table_ptr, float(thr), int(min_w), float(min_score), int(chunk0),
                    int(nchunks), C.cast(cfn, C.c_void_p), None, C.byref(sp), C.byref(ns)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(n_spans=int(ns.value), pos=pos, score=score)

    # ---- mirrors of the reference's R functions -------------------------------------------
    def kmer_counts(self, seq, k, with_f=True):
        """kmer.counts (kmer_spans.R:18-27): list(n = c(k, n), counts, f = counts / sum(counts))"""
        a = _SeqArgs(_as_bytes_list(seq))
        k = int(k)
        if not 1 <= k <= 15:
            raise KspansError(_lib.KS_ERR_ARG, "k must be a positive integer less than 1+MAX_K")
        counts = np.zeros(4 ** k, np.int32)
        n = C.c_double(0)
        self._ck(self.lib.ks_kmer_counts(self.h, a.ptrs, a.lens, a.n, k, counts.ctypes.data, C.byref(n)))
        out = dict(n=np.array([k, n.value]), counts=counts)
        if with_f:
            out["f"] = counts / counts.sum()
        return out

    def kmer_regions(self, seq, k, kmer_scores, min_width, min_score, counts=True):
        """kmer.regions (kmer_spans.R:41-52) -> kmer_regions_r.  kmer_scores is indexed by the
        2-bit code (the order of kmer.seq(k)); the name matching of the R wrapper is the caller's."""
        a = _SeqArgs(_as_bytes_list(seq))
        k = int(k)
        W = np.ascontiguousarray(kmer_scores, np.float64)
        if not 1 <= k <= 15:
            raise KspansError(_lib.KS_ERR_ARG, "kmer sizes larger than or equal to 16 not currently supported")
        if W.size != 4 ** k:
            raise KspansError(_lib.KS_ERR_ARG, "There should be a total of 4^k scores")
        cnt = np.zeros(4 ** k, np.int32)
        nuc = C.c_double(0)
        sp = KsSpans()
        self._ck(self.lib.ks_kmer_regions(self.h, a.ptrs, a.lens, a.n, k, W.ctypes.data, int(min_width),
                                          float(min_score), C.byref(nuc), cnt.ctypes.data if counts else None,
                                          C.byref(sp)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(n=nuc.value, counts=cnt, pos=pos, score=score)

    def kmer_low_comp_regions(self, seq, k, min_w, min_score, thr=0.75, want_tables=True):
        """kmer.low.comp.regions (kmer_spans.R:72-79): n, counts, w.rank, pos (R x 3), score (R x 2)"""
        a = _SeqArgs(_as_bytes_list(seq))
        k = int(k)
        if not 1 <= k <= 15:
            raise KspansError(_lib.KS_ERR_ARG, "k must be between 1 and 15")
        counts = np.zeros(4 ** k, np.int32) if want_tables else None
        ranks = np.zeros(4 ** k, np.float64) if want_tables else None
        n = (C.c_double * 2)()
        sp = KsSpans()
        self._ck(self.lib.ks_kmer_low_comp_regions(
            self.h, a.ptrs, a.lens, a.n, k, int(min_w), float(min_score), float(thr), n,
            counts.ctypes.data if want_tables else None, ranks.ctypes.data if want_tables else None, C.byref(sp)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(n=np.array([n[0], n[1]]), counts=counts, w_rank=ranks, pos=pos, score=score)

    def kmer_mode_regions(self, seq, k, mode, min_w, min_score, thr=0.0, param=float("nan"), want_tables=True,
                          counts_out=None, scores_out=None):
        """extension: counts -> scores(mode) -> scan on the device (README.md:27-49 modes).
        counts_out / scores_out: preallocated (e.g. pinned) numpy arrays to receive the tables."""
        a = _SeqArgs(_as_bytes_list(seq))
        k = int(k)
        if not 1 <= k <= 15:
            raise KspansError(_lib.KS_ERR_ARG, "k must be between 1 and 15")
        counts = counts_out if counts_out is not None else (np.zeros(4 ** k, np.int32) if want_tables else None)
        scores = scores_out if scores_out is not None else (np.zeros(4 ** k, np.float64) if want_tables else None)
        n = C.c_double(0)
        sp = KsSpans()
        self._ck(self.lib.ks_kmer_mode_regions(
            self.h, a.ptrs, a.lens, a.n, k, int(mode), float(param), float(thr), int(min_w), float(min_score),
            C.byref(n), counts.ctypes.data if counts is not None else None,
            scores.ctypes.data if scores is not None else None, C.byref(sp)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(n=n.value, counts=counts, scores=scores, pos=pos, score=score)

    def kmer_large_regions(self, seq, k, mode, min_w, min_score, thr=0.75, param=float("nan"), want_table=False):
        """large-k path (k up to 31; BASELINE.json configs[3]): hash-table counting, weighted rank over the k-mers
        that occur (mode 0) or +-1 around the frequency `param` (mode 2), scan.  want_table: also return the
        sparse table in (count, code) order (codes, counts, ranks)."""
        a = _SeqArgs(_as_bytes_list(seq))
        n, nd, sp = C.c_double(0), C.c_uint64(0), KsSpans()
        self._ck(self.lib.ks_kmer_large_regions(self.h, a.ptrs, a.lens, a.n, int(k), int(mode), float(param),
                                                float(thr), int(min_w), float(min_score), C.byref(n), C.byref(nd),
                                                C.byref(sp)))
        pos, score = _spans_to_numpy(self.lib, sp)
        out = dict(n=n.value, nd=int(nd.value), pos=pos, score=score)
        if want_table and out["nd"]:
            codes = np.zeros(out["nd"], np.uint64)
            counts = np.zeros(out["nd"], np.uint32)
            ranks = np.zeros(out["nd"], np.float64)
            self._ck(self.lib.ks_large_table(self.h, codes.ctypes.data, counts.ctypes.data, ranks.ctypes.data))
            out.update(codes=codes, counts=counts, ranks=ranks)
        return out

    def dev_large_regions(self, ss, k, mode, min_w, min_score, thr=0.75, param=float("nan"), fetch_spans=False):
        n, nd, ns, sp = C.c_double(0), C.c_uint64(0), C.c_uint64(0), KsSpans()
        self._ck(self.lib.ks_dev_large_regions(self.h, ss.h, int(k), int(mode), float(param), float(thr), int(min_w),
                                               float(min_score), C.byref(n), C.byref(nd),
                                               C.byref(sp) if fetch_spans else None, C.byref(ns)))
        res = dict(n=n.value, nd=int(nd.value), n_spans=int(ns.value))
        if fetch_spans:
            res["pos"], res["score"] = _spans_to_numpy(self.lib, sp)
        return res

    def kmer_scores(self, counts, k, total, mode=MODE_RANK, param=float("nan")):
        """rank_kmers_w (src/kmer_spans.c:189-202) / README modes as a table operator"""
        counts = np.ascontiguousarray(counts, np.int32)
        W = np.zeros(4 ** int(k), np.float64)
        self._ck(self.lib.ks_kmer_scores(self.h, int(k), counts.ctypes.data, float(total), int(mode), float(param),
                                         W.ctypes.data))
        return W

    def lr_regions(self, seq, params, kmers, kmer_scores, trans_scores):
        """lr.regions (kmer_spans.R:88-99) on tr_lr_regions_r (src/kmer_spans.c:649-713): params = (k,
        min_length); kmers = the 4^k k-mer strings in the order of the two score vectors.  Returns
        kmer_scores = the (4^k, 2) table re-ordered to 2-bit code order (init, trans), pos (n, 3) =
        seq.i, beg, end (all 1-based), score (n, 2)."""
        k, min_length = int(params[0]), int(params[1])
        n = 4 ** k
        kmers = [x.encode() if isinstance(x, str) else bytes(x) for x in kmers]
        ks_in = np.ascontiguousarray(kmer_scores, np.float64)
        tr_in = np.ascontiguousarray(trans_scores, np.float64)
        if not (1 <= k <= 15):
            raise KspansError(1, "k should be a positive value less than MAX_K")
        if len(kmers) != n or ks_in.size != n or tr_in.size != n:
            raise KspansError(1, "kmers_r, freq_a, freq_b should all be 4^k long")
        codes = np.fromiter((self.lib.ks_kmer_code(x, k) for x in kmers), np.int64, n)
        table = np.zeros((2, n), np.float64)
        table[0, codes] = ks_in  # later entries win, like the reference's loop (:689-696)
        table[1, codes] = tr_in
        a = _SeqArgs(_as_bytes_list(seq))
        sp = KsSpans()
        self._ck(self.lib.ks_tr_lr_regions(self.h, a.ptrs, a.lens, a.n, k, table[0].ctypes.data,
                                           table[1].ctypes.data, min_length, C.byref(sp)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(kmer_scores=table.T.copy(), pos=pos, score=score)

    def window_kmer_dist(self, seq, kmers, window, freq=True, ret_flag=0):
        """window.kmer.dist (kmer_spans.R:103-120) on windowed_kmer_count_distributions_r
        (src/kmer_spans.c:715-793): dist = (window+1) x len(kmers) occurrence histogram (column sums
        normalised when freq), seq_i = sequences longer than window, scores = per sequence the
        len x kmers matrix of window values when ret_flag & 1 (None for sequences left out)."""
        kmers = [x.encode() if isinstance(x, str) else bytes(x) for x in kmers]
        if len({len(x) for x in kmers}) != 1:
            raise ValueError("All kmers must be of the same size")
        k = len(kmers[0])
        seqs = _as_bytes_list(seq)
        a = _SeqArgs(seqs)
        codes = np.array([self.lib.ks_kmer_code(x, k) for x in kmers], np.uint32)
        window = int(window)
        dist = np.zeros((len(kmers), max(window, 0) + 1), np.int32)
        inc = np.zeros(a.n, np.int32)
        pos, pp = None, None
        if int(ret_flag) & 1:
            ln = a._len[:a.n]
            pos = [np.zeros((len(kmers), int(l)), np.int32) if l > window else None for l in ln]
            pp = (C.c_void_p * a.n)(*[p.ctypes.data if p is not None else None for p in pos])
        self._ck(self.lib.ks_windowed_kmer_count_distributions(
            self.h, a.ptrs, a.lens, a.n, k, codes.ctypes.data, len(kmers), window, dist.ctypes.data,
            inc.ctypes.data, pp))
        d = dist.T.copy()  # the R matrix: (window+1) x kmers
        if freq:
            # `dists$dist / colSums(dists$dist)` (kmer_spans.R:118): R recycles the divisor down the
            # columns, so element (r, c) is divided by colSums[(r + c * nrow) %% ncol] -- kept as is
            cs = d.sum(axis=0).astype(np.float64)
            flat = d.flatten("F").astype(np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                flat = flat / np.resize(cs, flat.size)
            d = flat.reshape(d.shape, order="F")
        return dict(dist=d, seq_i=inc, scores=[None if p is None else p.T for p in pos] if pos is not None else None)

    def dev_window_dist(self, ss, k, codes, window, d_dist, d_pos=0):
        codes = np.ascontiguousarray(codes, np.uint32)
        self._ck(self.lib.ks_dev_window_dist(self.h, ss.h, int(k), codes.ctypes.data, len(codes), int(window),
                                             C.c_void_p(d_dist), C.c_void_p(d_pos) if d_pos else None))

    # ---- device-resident -------------------------------------------------------------------
    def upload(self, seq):
        return SeqSet(self, _as_bytes_list(seq))

    def upload_window(self, seq, win_lo, win_hi):
        """only bytes [win_lo, win_hi) of the layout of these sequences (one shard, see plan_shard)"""
        return SeqSet(self, _as_bytes_list(seq), window=(int(win_lo), int(win_hi)))

    def dev_scan_ranks_shard(self, ss, k, thr, min_w, min_score, chunk0, nchunks, exchange):
        """rank-mode scan of one shard (rank order of the last dev_scores(mode RANK) on this context)"""
        def cb(user, what, mine, carry):
            try:
                out = exchange(int(what), C.string_at(mine, 48))
                C.memmove(carry, out, 48)
                return 0
            except Exception:  # pragma: no cover - surfaced as a C-side error
                import traceback
                traceback.print_exc()
                return 1
        cfn = _lib.EXCHANGE_FN(cb)
        ns = C.c_uint64(0)
        sp = KsSpans()
        self._ck(self.lib.ks_dev_scan_ranks_shard(self.h, ss.h, int(k), float(thr), int(min_w), float(min_score),
                                                  int(chunk0), int(nchunks), C.cast(cfn, C.c_void_p), None,
                                                  C.byref(sp), C.byref(ns)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(n_spans=int(ns.value), pos=pos, score=score)


class MultiContext:
    """Several GPUs behind one call (ks_mctx): one process, N devices, sharded upload, count tables summed over
    peer memory, scan carries folded exactly.  devices may repeat an index (several shards on one GPU)."""

    def __init__(self, devices):
        self.lib = _lib.load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.ks_mctx_create(C.byref(h), devs, len(devices))
        if rc:
            raise KspansError(rc, self.lib.ks_mctx_last_error(None).decode())
        self.h = h
        self.ndev = len(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ks_mctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise KspansError(rc, self.lib.ks_mctx_last_error(self.h).decode())

    def device_context(self, i):
        """borrowed view of the per-device context (profiling, launch counts); do not close it"""
        c = Context.__new__(Context)
        c.lib = self.lib
        c._owned = False
        c.h = C.c_void_p(self.lib.ks_mctx_ctx(self.h, int(i)))
        return c

    def kmer_counts(self, seq, k, with_f=True):
        a = _SeqArgs(_as_bytes_list(seq))
        k = int(k)
        counts = np.zeros(4 ** k if 1 <= k <= 15 else 1, np.int32)
        n = C.c_double(0)
        self._ck(self.lib.ks_m_kmer_counts(self.h, a.ptrs, a.lens, a.n, k, counts.ctypes.data, C.byref(n)))
        out = dict(n=np.array([k, n.value]), counts=counts)
        if with_f:
            out["f"] = counts / counts.sum()
        return out

    def kmer_mode_regions(self, seq, k, mode, min_w, min_score, thr=0.0, param=float("nan"), want_tables=True):
        a = _SeqArgs(_as_bytes_list(seq))
        k = int(k)
        ok = 1 <= k <= 15
        counts = np.zeros(4 ** k, np.int32) if want_tables and ok else None
        scores = np.zeros(4 ** k, np.float64) if want_tables and ok else None
        n = C.c_double(0)
        sp = KsSpans()
        self._ck(self.lib.ks_m_kmer_mode_regions(
            self.h, a.ptrs, a.lens, a.n, k, int(mode), float(param), float(thr), int(min_w), float(min_score),
            C.byref(n), counts.ctypes.data if counts is not None else None,
            scores.ctypes.data if scores is not None else None, C.byref(sp)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(n=n.value, counts=counts, scores=scores, pos=pos, score=score)

    def kmer_low_comp_regions(self, seq, k, min_w, min_score, thr=0.75, want_tables=True):
        a = _SeqArgs(_as_bytes_list(seq))
        k = int(k)
        ok = 1 <= k <= 15
        counts = np.zeros(4 ** k, np.int32) if want_tables and ok else None
        ranks = np.zeros(4 ** k, np.float64) if want_tables and ok else None
        n = (C.c_double * 2)()
        sp = KsSpans()
        self._ck(self.lib.ks_m_kmer_low_comp_regions(
            self.h, a.ptrs, a.lens, a.n, k, int(min_w), float(min_score), float(thr), n,
            counts.ctypes.data if counts is not None else None, ranks.ctypes.data if ranks is not None else None,
            C.byref(sp)))
        pos, score = _spans_to_numpy(self.lib, sp)
        return dict(n=np.array([n[0], n[1]]), counts=counts, w_rank=ranks, pos=pos, score=score)

    def load(self, seq):
        a = _SeqArgs(_as_bytes_list(seq))
        self._ck(self.lib.ks_m_load(self.h, a.ptrs, a.lens, a.n))

    def pipeline(self, k, mode, min_w, min_score, thr=0.0, param=float("nan"), fetch_spans=False, count_only=False):
        n, ns, sp = C.c_double(0), C.c_uint64(0), KsSpans()
        self._ck(self.lib.ks_m_pipeline(self.h, int(k), int(mode), float(param), float(thr), int(min_w),
                                        float(min_score), C.byref(n), C.byref(sp) if fetch_spans else None,
                                        C.byref(ns), 1 if count_only else 0))
        res = dict(n=n.value, n_spans=int(ns.value))
        if fetch_spans:
            res["pos"], res["score"] = _spans_to_numpy(self.lib, sp)
        return res

    def tables(self, i=0):
        """device pointers (int32 counts, double scores) of the tables device i holds after pipeline()"""
        c, s = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.ks_m_tables(self.h, int(i), C.byref(c), C.byref(s)))
        return c.value, s.value


class SeqSet:
    """Sequences resident in HBM (ks_seqset)."""

    def __init__(self, ctx, seqs, window=None):
        self.ctx = ctx
        a = _SeqArgs(seqs)
        h = C.c_void_p()
        if window is None:
            ctx._ck(ctx.lib.ks_seqset_upload(ctx.h, a.ptrs, a.lens, a.n, C.byref(h)))
        else:
            ctx._ck(ctx.lib.ks_seqset_upload_window(ctx.h, a.ptrs, a.lens, a.n, window[0], window[1], C.byref(h)))
        self.h = h
        self.bases = int(ctx.lib.ks_seqset_bases(h))
        self.chunks = int(ctx.lib.ks_seqset_chunks(h))
        self.buffer_bytes = int(ctx.lib.ks_seqset_buffer_bytes(h))
        self.positions = int(ctx.lib.ks_seqset_positions(h))
        self._keep = a

    def reupload(self, seqs, count_k=0, d_counts=0, d_nwords=0):
        """new sequences into the same device buffers; count_k > 0 counts behind the copies (asynchronous)"""
        ctx = self.ctx
        a = _SeqArgs(_as_bytes_list(seqs))
        ctx._ck(ctx.lib.ks_seqset_reupload(ctx.h, self.h, a.ptrs, a.lens, a.n, int(count_k),
                                           C.c_void_p(d_counts) if d_counts else None,
                                           C.c_void_p(d_nwords) if d_nwords else None))
        self._keep = a  # the copies may still be reading these buffers
        self.bases = int(ctx.lib.ks_seqset_bases(self.h))
        self.chunks = int(ctx.lib.ks_seqset_chunks(self.h))
        self.buffer_bytes = int(ctx.lib.ks_seqset_buffer_bytes(self.h))
        self.positions = int(ctx.lib.ks_seqset_positions(self.h))

    def start(self, seq):
        """buffer position of base 0 of sequence `seq`"""
        return int(self.ctx.lib.ks_seqset_start(self.h, int(seq)))

    def free(self):
        if getattr(self, "h", None):
            self.ctx.lib.ks_seqset_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def plan_shard(lens, nranks, rank):
    """(chunk0, nchunks, win_lo, win_hi) of shard `rank` of `nranks` for sequences of these lengths (ks_plan_shard)"""
    lens = np.ascontiguousarray(lens, np.int64)
    out = [C.c_int64(0) for _ in range(4)]
    rc = _lib.load().ks_plan_shard(lens.ctypes.data_as(C.POINTER(C.c_int64)), len(lens), int(nranks), int(rank),
                                   *[C.byref(o) for o in out])
    if rc:
        raise KspansError(rc, "ks_plan_shard: bad arguments")
    return tuple(int(o.value) for o in out)


def fold_carry(what, blobs, rank):
    """carry entering shard `rank` from the 48-byte aggregates of all shards (host, exact)"""
    lib = _lib.load()
    allb = b"".join(blobs)
    out = C.create_string_buffer(48)
    rc = lib.ks_fold_carry(int(what), allb, len(blobs), int(rank), out)
    if rc:
        raise KspansError(rc, "ks_fold_carry")
    return out.raw


def kmer_seq(k):
    """kmer.seq (kmer_spans.R:84-86): the 4^k k-mers in table order (A, C, T, G)"""
    lib = _lib.load()
    k = int(k)
    if not 1 <= k <= 16:
        raise KspansError(_lib.KS_ERR_ARG, "k_r (%d) should be smaller than MAX_K (16) and larger than 0" % k)
    buf = C.create_string_buffer(k + 1)
    out = []
    for i in range(4 ** k):
        lib.ks_kmer_seq(k, i, buf)
        out.append(buf.value.decode())
    return out


# ---- k-mer count files (kmer_spans.R:3-5,135-186): int32 LE -- magic, n_k, 4^k per k, the tables -----
KMER_MAGIC = 310572  # kmer.magic(), kmer_spans.R:5


def write_kmers(fname, tables):
    """the file body of kmers.to.file (kmer_spans.R:168-175) for a list of int32 count tables"""
    tables = [np.ascontiguousarray(t, "<i4") for t in tables]
    with open(fname, "wb") as f:
        np.array([KMER_MAGIC, len(tables)], "<i4").tofile(f)
        np.array([t.size for t in tables], "<i4").tofile(f)
        for t in tables:
            t.tofile(f)


def read_kmers(fname):
    """read.kmers (kmer_spans.R:178-186): dict(k=[...], counts=[...]) or False on a bad magic / count"""
    with open(fname, "rb") as f:
        head = np.fromfile(f, "<i4", 2)
        if head.size < 2 or head[0] != KMER_MAGIC or head[1] < 1:
            return False
        sizes = np.fromfile(f, "<i4", int(head[1]))
        counts = [np.fromfile(f, "<i4", int(n)) for n in sizes]
    return dict(k=[int(round(np.log2(n) / 2)) for n in sizes], counts=counts)


def kmers_to_file(ctx, seqs, out_prefix, ks, min_l=100_000):
    """kmers.to.file (kmer_spans.R:135-176) without the Biostrings FASTA reader: counts the sequences of
    length >= min_l for every k in `ks` on the GPU and writes <prefix>counts_<k1>_<k2>...bin"""
    seqs = _as_bytes_list(seqs)
    size = sum(len(s) for s in seqs)
    keep = [s for s in seqs if len(s) >= min_l]
    if not keep:
        raise ValueError("No sequence after length filtering")
    out = "%scounts_%s.bin" % (out_prefix, "_".join(str(int(k)) for k in ks))
    write_kmers(out, [ctx.kmer_counts(keep, int(k), with_f=False)["counts"] for k in ks])
    return dict(out=out, seq_size=size, seq_fsize=sum(len(s) for s in keep), seq_fl=len(keep))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def kmer_counts(seq, k, with_f=True):
    return default_context().kmer_counts(seq, k, with_f)


def kmer_regions(seq, k, kmer_scores, min_width, min_score):
    return default_context().kmer_regions(seq, k, kmer_scores, min_width, min_score)


def kmer_low_comp_regions(seq, k, min_w, min_score, thr=0.75):
    return default_context().kmer_low_comp_regions(seq, k, min_w, min_score, thr)

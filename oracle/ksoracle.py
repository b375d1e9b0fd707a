"""ctypes front-ends for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

Oracle  -- oracle/libks_oracle.so, this repo's C restatement (oracle/ks_oracle.c).
Ref     -- oracle/_ref/libkmer_spans_ref.so, the UNMODIFIED reference C file compiled against the
           mock R API (oracle/mockR); exposes both the core C functions and the .Call entries.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PAD = 64

RANK, LOG2, SIGN, RANK_REL = 0, 1, 2, 3


def build():
    """(Re)build both checkers; _ref only when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE], stdout=subprocess.DEVNULL)


def _as_bytes_list(seqs):
    if isinstance(seqs, (bytes, bytearray, str)):
        seqs = [seqs]
    return [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]


class _Spans(C.Structure):
    _fields_ = [("pos", C.POINTER(C.c_int32)), ("score", C.POINTER(C.c_double)),
                ("n", C.c_size_t), ("cap", C.c_size_t)]


def _padded(seqs):
    """zero-padded, NUL-terminated copies + pointer/len arrays"""
    bufs = [C.create_string_buffer(s, len(s) + PAD) for s in seqs]
    ptrs = (C.c_char_p * len(seqs))(*[C.cast(b, C.c_char_p) for b in bufs])
    lens = (C.c_int64 * len(seqs))(*[len(s) for s in seqs])
    return bufs, ptrs, lens


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "libks_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.kso_count_seq.restype = C.c_uint64
        L.kso_kmer_seq.argtypes = [C.c_int, C.c_uint64, C.c_char_p]

    @staticmethod
    def _spans_out(sp):
        n = sp.n
        pos = np.ctypeslib.as_array(sp.pos, shape=(n, 3)).copy() if n else np.zeros((0, 3), np.int32)
        sc = np.ctypeslib.as_array(sp.score, shape=(n, 2)).copy() if n else np.zeros((0, 2))
        return pos, sc

    def kmer_counts(self, seqs, k):
        seqs = _as_bytes_list(seqs)
        bufs, ptrs, lens = _padded(seqs)
        counts = np.zeros(4 ** k, np.int32)
        n = C.c_double(0)
        rc = self.lib.kso_kmer_counts(ptrs, lens, len(seqs), k, counts.ctypes.data_as(C.c_void_p), C.byref(n))
        if rc:
            raise ValueError("oracle kmer_counts rc=%d" % rc)
        return n.value, counts

    def rank(self, counts, k, total):
        counts = np.ascontiguousarray(counts, np.int32)
        ranks = np.zeros(4 ** k)
        self.lib.kso_rank(counts.ctypes.data_as(C.c_void_p), C.c_int(k), C.c_double(total),
                          ranks.ctypes.data_as(C.c_void_p))
        return ranks

    def rank_order(self, counts, k):
        counts = np.ascontiguousarray(counts, np.int32)
        order = np.zeros(4 ** k, np.uint32)
        self.lib.kso_rank_order(counts.ctypes.data_as(C.c_void_p), C.c_int(k), order.ctypes.data_as(C.c_void_p))
        return order

    def scores(self, counts, k, total, mode, param=float("nan")):
        counts = np.ascontiguousarray(counts, np.int32)
        W = np.zeros(4 ** k)
        rc = self.lib.kso_scores(counts.ctypes.data_as(C.c_void_p), C.c_int(k), C.c_double(total),
                                 C.c_int(mode), C.c_double(param), W.ctypes.data_as(C.c_void_p))
        if rc:
            raise ValueError("oracle scores rc=%d" % rc)
        return W

    def kmer_regions(self, seqs, k, W, min_width, min_score):
        seqs = _as_bytes_list(seqs)
        bufs, ptrs, lens = _padded(seqs)
        W = np.ascontiguousarray(W, np.float64)
        counts = np.zeros(4 ** k, np.int32)
        nuc = C.c_double(0)
        sp = _Spans()
        self.lib.kso_spans_init(C.byref(sp))
        rc = self.lib.kso_kmer_regions(ptrs, lens, len(seqs), k, W.ctypes.data_as(C.c_void_p),
                                       C.c_int(min_width), C.c_double(min_score), C.byref(nuc),
                                       counts.ctypes.data_as(C.c_void_p), C.byref(sp))
        pos, sc = self._spans_out(sp)
        self.lib.kso_spans_free(C.byref(sp))
        if rc:
            raise ValueError("oracle kmer_regions rc=%d" % rc)
        return dict(n=nuc.value, counts=counts, pos=pos, score=sc)

    def low_comp(self, seqs, k, min_width, min_score, thr=0.75):
        seqs = _as_bytes_list(seqs)
        bufs, ptrs, lens = _padded(seqs)
        counts = np.zeros(4 ** k, np.int32)
        ranks = np.zeros(4 ** k)
        n = (C.c_double * 2)()
        sp = _Spans()
        self.lib.kso_spans_init(C.byref(sp))
        rc = self.lib.kso_low_comp(ptrs, lens, len(seqs), k, C.c_int(min_width), C.c_double(min_score),
                                   C.c_double(thr), n, counts.ctypes.data_as(C.c_void_p),
                                   ranks.ctypes.data_as(C.c_void_p), C.byref(sp))
        pos, sc = self._spans_out(sp)
        self.lib.kso_spans_free(C.byref(sp))
        if rc:
            raise ValueError("oracle low_comp rc=%d" % rc)
        return dict(n=np.array([n[0], n[1]]), counts=counts, ranks=ranks, pos=pos, score=sc)

    def mode_regions(self, seqs, k, mode, min_width, min_score, thr=0.0, param=float("nan")):
        seqs = _as_bytes_list(seqs)
        bufs, ptrs, lens = _padded(seqs)
        counts = np.zeros(4 ** k, np.int32)
        W = np.zeros(4 ** k)
        n = C.c_double(0)
        sp = _Spans()
        self.lib.kso_spans_init(C.byref(sp))
        rc = self.lib.kso_mode_regions(ptrs, lens, len(seqs), k, C.c_int(mode), C.c_double(param),
                                       C.c_double(thr), C.c_int(min_width), C.c_double(min_score),
                                       C.byref(n), counts.ctypes.data_as(C.c_void_p),
                                       W.ctypes.data_as(C.c_void_p), C.byref(sp))
        pos, sc = self._spans_out(sp)
        self.lib.kso_spans_free(C.byref(sp))
        if rc:
            raise ValueError("oracle mode_regions rc=%d" % rc)
        return dict(n=n.value, counts=counts, scores=W, pos=pos, score=sc)

    def large_regions(self, seqs, k, mode, min_width, min_score, thr=0.75, param=float("nan")):
        """large-k checker (ks_oracle_large.c, parity by extension): sparse count table, rank over the k-mers that
        occur in (count, code) order, scan.  mode 0 rank, 2 +-1 around the frequency `param`."""
        class _Large(C.Structure):
            _fields_ = [("codes", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint32)),
                        ("ranks", C.POINTER(C.c_double)), ("nd", C.c_size_t), ("total", C.c_double)]
        seqs = _as_bytes_list(seqs)
        bufs, ptrs, lens = _padded(seqs)
        t = _Large()
        sp = _Spans()
        self.lib.kso_spans_init(C.byref(sp))
        rc = self.lib.kso_large_regions(ptrs, lens, len(seqs), C.c_int(k), C.c_int(mode), C.c_double(param),
                                        C.c_double(thr), C.c_int(min_width), C.c_double(min_score), C.byref(t),
                                        C.byref(sp))
        if rc:
            raise ValueError("oracle large_regions rc=%d" % rc)
        pos, sc = self._spans_out(sp)
        self.lib.kso_spans_free(C.byref(sp))
        nd = int(t.nd)
        out = dict(n=t.total, nd=nd, codes=np.ctypeslib.as_array(t.codes, (max(nd, 1),))[:nd].copy(),
                   counts=np.ctypeslib.as_array(t.counts, (max(nd, 1),))[:nd].copy(),
                   ranks=np.ctypeslib.as_array(t.ranks, (max(nd, 1),))[:nd].copy(), pos=pos, score=sc)
        self.lib.kso_large_free(C.byref(t))
        return out

    def kmer_seq(self, k, code):
        buf = C.create_string_buffer(k + 1)
        self.lib.kso_kmer_seq(k, code, buf)
        return buf.value.decode()

    # ---- SURVEY 8(f) rows 3 and 4 ----
    def kmer_code(self, kmer, k):
        kmer = kmer.encode() if isinstance(kmer, str) else bytes(kmer)
        self.lib.kso_kmer_code.restype = C.c_uint32
        return int(self.lib.kso_kmer_code(C.c_char_p(kmer), C.c_int(k)))

    def window_dist(self, seqs, kmers, k, window, want_pos=False):
        """kmers: strings of length k.  Returns dict(dist (kmer_n, window+1), included, pos)"""
        seqs = _as_bytes_list(seqs)
        bufs, ptrs, lens = _padded(seqs)
        codes = np.array([self.kmer_code(x, k) for x in kmers], np.uint32)
        dist = np.zeros((len(codes), window + 1), np.int32)
        inc = np.zeros(len(seqs), np.int32)
        pos = [np.zeros((len(codes), len(s)), np.int32) if len(s) > window else None for s in seqs] if want_pos else None
        pp = None
        if want_pos:
            pp = (C.c_void_p * len(seqs))(*[p.ctypes.data if p is not None else None for p in pos])
        rc = self.lib.kso_window_dist(ptrs, lens, len(seqs), C.c_int(k), codes.ctypes.data_as(C.c_void_p),
                                      C.c_int(len(codes)), C.c_int(window), dist.ctypes.data_as(C.c_void_p),
                                      inc.ctypes.data_as(C.c_void_p), pp)
        if rc:
            raise ValueError("oracle window_dist rc=%d" % rc)
        return dict(dist=dist, included=inc, pos=pos)

    def tr_lr_regions(self, seqs, k, init, trans, min_len):
        """init / trans in 2-bit code order"""
        seqs = _as_bytes_list(seqs)
        bufs, ptrs, lens = _padded(seqs)
        init = np.ascontiguousarray(init, np.float64)
        trans = np.ascontiguousarray(trans, np.float64)
        sp = _Spans()
        self.lib.kso_spans_init(C.byref(sp))
        rc = self.lib.kso_tr_lr_regions(ptrs, lens, len(seqs), C.c_int(k), init.ctypes.data_as(C.c_void_p),
                                        trans.ctypes.data_as(C.c_void_p), C.c_int(min_len), C.byref(sp))
        pos, sc = self._spans_out(sp)
        self.lib.kso_spans_free(C.byref(sp))
        if rc:
            raise ValueError("oracle tr_lr rc=%d" % rc)
        return dict(pos=pos, score=sc)


# ----------------------------------------------------------------------------------------------
class _SeqRegions(C.Structure):  # struct seq_regions, /root/reference/src/kmer_spans.c:46-58
    _fields_ = [("int_data", C.POINTER(C.c_int)), ("ints_nrow", C.c_size_t),
                ("doubles_nrow", C.c_size_t), ("double_data", C.POINTER(C.c_double)),
                ("n", C.c_size_t), ("capacity", C.c_size_t)]


class _Sexp(C.Structure):
    _fields_ = [("type", C.c_int), ("len", C.c_long), ("nrow", C.c_int), ("ncol", C.c_int),
                ("data", C.c_void_p)]


INTSXP, REALSXP, STRSXP, VECSXP = 13, 14, 16, 19


class Ref:
    """The compiled reference (default), or any other kmer_spans.so built against the mock R API
    (path=...: the replacement glue r/_build/kmer_spans.so), driven through mock SEXPs.
    Raises FileNotFoundError when the library was not built."""

    def __init__(self, path=None):
        self.is_reference = path is None
        if path is None:
            path = os.path.join(HERE, "_ref", "libkmer_spans_ref.so")
            if not os.path.exists(path) and os.path.exists("/root/reference/src/kmer_spans.c"):
                build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        P = C.POINTER(_Sexp)
        L.allocVector.restype = P
        L.allocVector.argtypes = [C.c_int, C.c_long]
        L.mkCharLen.restype = P
        L.mkCharLen.argtypes = [C.c_char_p, C.c_long]
        L.mockR_call.restype = P
        L.mockR_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(P)]
        L.mockR_last_error.restype = C.c_char_p
        self.P = P
        if not self.is_reference:
            return
        L.sequence_kmer_count.restype = C.c_size_t
        L.sequence_kmer_count.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.rank_kmers_w.restype = None
        L.rank_kmers_w.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double]
        L.init_regions.restype = _SeqRegions
        L.init_regions.argtypes = [C.c_size_t]
        L.kmer_regions.restype = None
        L.kmer_regions.argtypes = [C.c_char_p, C.c_int, C.c_size_t, C.c_size_t, C.c_double, C.c_void_p,
                                   C.c_int, C.c_double, C.POINTER(_SeqRegions), C.c_void_p]
        L.seq_regions_free.argtypes = [C.POINTER(_SeqRegions)]

    # ---- core C functions on plain buffers ----
    def sequence_kmer_count(self, seq, k, counts):
        buf = C.create_string_buffer(seq, len(seq) + PAD)
        return self.lib.sequence_kmer_count(C.cast(buf, C.c_char_p), counts.ctypes.data_as(C.c_void_p), k)

    def rank_kmers_w(self, counts, k, total):
        counts = np.ascontiguousarray(counts, np.int32)
        ranks = np.zeros(4 ** k)  # zero-filled: SURVEY T5
        self.lib.rank_kmers_w(counts.ctypes.data_as(C.c_void_p), k, ranks.ctypes.data_as(C.c_void_p), total)
        return ranks

    def kmer_regions_core(self, seqs, k, W, thr, min_width, min_score, count=False):
        seqs = _as_bytes_list(seqs)
        W = np.ascontiguousarray(W, np.float64)
        counts = np.zeros(4 ** k, np.int32) if count else None
        reg = self.lib.init_regions(100)
        for i, s in enumerate(seqs):
            if len(s) < k:
                continue
            buf = C.create_string_buffer(s, len(s) + PAD)
            self.lib.kmer_regions(C.cast(buf, C.c_char_p), i, 0, C.c_size_t(min_width & (2 ** 64 - 1)),
                                  min_score, W.ctypes.data_as(C.c_void_p), k, thr, C.byref(reg),
                                  counts.ctypes.data_as(C.c_void_p) if count else None)
        n = reg.n
        pos = np.ctypeslib.as_array(reg.int_data, shape=(n, 3)).copy() if n else np.zeros((0, 3), np.int32)
        sc = np.ctypeslib.as_array(reg.double_data, shape=(n, 2)).copy() if n else np.zeros((0, 2))
        self.lib.seq_regions_free(C.byref(reg))
        return pos.astype(np.int32), sc, counts

    # ---- .Call entry points through mock SEXPs ----
    def _strsxp(self, seqs):
        v = self.lib.allocVector(STRSXP, len(seqs))
        arr = C.cast(v.contents.data, C.POINTER(self.P))
        for i, s in enumerate(seqs):
            arr[i] = self.lib.mkCharLen(s, len(s))
        return v

    def _intsxp(self, vals):
        vals = np.atleast_1d(np.asarray(vals, np.int32))
        v = self.lib.allocVector(INTSXP, len(vals))
        C.memmove(v.contents.data, vals.ctypes.data, vals.nbytes)
        return v

    def _realsxp(self, vals):
        vals = np.atleast_1d(np.asarray(vals, np.float64))
        v = self.lib.allocVector(REALSXP, len(vals))
        C.memmove(v.contents.data, vals.ctypes.data, vals.nbytes)
        return v

    def _np(self, sexp):
        if not sexp:
            return None  # list element never set (R_NilValue)
        s = sexp.contents
        if s.type == VECSXP:
            return self._list(sexp)
        if s.type == INTSXP:
            a = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_int32)), shape=(max(s.len, 0),)).copy() if s.len else np.zeros(0, np.int32)
        elif s.type == REALSXP:
            a = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(max(s.len, 0),)).copy() if s.len else np.zeros(0)
        else:
            raise TypeError(s.type)
        if s.ncol != 1 or s.nrow != s.len:
            a = a.reshape(s.ncol, s.nrow)  # column-major R matrix nrow x ncol -> rows = columns
        return a

    def _call(self, name, args, lib=None):
        lib = lib or self.lib
        fn = C.cast(getattr(lib, name), C.c_void_p)
        arr = (self.P * len(args))(*args)
        r = self.lib.mockR_call(fn, len(args), arr)
        if not r:
            msg = self.lib.mockR_last_error().decode()
            self.lib.mockR_free_all()
            raise RuntimeError(msg)
        return r

    def _list(self, r):
        arr = C.cast(r.contents.data, C.POINTER(self.P))
        return [self._np(arr[i]) for i in range(r.contents.len)]

    def call_kmer_counts(self, seqs, k):
        seqs = _as_bytes_list(seqs)
        r = self._call("kmer_counts", [self._strsxp(seqs), self._intsxp(k)])
        out = self._list(r)
        self.lib.mockR_free_all()
        return dict(n=out[0][0], counts=out[1])

    def call_kmer_regions_r(self, seqs, k, W, min_width, min_score):
        seqs = _as_bytes_list(seqs)
        r = self._call("kmer_regions_r", [self._strsxp(seqs), self._intsxp(k), self._realsxp(W),
                                          self._intsxp(min_width), self._realsxp(min_score)])
        out = self._list(r)
        self.lib.mockR_free_all()
        return dict(n=out[0][0], counts=out[1], pos=out[2].reshape(-1, 3), score=out[3].reshape(-1, 2))

    def call_kmer_low_comp_regions(self, seqs, k, min_width, min_score, thr=0.75):
        seqs = _as_bytes_list(seqs)
        r = self._call("kmer_low_comp_regions", [self._strsxp(seqs), self._intsxp(k), self._intsxp(min_width),
                                                 self._realsxp(min_score), self._realsxp(thr)])
        out = self._list(r)
        self.lib.mockR_free_all()
        return dict(n=out[0], counts=out[1], ranks=out[2], pos=out[3].reshape(-1, 3), score=out[4].reshape(-1, 2))

    def registered(self):
        """[(name, arity)] as registered by R_init_kmer_spans"""
        class Def(C.Structure):
            _fields_ = [("name", C.c_char_p), ("fun", C.c_void_p), ("n", C.c_int)]
        self.lib.R_init_kmer_spans(None)
        self.lib.mockR_registered.restype = C.POINTER(Def)
        d = self.lib.mockR_registered()
        out, i = [], 0
        while d[i].name:
            out.append((d[i].name.decode(), d[i].n))
            i += 1
        return out

    def call_raw(self, name, args):
        """args: list of ('s', [bytes]) | ('i', ints) | ('d', doubles); returns list of numpy arrays"""
        sx = []
        for kind, val in args:
            sx.append(self._strsxp(_as_bytes_list(val)) if kind == "s" else
                      self._intsxp(val) if kind == "i" else self._realsxp(val))
        r = self._call(name, sx)
        out = self._list(r)
        self.lib.mockR_free_all()
        return out

    def call_window_dist(self, seqs, kmers, k, window, ret_flag=0):
        out = self.call_raw("windowed_kmer_count_distributions_r",
                            [("s", _as_bytes_list(seqs)), ("s", _as_bytes_list(kmers)), ("i", k), ("i", window),
                             ("i", ret_flag)])
        return dict(dist=out[0].reshape(-1, window + 1), included=out[1], pos=out[2])

    def call_tr_lr(self, seqs, k, min_len, kmers, kmer_scores, trans_scores):
        out = self.call_raw("tr_lr_regions_r", [("s", _as_bytes_list(seqs)), ("i", [k, min_len]),
                                                ("s", _as_bytes_list(kmers)), ("d", kmer_scores),
                                                ("d", trans_scores)])
        return dict(tables=out[0].reshape(2, -1), pos=out[1].reshape(-1, 3), score=out[2].reshape(-1, 2))

    def call_kmer_seq_r(self, k):
        r = self._call("kmer_seq_r", [self._intsxp(k)])
        arr = C.cast(r.contents.data, C.POINTER(self.P))
        out = [C.string_at(arr[i].contents.data).decode() for i in range(r.contents.len)]
        self.lib.mockR_free_all()
        return out

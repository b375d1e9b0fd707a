"""ctypes front-end of the host emulation (tests/emu/ks_emu.cpp) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(seqs):
    """concatenated buffer in the product's HBM layout (csrc/ks_layout.h)"""
    lens = np.array([len(s) for s in seqs], np.int64)
    starts = np.zeros(len(seqs) + 1, np.int64)
    cur = 16
    for i, l in enumerate(lens):
        starts[i] = cur
        cur += int(l) + 1
    starts[len(seqs)] = cur
    tot = ((cur + 15) // 16) * 16 + 16
    buf = np.zeros(tot + 64, np.uint8)
    for s, st in zip(seqs, starts):
        buf[st:st + len(s)] = np.frombuffer(bytes(s), np.uint8)
    return buf, tot, starts


class Emu:
    def __init__(self):
        subprocess.check_call(["make", "-s", "-C", HERE], stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(os.path.join(HERE, "libks_emu.so"))
        self.lib.emu_fx_roundtrip.restype = C.c_double
        self.lib.emu_fx_roundtrip.argtypes = [C.c_double, C.c_int]

    def count(self, seqs, k):
        buf, tot, starts = pack(seqs)
        counts = np.zeros(4 ** k, np.int32)
        n = C.c_uint64(0)
        self.lib.emu_count(buf.ctypes.data_as(C.c_void_p), C.c_int64(tot), C.c_int(k),
                           counts.ctypes.data_as(C.c_void_p), C.byref(n))
        return float(n.value), counts

    def count_pairs(self, seqs, k, row_cap=24):
        """the bucketed count (one sub-key per pair of k-mers, two tables per bucket, fold) on the host"""
        buf, tot, starts = pack(seqs)
        counts = np.zeros(4 ** k, np.int32)
        n = C.c_uint64(0)
        rc = self.lib.emu_count_pairs(buf.ctypes.data_as(C.c_void_p), C.c_int64(tot), C.c_int(k),
                                      counts.ctypes.data_as(C.c_void_p), C.byref(n), C.c_uint32(row_cap))
        if rc or n.value == 2 ** 64 - 1:
            raise ValueError("emu_count_pairs: rc=%d (geometry round trip failed)" % rc)
        return float(n.value), counts

    def rank(self, counts, k, total):
        counts = np.ascontiguousarray(counts, np.int32)
        ranks = np.zeros(4 ** k)
        self.lib.emu_rank_exact(counts.ctypes.data_as(C.c_void_p), C.c_int(k), C.c_double(total),
                                ranks.ctypes.data_as(C.c_void_p))
        return ranks

    def num_rank_segments(self, counts, k, total):
        counts = np.ascontiguousarray(counts, np.int32)
        return self.lib.emu_num_rank_segments(counts.ctypes.data_as(C.c_void_p), C.c_int(k), C.c_double(total))

    def scan_fast(self, seqs, k, W, thr, min_width, min_score):
        """the summary-based walk (scan_walk_fast_kernel / scan_detail_kernel) folded on the host; every
        summary-derived chunk element is checked against the position-by-position walk inside"""
        buf, tot, starts = pack(seqs)
        W = np.ascontiguousarray(W, np.float64)
        n = C.c_int64(0)
        pb, pp = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        ps = C.POINTER(C.c_double)()
        lv = C.c_int(0)
        nd = C.c_int64(0)
        rc = self.lib.emu_scan_fast(buf.ctypes.data_as(C.c_void_p), C.c_int64(tot), C.c_int(k),
                                    W.ctypes.data_as(C.c_void_p), C.c_double(thr),
                                    C.c_uint64(min_width & (2 ** 64 - 1)), C.c_double(min_score),
                                    C.byref(n), C.byref(pb), C.byref(pp), C.byref(ps), C.byref(lv), C.byref(nd))
        if rc:
            raise ValueError("emu_scan_fast rc=%d" % rc)
        m = n.value
        beg = np.ctypeslib.as_array(pb, shape=(m + 1,))[:m].copy()
        pk = np.ctypeslib.as_array(pp, shape=(m + 1,))[:m].copy()
        sc = np.ctypeslib.as_array(ps, shape=(m + 1,))[:m].copy()
        for p in (pb, pp, ps):
            self.lib.emu_free(p)
        sid = np.searchsorted(starts, beg, side="right") - 1
        pos = np.stack([sid, beg - starts[sid], pk - starts[sid]], 1).astype(np.int32) if m else np.zeros((0, 3), np.int32)
        score = np.stack([sc, np.zeros(m)], 1) if m else np.zeros((0, 2))
        return dict(pos=pos, score=score, levels=lv.value, detail_chunks=nd.value, chunks=tot // 16)

    def tr_scan(self, seqs, k, init, trans, min_len):
        """transition-score scan (chunk_walk_tr level loop) folded on the host; 1-based ids / coordinates"""
        buf, tot, starts = pack(seqs)
        init = np.ascontiguousarray(init, np.float64)
        trans = np.ascontiguousarray(trans, np.float64)
        n = C.c_int64(0)
        pb, pp = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        ps = C.POINTER(C.c_double)()
        lv = C.c_int(0)
        rc = self.lib.emu_tr_scan(buf.ctypes.data_as(C.c_void_p), C.c_int64(tot), C.c_int(k),
                                  init.ctypes.data_as(C.c_void_p), trans.ctypes.data_as(C.c_void_p), C.c_int(min_len),
                                  C.byref(n), C.byref(pb), C.byref(pp), C.byref(ps), C.byref(lv))
        if rc:
            raise ValueError("emu_tr_scan rc=%d" % rc)
        m = n.value
        beg = np.ctypeslib.as_array(pb, shape=(m + 1,))[:m].copy()
        pk = np.ctypeslib.as_array(pp, shape=(m + 1,))[:m].copy()
        sc = np.ctypeslib.as_array(ps, shape=(m + 1,))[:m].copy()
        for p in (pb, pp, ps):
            self.lib.emu_free(p)
        sid = np.searchsorted(starts, beg, side="right") - 1
        pos = np.stack([sid + 1, beg - starts[sid] + 1, pk - starts[sid] + 1], 1).astype(np.int32) if m else np.zeros((0, 3), np.int32)
        score = np.stack([sc, np.zeros(m)], 1) if m else np.zeros((0, 2))
        return dict(pos=pos, score=score, levels=lv.value)

    def scan(self, seqs, k, W, thr, min_width, min_score, inscan=False):
        buf, tot, starts = pack(seqs)
        W = np.ascontiguousarray(W, np.float64)
        cnt = np.zeros(4 ** k, np.int32) if inscan else None
        n = C.c_int64(0)
        pb, pp = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        ps = C.POINTER(C.c_double)()
        lv = C.c_int(0)
        rv = C.c_int64(0)
        rc = self.lib.emu_scan(buf.ctypes.data_as(C.c_void_p), C.c_int64(tot), C.c_int(k),
                               W.ctypes.data_as(C.c_void_p), C.c_double(thr),
                               C.c_uint64(min_width & (2 ** 64 - 1)), C.c_double(min_score),
                               cnt.ctypes.data_as(C.c_void_p) if inscan else None,
                               C.byref(n), C.byref(pb), C.byref(pp), C.byref(ps), C.byref(lv), C.byref(rv))
        if rc:
            raise ValueError("emu_scan rc=%d" % rc)
        m = n.value
        beg = np.ctypeslib.as_array(pb, shape=(m + 1,))[:m].copy()
        pk = np.ctypeslib.as_array(pp, shape=(m + 1,))[:m].copy()
        sc = np.ctypeslib.as_array(ps, shape=(m + 1,))[:m].copy()
        for p in (pb, pp, ps):
            self.lib.emu_free(p)
        sid = np.searchsorted(starts, beg, side="right") - 1
        pos = np.stack([sid, beg - starts[sid], pk - starts[sid]], 1).astype(np.int32) if m else np.zeros((0, 3), np.int32)
        score = np.stack([sc, np.zeros(m)], 1) if m else np.zeros((0, 2))
        return dict(pos=pos, score=score, counts=cnt, levels=lv.value, revisits=rv.value)

"""ctypes binding of the C ABI in include/kspans.h (kmer_spans_b200/csrc/libkspans_cuda.so).

The library is the product: hand-written sm_100a kernels, no CPU path.  Loading works anywhere
(the CPU test tier checks the exported symbols); creating a Context needs a CUDA device and
raises otherwise.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libkspans_cuda.so")

KS_OK, KS_ERR_ARG, KS_ERR_CUDA, KS_ERR_RANGE, KS_ERR_NOMEM = 0, 1, 2, 3, 4
MODE_RANK, MODE_LOG2, MODE_SIGN, MODE_RANK_REL = 0, 1, 2, 3


class KsSpans(C.Structure):
    _fields_ = [("pos", C.POINTER(C.c_int32)), ("score", C.POINTER(C.c_double)), ("n", C.c_size_t)]


class KspansError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kspans error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


_lib = None
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)
GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

_vp, _i, _d, _i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
_pd = C.POINTER(C.c_double)
_pu64 = C.POINTER(C.c_uint64)
_SEQS = [C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int]

SIGNATURES = {
    "ks_spans_free": (None, [C.POINTER(KsSpans)]),
    "ks_ctx_create": (_i, [C.POINTER(_vp), _i]),
    "ks_ctx_destroy": (None, [_vp]),
    "ks_last_error": (C.c_char_p, [_vp]),
    "ks_ctx_stream": (_vp, [_vp]),
    "ks_ctx_sync": (_i, [_vp]),
    "ks_ctx_launches": (C.c_uint64, [_vp]),
    "ks_ctx_reset_launches": (None, [_vp]),
    "ks_ctx_scan_stats": (None, [_vp, C.POINTER(_i), _pu64]),
    "ks_ctx_timer_start": (_i, [_vp]),
    "ks_ctx_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "ks_ctx_set_profile": (None, [_vp, _i]),
    "ks_ctx_side_table": (None, [_vp, _i]),
    "ks_ctx_profile_get": (_i, [_vp, _i, _pd, _pu64]),
    "ks_ctx_profile_reset": (None, [_vp]),
    "ks_kmer_counts": (_i, [_vp] + _SEQS + [_i, _vp, _pd]),
    "ks_kmer_regions": (_i, [_vp] + _SEQS + [_i, _vp, _i, _d, _pd, _vp, C.POINTER(KsSpans)]),
    "ks_kmer_low_comp_regions": (_i, [_vp] + _SEQS + [_i, _i, _d, _d, _pd, _vp, _vp, C.POINTER(KsSpans)]),
    "ks_kmer_mode_regions": (_i, [_vp] + _SEQS + [_i, _i, _d, _d, _i, _d, _pd, _vp, _vp, C.POINTER(KsSpans)]),
    "ks_kmer_scores": (_i, [_vp, _i, _vp, _d, _i, _d, _vp]),
    "ks_kmer_seq": (_i, [_i, C.c_uint64, C.c_char_p]),
    "ks_seqset_upload": (_i, [_vp] + _SEQS + [C.POINTER(_vp)]),
    "ks_seqset_wrap": (_i, [_vp, _vp, _i64, C.POINTER(C.c_int64), _i, C.POINTER(_vp)]),
    "ks_seqset_reupload": (_i, [_vp, _vp] + _SEQS + [_i, _vp, _vp]),
    "ks_seqset_free": (None, [_vp]),
    "ks_seqset_bases": (_i64, [_vp]),
    "ks_seqset_buffer_bytes": (_i64, [_vp]),
    "ks_dev_count": (_i, [_vp, _vp, _i, _vp, _pd]),
    "ks_dev_scores": (_i, [_vp, _i, _vp, _d, _i, _d, _vp]),
    "ks_dev_count_async": (_i, [_vp, _vp, _i, _vp, _vp]),
    "ks_dev_xsum": (_i, [_vp, _vp, _i, _i, _vp, C.c_uint64]),
    "ks_dev_scores_devtotal": (_i, [_vp, _i, _vp, _vp, _i, _d, _vp, _pd]),
    "ks_dev_scan": (_i, [_vp, _vp, _i, _vp, _d, _i, _d, _vp, C.POINTER(KsSpans), _pu64]),
    "ks_dev_scan_counts": (_i, [_vp, _vp, _i, _vp, _d, _i, _d, C.POINTER(KsSpans), _pu64]),
    "ks_dev_scan_ranks": (_i, [_vp, _vp, _i, _d, _i, _d, C.POINTER(KsSpans), _pu64]),
    "ks_dev_scan_ranks_shard": (_i, [_vp, _vp, _i, _d, _i, _d, _i64, _i64, _vp, _vp, C.POINTER(KsSpans), _pu64]),
    "ks_dev_scores_rank_sliced": (_i, [_vp, _i, _vp, _d, _i, _i, _vp, _vp, _vp]),
    "ks_ctx_rank_positions": (_vp, [_vp]),
    "ks_seqset_chunks": (_i64, [_vp]),
    "ks_plan_shard": (_i, [C.POINTER(C.c_int64), _i, _i, _i] + [C.POINTER(C.c_int64)] * 4),
    "ks_seqset_upload_window": (_i, [_vp] + _SEQS + [_i64, _i64, C.POINTER(_vp)]),
    "ks_dev_count_range": (_i, [_vp, _vp, _i, _i64, _i64, _vp, _pd]),
    "ks_dev_count_range_async": (_i, [_vp, _vp, _i, _i64, _i64, _vp, _vp]),
    "ks_dev_scan_shard": (_i, [_vp, _vp, _i, _vp, _d, _i, _d, _i64, _i64, _vp, _vp, C.POINTER(KsSpans), _pu64]),
    "ks_dev_scan_counts_shard": (_i, [_vp, _vp, _i, _vp, _d, _i, _d, _i64, _i64, _vp, _vp, C.POINTER(KsSpans), _pu64]),
    "ks_fold_carry": (_i, [_i, _vp, _i, _i, _vp]),
    "ks_dev_pipeline": (_i, [_vp, _vp, _i, _i, _d, _d, _i, _d, _vp, _vp, _pd, C.POINTER(KsSpans), _pu64]),
    "ks_kmer_code": (C.c_uint32, [C.c_char_p, _i]),
    "ks_tr_lr_regions": (_i, [_vp] + _SEQS + [_i, _vp, _vp, _i, C.POINTER(KsSpans)]),
    "ks_dev_tr_lr_regions": (_i, [_vp, _vp, _i, _vp, _vp, _i, C.POINTER(KsSpans), _pu64]),
    "ks_windowed_kmer_count_distributions": (_i, [_vp] + _SEQS + [_i, _vp, _i, _i, _vp, _vp, _vp]),
    "ks_dev_window_dist": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp]),
    "ks_dev_large_regions": (_i, [_vp, _vp, _i, _i, _d, _d, _i, _d, _pd, _pu64, C.POINTER(KsSpans), _pu64]),
    "ks_kmer_large_regions": (_i, [_vp] + _SEQS + [_i, _i, _d, _d, _i, _d, _pd, _pu64, C.POINTER(KsSpans)]),
    "ks_large_table": (_i, [_vp, _vp, _vp, _vp]),
    "ks_mctx_create": (_i, [C.POINTER(_vp), C.POINTER(_i), _i]),
    "ks_mctx_destroy": (None, [_vp]),
    "ks_mctx_last_error": (C.c_char_p, [_vp]),
    "ks_mctx_ndev": (_i, [_vp]),
    "ks_mctx_ctx": (_vp, [_vp, _i]),
    "ks_m_kmer_counts": (_i, [_vp] + _SEQS + [_i, _vp, _pd]),
    "ks_m_kmer_mode_regions": (_i, [_vp] + _SEQS + [_i, _i, _d, _d, _i, _d, _pd, _vp, _vp, C.POINTER(KsSpans)]),
    "ks_m_kmer_low_comp_regions": (_i, [_vp] + _SEQS + [_i, _i, _d, _d, _pd, _vp, _vp, C.POINTER(KsSpans)]),
    "ks_m_load": (_i, [_vp] + _SEQS),
    "ks_m_pipeline": (_i, [_vp, _i, _i, _d, _d, _i, _d, _pd, C.POINTER(KsSpans), _pu64, _i]),
    "ks_m_tables": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp)]),
    "ks_seqset_positions": (_i64, [_vp]),
    "ks_seqset_start": (_i64, [_vp, _i]),
}


def load():
    """Load the shared library (building it in-tree if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export the ABI
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

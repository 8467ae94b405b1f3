"""ctypes binding of libdbga_b200.so (the C ABI in include/dbga_b200.h).

There is deliberately no CPU fallback: if the CUDA library is missing or no B200 is
visible, importing succeeds (so host-side code and symbol checks run anywhere) but the
first compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdbga_b200.so")

OK, CYCLE, K_INVALID, BAD_CHAR, TOO_LARGE = 0, 1, 2, 3, 4
FLAG_NO_RUNS, FLAG_COMPACT_ROWS, FLAG_SOP = 1, 2, 4
ORDERS = ("XMY", "XYM", "MXY", "MYX", "YXM", "YMX")

# every symbol include/dbga_b200.h declares
EXPORTS = (
    "dbga_create", "dbga_destroy", "dbga_last_error", "dbga_set_stream", "dbga_set_scratch_limit",
    "dbga_host_alloc", "dbga_host_free", "dbga_align_batch", "dbga_plan", "dbga_plan_global",
    "dbga_plan_run", "dbga_plan_fetch", "dbga_measure_int_peak", "dbga_global_align_batch",
    "dbga_kmer_stats", "dbga_version",
    "dbga_align_batch_packed", "dbga_packed_words", "dbga_pack_ascii", "dbga_expand_rows",
    "dbga_multi_create", "dbga_multi_destroy", "dbga_multi_last_error", "dbga_align_batch_multi",
    "dbga_fasta_read", "dbga_fasta_free", "dbga_fasta_write",
)
MAX_SHARDS = 16


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("score", C.c_int32 * 16), ("gap_open", C.c_int32),
                ("gap_extend", C.c_int32), ("gap_len_mode", C.c_int32), ("xy_allowed", C.c_int32),
                ("pred_order", C.c_int32), ("end_order", C.c_int32), ("stages", C.c_int32),
                ("want_graph", C.c_int32), ("flags", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("num_pairs", C.c_int64),
                ("status", C.POINTER(C.c_int32)), ("score", C.POINTER(C.c_int64)),
                ("dp_cells", C.POINTER(C.c_int64)), ("dp_segments", C.POINTER(C.c_int32)),
                ("n_anchors", C.POINTER(C.c_int64)), ("run_off", C.POINTER(C.c_int64)),
                ("runs", C.POINTER(C.c_int32)), ("aln_off", C.POINTER(C.c_int64)),
                ("aln0", C.POINTER(C.c_uint8)), ("aln1", C.POINTER(C.c_uint8)),
                ("p_off", C.POINTER(C.c_int64)), ("nondup0", C.POINTER(C.c_uint8)),
                ("q_off", C.POINTER(C.c_int64)), ("nondup1", C.POINTER(C.c_uint8)),
                ("ms_h2d", C.c_float), ("ms_stage1", C.c_float), ("ms_stage2", C.c_float),
                ("ms_stage3", C.c_float), ("ms_assemble", C.c_float), ("ms_d2h", C.c_float),
                ("ms_total", C.c_float), ("dp_cells_total", C.c_int64),
                ("dp_segments_total", C.c_int64), ("kernel_launches", C.c_int64),
                ("sop", C.POINTER(C.c_int32)), ("seg_cols", C.POINTER(C.c_int32)), ("ms_plan", C.c_float)]


class MultiResult(C.Structure):
    _fields_ = [("num_shards", C.c_int32), ("pad", C.c_int32), ("pair_lo", C.c_int64 * (MAX_SHARDS + 1)),
                ("shard", Result * MAX_SHARDS), ("ms_total", C.c_float)]


class Fasta(C.Structure):
    _fields_ = [("num_records", C.c_int64), ("offsets", C.POINTER(C.c_int64)), ("seqs", C.POINTER(C.c_uint8)),
                ("packed", C.POINTER(C.c_uint32)), ("packed_words", C.c_int64), ("name_off", C.POINTER(C.c_int64)),
                ("names", C.POINTER(C.c_char)), ("bad_records", C.c_int64)]


_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load the CUDA library.  Raises LibraryMissing (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m dbga_b200.build` "
            "(dbga_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.dbga_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.dbga_create.restype = C.c_int
    lib.dbga_destroy.argtypes = [vp]
    lib.dbga_destroy.restype = None
    lib.dbga_last_error.argtypes = [vp]
    lib.dbga_last_error.restype = C.c_char_p
    lib.dbga_set_stream.argtypes = [vp, vp]
    lib.dbga_set_scratch_limit.argtypes = [vp, C.c_int64]
    lib.dbga_host_alloc.argtypes = [C.c_int64]
    lib.dbga_host_alloc.restype = vp
    lib.dbga_host_free.argtypes = [vp]
    lib.dbga_host_free.restype = None
    for name in ("dbga_align_batch", "dbga_global_align_batch"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, C.POINTER(C.c_int64), C.c_int64, C.POINTER(Params), C.POINTER(Result)]
        f.restype = C.c_int
    for name in ("dbga_plan", "dbga_plan_global"):
        f = getattr(lib, name)
        f.argtypes = [vp, C.POINTER(C.c_int64), C.c_int64, C.POINTER(Params)]
        f.restype = C.c_int
    lib.dbga_plan_run.argtypes = [vp, vp]
    lib.dbga_plan_run.restype = C.c_int
    lib.dbga_plan_fetch.argtypes = [vp, C.POINTER(Result)]
    lib.dbga_plan_fetch.restype = C.c_int
    lib.dbga_measure_int_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.dbga_measure_int_peak.restype = C.c_int
    lib.dbga_kmer_stats.argtypes = [vp, vp, C.POINTER(C.c_int64), C.c_int64, C.c_int32, C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.dbga_kmer_stats.restype = C.c_int
    lib.dbga_version.restype = C.c_int
    i64p, u32p = C.POINTER(C.c_int64), C.POINTER(C.c_uint32)
    lib.dbga_align_batch_packed.argtypes = [vp, vp, i64p, C.c_int64, C.POINTER(Params), C.POINTER(Result)]
    lib.dbga_align_batch_packed.restype = C.c_int
    lib.dbga_packed_words.argtypes = [i64p, C.c_int64]
    lib.dbga_packed_words.restype = C.c_int64
    lib.dbga_pack_ascii.argtypes = [vp, i64p, C.c_int64, vp, C.c_int]
    lib.dbga_pack_ascii.restype = C.c_int64
    lib.dbga_expand_rows.argtypes = [vp, i64p, C.POINTER(Result), vp, vp, C.c_int64, i64p, C.c_int]
    lib.dbga_expand_rows.restype = C.c_int
    lib.dbga_fasta_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.POINTER(Fasta))]
    lib.dbga_fasta_read.restype = C.c_int
    lib.dbga_fasta_free.argtypes = [C.POINTER(Fasta)]
    lib.dbga_fasta_free.restype = None
    lib.dbga_fasta_write.argtypes = [C.c_char_p, C.POINTER(Fasta), C.POINTER(Result), C.c_int64, C.c_int, C.c_int, C.c_int]
    lib.dbga_fasta_write.restype = C.c_int64
    lib.dbga_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.dbga_multi_create.restype = C.c_int
    lib.dbga_multi_destroy.argtypes = [vp]
    lib.dbga_multi_destroy.restype = None
    lib.dbga_multi_last_error.argtypes = [vp]
    lib.dbga_multi_last_error.restype = C.c_char_p
    lib.dbga_align_batch_multi.argtypes = [vp, vp, vp, i64p, C.c_int64, C.POINTER(Params), C.POINTER(MultiResult)]
    lib.dbga_align_batch_multi.restype = C.c_int
    _lib = lib
    return lib


def make_params(k, score16, d=10, e=2, gap_len_mode=0, xy_allowed=False, pred_order="XMY",
                end_order="XMY", stages=3, want_graph=False, flags=0) -> Params:
    p = Params()
    p.k = int(k)
    for i, v in enumerate(score16):
        p.score[i] = int(v)
    p.gap_open, p.gap_extend = int(d), int(e)
    p.gap_len_mode, p.xy_allowed = int(gap_len_mode), int(bool(xy_allowed))
    p.pred_order = ORDERS.index(pred_order) if isinstance(pred_order, str) else int(pred_order)
    p.end_order = ORDERS.index(end_order) if isinstance(end_order, str) else int(end_order)
    p.stages, p.want_graph = int(stages), int(bool(want_graph))
    p.flags = int(flags)
    return p


def as_array(ptr, n, dtype):
    """Copy n elements out of a ctypes pointer owned by the context."""
    if n <= 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)

"""ctypes binding of oracle/dbga_oracle.c (TEST INFRASTRUCTURE ONLY, see its header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import dbga_oracle as pyo

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


class Conv(C.Structure):
    _fields_ = [("gap_len_mode", C.c_int32), ("xy_allowed", C.c_int32),
                ("pred_order", C.c_int32), ("end_order", C.c_int32)]


class Pair(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_anchors", C.c_int64),
                ("anchors", C.POINTER(C.c_int32)), ("score", C.c_int64),
                ("aln_len", C.c_int64), ("aln0", C.POINTER(C.c_uint8)),
                ("aln1", C.POINTER(C.c_uint8)), ("dp_cells", C.c_int64),
                ("dp_segments", C.c_int64), ("id_count", C.c_int64),
                ("nondup0", C.POINTER(C.c_uint8)), ("nondup1", C.POINTER(C.c_uint8)),
                ("nk0", C.c_int64), ("nk1", C.c_int64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dbga_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_align_pair.restype = C.c_int32
        _lib.oracle_align_pair.argtypes = [
            C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int32,
            C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(Conv),
            C.c_int32, C.c_int32, C.POINTER(Pair)]
        _lib.oracle_free_pair.argtypes = [C.POINTER(Pair)]
        _lib.oracle_gotoh.restype = C.c_int32
        _lib.oracle_align_batch.restype = None
        _lib.oracle_check_batch.restype = None
    return _lib


def conv_struct(conv: pyo.Convention = pyo.DEFAULT) -> Conv:
    return Conv(conv.gap_len_mode, int(conv.xy_allowed),
                pyo.ORDERS.index(conv.pred_order), pyo.ORDERS.index(conv.end_order))


def score_array(match=10, transition=-1, transversion=-8, s=None):
    if s is None:
        s = pyo.make_dna_scoring_dict(match, transition, transversion)
    return (C.c_int32 * 16)(*pyo.score_table(s))


@dataclass
class CResult:
    status: int
    anchors: np.ndarray   # [M, 2] int32
    score: int
    aln0: str
    aln1: str
    dp_cells: int
    dp_segments: int
    id_count: int = 0
    nondup0: Optional[np.ndarray] = None
    nondup1: Optional[np.ndarray] = None


def align_pair(s0: str, s1: str, k: int, match=10, transition=-1, transversion=-8, d=10, e=2,
               conv: pyo.Convention = pyo.DEFAULT, s=None, want_graph=False, stages=3) -> CResult:
    L = lib()
    r = Pair()
    cv = conv_struct(conv)
    b0, b1 = s0.encode(), s1.encode()
    L.oracle_align_pair(b0, len(b0), b1, len(b1), k, score_array(match, transition, transversion, s),
                        d, e, C.byref(cv), int(want_graph), stages, C.byref(r))
    M = r.n_anchors
    anchors = (np.ctypeslib.as_array(r.anchors, shape=(2 * M,)).reshape(M, 2).copy()
               if M and r.anchors else np.zeros((0, 2), np.int32))
    a0 = C.string_at(r.aln0, r.aln_len).decode() if r.aln0 else ""
    a1 = C.string_at(r.aln1, r.aln_len).decode() if r.aln1 else ""
    nd0 = nd1 = None
    if want_graph and r.nondup0:
        nd0 = np.ctypeslib.as_array(r.nondup0, shape=(r.nk0,)).copy()
        nd1 = np.ctypeslib.as_array(r.nondup1, shape=(r.nk1,)).copy()
    out = CResult(r.status, anchors, r.score, a0, a1, r.dp_cells, r.dp_segments, r.id_count, nd0, nd1)
    L.oracle_free_pair(C.byref(r))
    return out


def gotoh(x: str, y: str, match=10, transition=-1, transversion=-8, d=10, e=2,
          conv: pyo.Convention = pyo.DEFAULT, s=None):
    """oracle_gotoh on two non-empty strings -> (score, aligned_x, aligned_y)."""
    L = lib()
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    xb = bytes(code[c] for c in x.encode())
    yb = bytes(code[c] for c in y.encode())
    ox = C.create_string_buffer(len(x) + len(y) + 1)
    oy = C.create_string_buffer(len(x) + len(y) + 1)
    n = C.c_int64(0)
    cv = conv_struct(conv)
    L.oracle_gotoh.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int32),
                               C.c_int32, C.c_int32, C.POINTER(Conv), C.c_char_p, C.c_char_p,
                               C.POINTER(C.c_int64)]
    sc = L.oracle_gotoh(xb, len(xb), yb, len(yb), score_array(match, transition, transversion, s),
                        d, e, C.byref(cv), ox, oy, C.byref(n))
    return int(sc), ox.raw[:n.value].decode(), oy.raw[:n.value].decode()


def align_batch(seqs: np.ndarray, offsets: np.ndarray, k: int, match=10, transition=-1,
                transversion=-8, d=10, e=2, conv: pyo.Convention = pyo.DEFAULT, s=None, threads: int = 0):
    """Summary-only batch (status, score, aligned length, DP cells, anchors per pair);
    fanned out over `threads` pthreads.  This is the timed CPU baseline."""
    L = lib()
    P = (len(offsets) - 1) // 2
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    status = np.zeros(P, np.int32)
    score = np.zeros(P, np.int64)
    alen = np.zeros(P, np.int64)
    cells = np.zeros(P, np.int64)
    nanch = np.zeros(P, np.int64)
    cv = conv_struct(conv)
    threads = threads or (os.cpu_count() or 1)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    L.oracle_align_batch(p(seqs, C.c_uint8), p(offsets, C.c_int64), C.c_int64(P), C.c_int32(k),
                         score_array(match, transition, transversion, s), C.c_int32(d), C.c_int32(e),
                         C.byref(cv), C.c_int32(threads), p(status, C.c_int32), p(score, C.c_int64), p(alen, C.c_int64),
                         p(cells, C.c_int64), p(nanch, C.c_int64))
    return dict(status=status, score=score, aln_len=alen, dp_cells=cells, n_anchors=nanch,
                threads=threads)


def check_batch(seqs: np.ndarray, offsets: np.ndarray, k: int, aln_off=None, aln0=None, aln1=None,
                run_off=None, runs=None, match=10, transition=-1, transversion=-8, d=10, e=2,
                conv: pyo.Convention = pyo.DEFAULT, s=None, threads: int = 0):
    """align_batch plus a pair-by-pair comparison, in C, of the gapped rows (aln_off / aln0 / aln1)
    and of the anchors (run_off / runs as (a, b, len) triples) that another implementation produced.
    Returns align_batch's columns and `diff` (bit 0: rows differ, bit 1: anchors differ)."""
    L = lib()
    P = (len(offsets) - 1) // 2
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    status = np.zeros(P, np.int32)
    score = np.zeros(P, np.int64)
    alen = np.zeros(P, np.int64)
    cells = np.zeros(P, np.int64)
    nanch = np.zeros(P, np.int64)
    diff = np.zeros(P, np.int32)
    cv = conv_struct(conv)
    threads = threads or (os.cpu_count() or 1)
    keep = []

    def p(a, t, dt):
        if a is None:
            return C.cast(None, C.POINTER(t))
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(t))

    L.oracle_check_batch(p(seqs, C.c_uint8, np.uint8), p(offsets, C.c_int64, np.int64), C.c_int64(P), C.c_int32(k),
                         score_array(match, transition, transversion, s), C.c_int32(d), C.c_int32(e),
                         C.byref(cv), C.c_int32(threads), p(status, C.c_int32, np.int32), p(score, C.c_int64, np.int64),
                         p(alen, C.c_int64, np.int64), p(cells, C.c_int64, np.int64), p(nanch, C.c_int64, np.int64),
                         p(aln_off, C.c_int64, np.int64), p(aln0, C.c_uint8, np.uint8), p(aln1, C.c_uint8, np.uint8),
                         p(run_off, C.c_int64, np.int64), p(runs, C.c_int32, np.int32), p(diff, C.c_int32, np.int32))
    return dict(status=status, score=score, aln_len=alen, dp_cells=cells, n_anchors=nanch, diff=diff,
                threads=threads)

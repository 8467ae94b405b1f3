"""Batch front-end: the reference's  [DeBruijnPairwise(p, k, "dna").alignment(...) for p in
pairs]  as one call into the CUDA library (src/dbga/debruijn_pairwise.py:42-76, 418-543)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import OK, CYCLE, K_INVALID, BAD_CHAR, TOO_LARGE  # noqa: F401

ALIGN = 16   # sequences start on 16-byte boundaries so that stage 1 uses 128-bit loads


def make_dna_scoring_dict(match: int, transition: int, transversion: int):
    """cogent3.align.make_dna_scoring_dict as called at debruijn_pairwise.py:448-450
    (transition = both purines A,G or both pyrimidines C,T)."""
    out = {}
    for a in "ACGT":
        for b in "ACGT":
            if a == b:
                out[a, b] = match
            elif (a in "AG") == (b in "AG"):
                out[a, b] = transition
            else:
                out[a, b] = transversion
    return out


def score_table(s) -> List[int]:
    return [int(s[a, b]) for a in "ACGT" for b in "ACGT"]


def pack_pairs(pairs: Sequence[Tuple[str, str]], align: int = ALIGN):
    """Concatenate pairs into (uint8 buffer, int64 offsets[2P+1]).  Because offsets are
    boundaries, alignment padding is appended to the *previous* sequence's slot only in the
    buffer, never in the offsets: each sequence gets its own [start, end) below."""
    P = len(pairs)
    starts = np.zeros(2 * P + 1, np.int64)
    lens = np.fromiter((len(x) for p in pairs for x in p), np.int64, 2 * P)
    # the C ABI defines sequence j as [offsets[j], offsets[j+1]); to keep that simple the
    # buffer is dense (no padding) unless every length is already a multiple of `align`.
    starts[1:] = np.cumsum(lens)
    buf = np.frombuffer("".join(x for p in pairs for x in p).encode("ascii"), np.uint8)
    return buf, starts


@dataclass
class BatchResult:
    status: np.ndarray
    score: np.ndarray
    dp_cells: np.ndarray
    dp_segments: np.ndarray
    n_anchors: np.ndarray
    run_off: np.ndarray
    runs: np.ndarray          # [R, 3] (a, b, len)
    aln_off: np.ndarray
    aln0: np.ndarray          # uint8
    aln1: np.ndarray
    p_off: Optional[np.ndarray] = None
    nondup0: Optional[np.ndarray] = None
    q_off: Optional[np.ndarray] = None
    nondup1: Optional[np.ndarray] = None
    sop: Optional[np.ndarray] = None       # [P, 4] match, mismatch, gap_open, gap_extend (utils.py:440-504)
    seg_cols: Optional[np.ndarray] = None  # FLAG_COMPACT_ROWS: columns per segment
    timing: dict = field(default_factory=dict)

    def __len__(self):
        return len(self.status)

    def alignment(self, i: int) -> Tuple[str, str]:
        lo, hi = int(self.aln_off[i]), int(self.aln_off[i + 1])
        return self.aln0[lo:hi].tobytes().decode(), self.aln1[lo:hi].tobytes().decode()

    def anchors(self, i: int) -> np.ndarray:
        """Merge nodes of pair i as (pos in seq1, pos in seq2), in seq-2 order."""
        r = self.runs[int(self.run_off[i]):int(self.run_off[i + 1])]
        if len(r) == 0:
            return np.zeros((0, 2), np.int32)
        reps = r[:, 2]
        base = np.repeat(r[:, :2], reps, axis=0)
        within = np.arange(int(reps.sum())) - np.repeat(np.cumsum(reps) - reps, reps)
        return (base + within[:, None]).astype(np.int32)

    def graph_flags(self, i: int):
        return (self.nondup0[int(self.p_off[i]):int(self.p_off[i + 1])],
                self.nondup1[int(self.q_off[i]):int(self.q_off[i + 1])])


class Aligner:
    """One CUDA context (device, stream, grow-only workspaces).  Not thread-safe; use one
    per GPU / host thread."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.dbga_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise RuntimeError(f"dbga_create(device={device}) failed (rc={rc}): no usable CUDA device; "
                               "dbga_b200 has no CPU fallback")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.dbga_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed (rc={rc}): {self.lib.dbga_last_error(self.h).decode()}")

    # ------------------------------------------------------------------ params
    @staticmethod
    def params(k, match=10, transition=-1, transversion=-8, d=10, e=2, s=None, **conv):
        if s is None:
            s = make_dna_scoring_dict(match, transition, transversion)
        return _lib.make_params(k, score_table(s), d, e, **conv)

    # ------------------------------------------------------------------ one-shot
    def _unpack(self, res: _lib.Result, want_graph: bool) -> BatchResult:
        P = int(res.num_pairs)
        aa = _lib.as_array
        run_off = aa(res.run_off, P + 1, np.int64)
        aln_off = aa(res.aln_off, P + 1, np.int64)
        nr = int(run_off[-1]) if P else 0
        na = int(aln_off[-1]) if P else 0
        out = BatchResult(
            status=aa(res.status, P, np.int32), score=aa(res.score, P, np.int64),
            dp_cells=aa(res.dp_cells, P, np.int64), dp_segments=aa(res.dp_segments, P, np.int32),
            n_anchors=aa(res.n_anchors, P, np.int64), run_off=run_off,
            runs=aa(res.runs, 3 * nr if res.runs else 0, np.int32).reshape(-1, 3), aln_off=aln_off,
            aln0=aa(res.aln0, na, np.uint8), aln1=aa(res.aln1, na, np.uint8))
        if res.sop:
            out.sop = aa(res.sop, 4 * P, np.int32).reshape(P, 4)
        if res.seg_cols:
            out.seg_cols = aa(res.seg_cols, nr + P, np.int32)
        if want_graph:
            out.p_off = aa(res.p_off, P + 1, np.int64)
            out.q_off = aa(res.q_off, P + 1, np.int64)
            out.nondup0 = aa(res.nondup0, int(out.p_off[-1]) if P else 0, np.uint8)
            out.nondup1 = aa(res.nondup1, int(out.q_off[-1]) if P else 0, np.uint8)
        out.timing = dict(ms_h2d=res.ms_h2d, ms_stage1=res.ms_stage1, ms_stage2=res.ms_stage2,
                          ms_stage3=res.ms_stage3, ms_assemble=res.ms_assemble, ms_d2h=res.ms_d2h,
                          ms_total=res.ms_total, ms_plan=res.ms_plan, dp_cells=int(res.dp_cells_total),
                          dp_segments=int(res.dp_segments_total), launches=int(res.kernel_launches))
        return out

    def align_buffers(self, seqs, offsets: np.ndarray, prm: _lib.Params,
                      global_only: bool = False, unpack: bool = True, packed: bool = False):
        """seqs: uint8 host buffer (numpy or pinned; or its address), offsets int64[2P+1].
        packed=True: `seqs` is the 2-bit packed form of the batch (pack_ascii), a uint32 buffer."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        P = (len(offsets) - 1) // 2
        res = _lib.Result()
        fn = (self.lib.dbga_align_batch_packed if packed else
              self.lib.dbga_global_align_batch if global_only else self.lib.dbga_align_batch)
        ptr = seqs if isinstance(seqs, int) else seqs.ctypes.data
        rc = fn(self.h, C.c_void_p(ptr), offsets.ctypes.data_as(C.POINTER(C.c_int64)), P, C.byref(prm), C.byref(res))
        self._check(rc, fn.__name__)
        self.last_raw = res
        return self._unpack(res, bool(prm.want_graph) and not global_only and not packed) if unpack else res

    def expand_rows(self, seqs, offsets: np.ndarray, raw: "_lib.Result", threads: int = 0):
        """Full gapped rows of a FLAG_COMPACT_ROWS result (dbga_expand_rows): -> (aln_off, row0, row1)."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        P = int(raw.num_pairs)
        cap = int(offsets[-1] - offsets[0]) + 16
        row0, row1 = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
        off = np.zeros(P + 1, np.int64)
        ptr = seqs if isinstance(seqs, int) else seqs.ctypes.data
        rc = self.lib.dbga_expand_rows(C.c_void_p(ptr), offsets.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(raw),
                                       C.c_void_p(row0.ctypes.data), C.c_void_p(row1.ctypes.data), cap,
                                       off.ctypes.data_as(C.POINTER(C.c_int64)), int(threads))
        if rc != 0:
            raise RuntimeError(f"dbga_expand_rows failed (rc={rc})")
        n = int(off[-1])
        return off, row0[:n], row1[:n]

    def align_pairs(self, pairs: Sequence[Tuple[str, str]], k: int, match=10, transition=-1,
                    transversion=-8, d=10, e=2, s=None, stages=3, want_graph=False, **conv) -> BatchResult:
        buf, offs = pack_pairs(pairs)
        prm = self.params(k, match, transition, transversion, d, e, s, stages=stages,
                          want_graph=want_graph, **conv)
        return self.align_buffers(buf, offs, prm)

    def global_align_pairs(self, pairs, match=10, transition=-1, transversion=-8, d=10, e=2, s=None,
                           **conv) -> BatchResult:
        """Batched dna_global_aln (utils.py:157-192): every pair is one whole segment."""
        buf, offs = pack_pairs(pairs)
        prm = self.params(1, match, transition, transversion, d, e, s, **conv)
        return self.align_buffers(buf, offs, prm, global_only=True)

    # ------------------------------------------------------------------ split form
    def plan(self, offsets: np.ndarray, prm: _lib.Params, global_only: bool = False):
        offsets = np.ascontiguousarray(offsets, np.int64)
        P = (len(offsets) - 1) // 2
        fn = self.lib.dbga_plan_global if global_only else self.lib.dbga_plan
        self._check(fn(self.h, offsets.ctypes.data_as(C.POINTER(C.c_int64)), P, C.byref(prm)), fn.__name__)
        self._plan_graph = bool(prm.want_graph) and not global_only

    def run(self, device_ptr: int):
        self._check(self.lib.dbga_plan_run(self.h, C.c_void_p(int(device_ptr))), "dbga_plan_run")

    def fetch(self, unpack: bool = True):
        res = _lib.Result()
        self._check(self.lib.dbga_plan_fetch(self.h, C.byref(res)), "dbga_plan_fetch")
        self.last_raw = res
        return self._unpack(res, self._plan_graph) if unpack else res

    def kmer_stats(self, pairs: Sequence[Tuple[str, str]], k: int):
        """|K0|, |K1|, |K0 & K1| of the k-mer sets of every pair (what predict_p and the k-size
        heuristics are computed from, utils.py:371-413, kmersize.py:10-14).
        Returns (status, distinct0, distinct1, shared) as numpy arrays."""
        buf, offs = pack_pairs(pairs)
        P = len(pairs)
        status = np.zeros(P, np.int32)
        d0, d1, sh = np.zeros(P, np.int64), np.zeros(P, np.int64), np.zeros(P, np.int64)
        i64 = C.POINTER(C.c_int64)
        rc = self.lib.dbga_kmer_stats(self.h, C.c_void_p(buf.ctypes.data), offs.ctypes.data_as(i64), P, int(k),
                                      status.ctypes.data_as(C.POINTER(C.c_int32)), d0.ctypes.data_as(i64),
                                      d1.ctypes.data_as(i64), sh.ctypes.data_as(i64))
        self._check(rc, "dbga_kmer_stats")
        return status, d0, d1, sh

    def set_stream(self, stream_ptr: int):
        self._check(self.lib.dbga_set_stream(self.h, C.c_void_p(int(stream_ptr))), "dbga_set_stream")

    def set_scratch_limit(self, nbytes: int):
        self._check(self.lib.dbga_set_scratch_limit(self.h, int(nbytes)), "dbga_set_scratch_limit")

    def int_peak(self):
        out = (C.c_double * 3)()
        self._check(self.lib.dbga_measure_int_peak(self.h, out), "dbga_measure_int_peak")
        return dict(add_ops_per_s=out[0], max_ops_per_s=out[1], mix_ops_per_s=out[2])


def pack_ascii(buf: np.ndarray, offsets: np.ndarray, threads: int = 0) -> np.ndarray:
    """2-bit packed form of a batch in the C ABI's canonical layout (dbga_pack_ascii): sequence j
    starts at word (offsets[j] >> 4) + j.  Raises ValueError on a character outside ACGT."""
    lib = _lib.load()
    offsets = np.ascontiguousarray(offsets, np.int64)
    nseq = len(offsets) - 1
    i64 = C.POINTER(C.c_int64)
    words = int(lib.dbga_packed_words(offsets.ctypes.data_as(i64), nseq))
    out = np.zeros(words, np.uint32)
    buf = np.ascontiguousarray(buf, np.uint8)
    bad = lib.dbga_pack_ascii(C.c_void_p(buf.ctypes.data), offsets.ctypes.data_as(i64), nseq, C.c_void_p(out.ctypes.data), int(threads))
    if bad != 0:
        raise ValueError(f"{bad} sequence(s) hold characters outside ACGT")
    return out


class MultiAligner:
    """One host process, several GPUs (dbga_align_batch_multi): the batch is cut into contiguous
    shards of equal bytes, one per device, each driven by its own host thread; no collective."""

    def __init__(self, devices: Sequence[int]):
        self.lib = _lib.load()
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        rc = self.lib.dbga_multi_create(arr, len(self.devices), C.byref(h))
        if rc != 0 or not h:
            raise RuntimeError(f"dbga_multi_create({self.devices}) failed (rc={rc}); dbga_b200 has no CPU fallback")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.dbga_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align_buffers(self, seqs, offsets: np.ndarray, prm: _lib.Params, packed=None, unpack: bool = True):
        """-> list of (pair_lo, BatchResult) per shard (unpack=False: the raw MultiResult)."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        P = (len(offsets) - 1) // 2
        res = _lib.MultiResult()
        sp = None if seqs is None else (seqs if isinstance(seqs, int) else seqs.ctypes.data)
        pp = None if packed is None else (packed if isinstance(packed, int) else packed.ctypes.data)
        rc = self.lib.dbga_align_batch_multi(self.h, C.c_void_p(sp), C.c_void_p(pp), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                             P, C.byref(prm), C.byref(res))
        if rc != 0:
            raise RuntimeError(f"dbga_align_batch_multi failed (rc={rc}): {self.lib.dbga_multi_last_error(self.h).decode()}")
        self.last_raw = res
        if not unpack:
            return res
        un = Aligner._unpack
        return [(int(res.pair_lo[d]), un(None, res.shard[d], False)) for d in range(res.num_shards)]


_default: dict = {}


def default_aligner(device: int = 0) -> Aligner:
    if device not in _default:
        _default[device] = Aligner(device)
    return _default[device]


def align_batch(pairs, k, device: int = 0, **kw) -> BatchResult:
    """Addition to the reference API: align many independent pairs in one call."""
    return default_aligner(device).align_pairs(pairs, k, **kw)

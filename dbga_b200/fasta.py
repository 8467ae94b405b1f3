"""Batched FASTA in / out around the batch API: the reference reads one
pair per file through cogent3 (src/dbga/utils.py:95-96) and writes `aln.to_fasta()`
(src/dbga/cli.py:135-136); a batch run wants many pairs per file, parsed straight into the
C-ABI layout (one uint8 buffer + offsets) without per-record Python objects."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def read_fasta_buffer(path) -> Tuple[List[str], np.ndarray, np.ndarray]:
    """Parse a FASTA file into (names, buf uint8, offsets int64[R+1]); record r is
    buf[offsets[r]:offsets[r+1]], upper-cased, line breaks removed.  Vectorised: no Python
    loop over sequence bytes."""
    data = np.fromfile(path, dtype=np.uint8)
    if data.size == 0:
        return [], np.zeros(0, np.uint8), np.zeros(1, np.int64)
    nl = data == 10
    line_start = np.empty(data.size, bool)
    line_start[0] = True
    line_start[1:] = nl[:-1]
    hdr_start = line_start & (data == ord(">"))
    # a byte belongs to a header line iff the most recent line start is a header start
    line_id = np.cumsum(line_start) - 1
    hdr_line = hdr_start[line_start]                      # per line: is it a header
    in_hdr = hdr_line[line_id]
    rec_id = np.cumsum(hdr_start) - 1                     # record index of every byte
    keep = ~in_hdr & ~nl & (data != 13) & (data != 32) & (rec_id >= 0)
    buf = data[keep].copy()
    lower = (buf >= 97) & (buf <= 122)
    buf[lower] -= 32
    n_rec = int(hdr_start.sum())
    counts = np.bincount(rec_id[keep], minlength=n_rec).astype(np.int64)
    offsets = np.zeros(n_rec + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    starts = np.flatnonzero(hdr_start)
    ends = np.flatnonzero(nl)
    names = []
    for st in starts:
        j = np.searchsorted(ends, st)
        e = ends[j] if j < len(ends) else data.size
        head = data[st + 1:e].tobytes().decode("ascii", "replace").strip()
        names.append(head.split()[0] if head else "")
    return names, buf, offsets


def read_pairs(path):
    """FASTA with records (a1, b1, a2, b2, ...) -> (names, buf, offsets[2P+1]) ready for
    `Aligner.align_buffers`."""
    names, buf, offsets = read_fasta_buffer(path)
    if len(names) % 2:
        raise ValueError("a pair file needs an even number of FASTA records")
    return names, buf, offsets


def format_record(name, row: bytes, width: int = 60) -> bytes:
    body = b"\n".join(row[i:i + width] for i in range(0, len(row), width)) if row else b""
    return b">" + str(name).encode() + b"\n" + body + b"\n"


def write_alignments(path, names: Sequence, result, width: int = 60, only_ok: bool = True) -> int:
    """Write every aligned pair of a BatchResult as two 60-column-wrapped FASTA records
    (cogent3's to_fasta block size).  Returns the number of pairs written."""
    n = 0
    with open(path, "wb") as f:
        for i in range(len(result)):
            if only_ok and result.status[i] != 0:
                continue
            lo, hi = int(result.aln_off[i]), int(result.aln_off[i + 1])
            f.write(format_record(names[2 * i], result.aln0[lo:hi].tobytes(), width))
            f.write(format_record(names[2 * i + 1], result.aln1[lo:hi].tobytes(), width))
            n += 1
    return n


# ---------------------------------------------------------------------------------------------
# Native path (csrc/fasta_io.cu behind dbga_fasta_read / dbga_fasta_write): the parser fills a pinned
# buffer in the C ABI's batch layout -- and the 2-bit packed words dbga_align_batch_packed uploads --
# in parallel over host threads; the writer wraps the rows of a result at 60 columns, expanding a
# COMPACT_ROWS result on the fly.  No Python object per record on either side.
class FastaBatch:
    """A FASTA file parsed by dbga_fasta_read.  Records 2i and 2i+1 are pair i."""

    def __init__(self, path, want_packed: bool = True, threads: int = 0):
        import ctypes as C
        from . import _lib
        self._lib = _lib.load()
        self._h = C.POINTER(_lib.Fasta)()
        rc = self._lib.dbga_fasta_read(str(path).encode(), int(want_packed), int(threads), C.byref(self._h))
        if rc != 0:
            raise ValueError(f"could not read FASTA file {path} (rc={rc})")
        f = self._h.contents
        self.num_records = int(f.num_records)
        self.offsets = np.ctypeslib.as_array(f.offsets, shape=(self.num_records + 1,)) if self.num_records else np.zeros(1, np.int64)
        self.seqs_ptr = C.cast(f.seqs, C.c_void_p).value
        self.packed_ptr = C.cast(f.packed, C.c_void_p).value if f.packed else None
        self.bad_records = int(f.bad_records)

    @property
    def num_pairs(self) -> int:
        return self.num_records // 2

    @property
    def seqs(self) -> np.ndarray:
        import ctypes as C
        n = int(self.offsets[-1])
        return np.ctypeslib.as_array(C.cast(self.seqs_ptr, C.POINTER(C.c_uint8)), shape=(max(n, 1),))[:n]

    @property
    def names(self) -> List[str]:
        import ctypes as C
        f = self._h.contents
        no = np.ctypeslib.as_array(f.name_off, shape=(self.num_records + 1,)) if self.num_records else np.zeros(1, np.int64)
        blob = C.string_at(f.names, int(no[-1]))
        return [blob[int(no[i]):int(no[i + 1])].decode("ascii", "replace") for i in range(self.num_records)]

    def write(self, path, raw_result, first_record: int = 0, width: int = 60, append: bool = False, threads: int = 0) -> int:
        """Two wrapped records per DBGA_OK pair of `raw_result` (a _lib.Result, full or COMPACT_ROWS)."""
        import ctypes as C
        n = self._lib.dbga_fasta_write(str(path).encode(), self._h, C.byref(raw_result), int(first_record), int(width),
                                       int(append), int(threads))
        if n < 0:
            raise RuntimeError(f"dbga_fasta_write({path}) failed")
        return int(n)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dbga_fasta_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

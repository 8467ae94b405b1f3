"""Alignment quality counts: `pairwise_alignment_score` / `sop` of the
reference (src/dbga/utils.py:440-534) as vectorised integer column classification.  A gap
column extends a gap iff the previous column has a gap in the same row and a character in
the other; so gap_open = number of maximal same-row gap runs, gap_extend = the rest."""
from __future__ import annotations

from itertools import combinations
from typing import Dict

import numpy as np


def _rows(aln, names=None):
    d = aln if isinstance(aln, dict) else aln.to_dict()
    names = list(d.keys()) if names is None else list(names)
    return names, [np.frombuffer(str(d[n]).encode(), np.uint8) for n in names]


def score_rows(r1: np.ndarray, r2: np.ndarray) -> Dict[str, int]:
    g1, g2 = r1 == ord("-"), r2 == ord("-")
    both = ~g1 & ~g2
    kind = np.where(both, 0, np.where(g1 & ~g2, 1, np.where(g2 & ~g1, 2, 3)))   # 3 = gap in both rows
    prev = np.concatenate([[0], kind[:-1]]) if len(kind) else kind
    gap = (kind == 1) | (kind == 2)
    extend = gap & (prev == kind)
    return {"match": int((both & (r1 == r2)).sum()), "mismatch": int((both & (r1 != r2)).sum()),
            "gap_open": int((gap & ~extend).sum()), "gap_extend": int(extend.sum())}


def pairwise_alignment_score(seqs) -> Dict[str, int]:
    """Counts for an alignment of exactly two rows (utils.py:440-504)."""
    names, rows = _rows(seqs)
    assert len(rows) == 2
    return score_rows(rows[0], rows[1])


def sop(seqs) -> Dict[str, int]:
    """Sum of pairs over every pair of rows (utils.py:507-534)."""
    names, rows = _rows(seqs)
    total = {"match": 0, "mismatch": 0, "gap_open": 0, "gap_extend": 0}
    for i, j in combinations(range(len(rows)), 2):
        for k, v in score_rows(rows[i], rows[j]).items():
            total[k] += v
    return total


def batch_scores(result) -> Dict[str, np.ndarray]:
    """The four counts for every aligned pair of a BatchResult."""
    P = len(result)
    out = {k: np.zeros(P, np.int64) for k in ("match", "mismatch", "gap_open", "gap_extend")}
    for i in range(P):
        if result.status[i] != 0:
            continue
        lo, hi = int(result.aln_off[i]), int(result.aln_off[i + 1])
        for k, v in score_rows(result.aln0[lo:hi], result.aln1[lo:hi]).items():
            out[k][i] = v
    return out

"""CPU oracle for DBGA's pairwise alignment hot path  --  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``dbga_b200/`` imports it, and the product path
raises when the CUDA library is missing instead of falling back to this code.

It restates, position-based, what the reference computes with Python dicts/objects:

* stages 1-2 (k-mers, duplicate k-mers, node ids, merge nodes, cycle check):
  reference ``src/dbga/utils.py:108-154`` (``get_kmers``, ``duplicate_kmers``),
  ``src/dbga/debruijn_pairwise.py:78-225`` (``_add_node`` / ``_get_seq_kmer_indices`` /
  ``add_debruijn``), ``:277-300`` (``extract_bubble``) and ``utils.py:266-291``
  (``debruijn_merge_correctness``);
* stage 3 orchestration (segment extraction, shortcut, first-column strip, tail):
  ``debruijn_pairwise.py:302-395, 418-543`` and ``utils.py:157-192``
  (``dna_global_aln``, including its empty-side cases);
* stage 3 arithmetic: the reference delegates to the third-party
  ``cogent3.align.global_pairwise`` (``utils.py:185``), pinned by the reference to the
  moving branch ``cogent3 @ git+https://github.com/cogent3/cogent3.git@develop``
  (``pyproject.toml:15``) and NOT available in this image.  It is restated here as an
  integer three-state affine-gap (Gotoh) global alignment whose every convention is an
  explicit parameter (:class:`Convention`).

Parity status
-------------
* Stages 1-2 + orchestration: pinned against the reference's own unmodified code run
  under a cogent3 shim (``oracle/make_golden.py`` -> ``tests/golden/ref_graph_*.json``)
  and against the reference's golden tests (``tests/test_debruijn_pairwise.py:50-214``).
* Stage 3 arithmetic vs *real* cogent3: **parity unpinned beyond the reference's golden
  vectors** (10 end-to-end goldens, ``tests/test_utils.py:82-95``).  The default
  convention (X, M, Y first-wins; gap(L) = d + (L-1) e; no X<->Y) is the only uniform
  tie order that reproduces all of them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

# status codes (shared with include/dbga_b200.h)
OK = 0
CYCLE = 1
K_INVALID = 2
BAD_CHAR = 3
TOO_LARGE = 4

NEG = -(1 << 30)
ORDERS = ("XMY", "XYM", "MXY", "MYX", "YXM", "YMX")
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


@dataclass(frozen=True)
class Convention:
    """Every DP convention cogent3 might use and the goldens do not pin.

    gap_len_mode : 0 -> gap of length L costs d + (L-1) e;  1 -> d + L e
    xy_allowed   : allow X->Y / Y->X transitions (gap in one row directly after a gap
                   in the other), charged as a gap open
    pred_order   : tie order among predecessor states, first wins, strict '>'
    end_order    : tie order among the three states at the terminal cell
    X = consumes a seq1 character (gap in seq2's row); Y = consumes a seq2 character.
    """

    gap_len_mode: int = 0
    xy_allowed: bool = False
    pred_order: str = "XMY"
    end_order: str = "XMY"


DEFAULT = Convention()


def make_dna_scoring_dict(match: int, transition: int, transversion: int) -> Dict[Tuple[str, str], int]:
    """cogent3.align.make_dna_scoring_dict, called at debruijn_pairwise.py:448-450.

    Transition = both purines (A, G) or both pyrimidines (C, T).
    """
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


def score_table(s: Dict[Tuple[str, str], int]) -> List[int]:
    """Flatten a scoring dict to the 16-entry ACGT x ACGT table the C-ABI takes."""
    return [int(s[a, b]) for a in "ACGT" for b in "ACGT"]


# --------------------------------------------------------------------------- stage 3 DP
def gotoh(x: str, y: str, s: Dict[Tuple[str, str], int], d: int, e: int,
          conv: Convention = DEFAULT) -> Tuple[int, str, str]:
    """Integer 3-state global affine-gap alignment; stands in for
    ``cogent3.align.global_pairwise(seq1, seq2, s, d, e)`` (utils.py:185).

    Both inputs must be non-empty (the empty cases are handled by
    :func:`dna_global_aln`, as in the reference).  Returns (score, aligned_x, aligned_y).
    """
    m, n = len(x), len(y)
    assert m > 0 and n > 0
    go = d if conv.gap_len_mode == 0 else d + e
    ge = e
    # state values; index [i][j]
    M = [[NEG] * (n + 1) for _ in range(m + 1)]
    X = [[NEG] * (n + 1) for _ in range(m + 1)]
    Y = [[NEG] * (n + 1) for _ in range(m + 1)]
    pM = [[""] * (n + 1) for _ in range(m + 1)]
    pX = [[""] * (n + 1) for _ in range(m + 1)]
    pY = [[""] * (n + 1) for _ in range(m + 1)]
    M[0][0] = 0
    for i in range(1, m + 1):
        X[i][0] = -(go + (i - 1) * ge)
        pX[i][0] = "X" if i > 1 else "M"
    for j in range(1, n + 1):
        Y[0][j] = -(go + (j - 1) * ge)
        pY[0][j] = "Y" if j > 1 else "M"
    order = conv.pred_order

    def pick(cands):
        best, who = None, ""
        for st in order:
            if st in cands:
                v = cands[st]
                if best is None or v > best:
                    best, who = v, st
        return best, who

    for i in range(1, m + 1):
        xi = x[i - 1]
        for j in range(1, n + 1):
            v, w = pick({"M": M[i - 1][j - 1], "X": X[i - 1][j - 1], "Y": Y[i - 1][j - 1]})
            M[i][j] = v + s[xi, y[j - 1]]
            pM[i][j] = w
            c = {"M": M[i - 1][j] - go, "X": X[i - 1][j] - ge}
            if conv.xy_allowed:
                c["Y"] = Y[i - 1][j] - go
            X[i][j], pX[i][j] = pick(c)
            c = {"M": M[i][j - 1] - go, "Y": Y[i][j - 1] - ge}
            if conv.xy_allowed:
                c["X"] = X[i][j - 1] - go
            Y[i][j], pY[i][j] = pick(c)

    best, st = None, ""
    for cand in conv.end_order:
        v = {"M": M, "X": X, "Y": Y}[cand][m][n]
        if best is None or v > best:
            best, st = v, cand
    score = best
    ax, ay = [], []
    i, j = m, n
    while i > 0 or j > 0:
        if st == "M":
            nxt = pM[i][j]
            ax.append(x[i - 1]); ay.append(y[j - 1])
            i -= 1; j -= 1
        elif st == "X":
            nxt = pX[i][j]
            ax.append(x[i - 1]); ay.append("-")
            i -= 1
        else:
            nxt = pY[i][j]
            ax.append("-"); ay.append(y[j - 1])
            j -= 1
        st = nxt
    return score, "".join(reversed(ax)), "".join(reversed(ay))


def dna_global_aln(seq1: str, seq2: str, s, d: int, e: int,
                   conv: Convention = DEFAULT) -> Tuple[str, str]:
    """utils.py:157-192, with global_pairwise replaced by :func:`gotoh`."""
    if seq1 and seq2:
        _, a, b = gotoh(seq1, seq2, s, d, e, conv)
        return a, b
    elif seq1:
        return seq1, "-" * len(seq1)
    elif seq2:
        return "-" * len(seq2), seq2
    return "", ""


# --------------------------------------------------------------------- stages 1 and 2
@dataclass
class Graph:
    """Position-based restatement of DeBruijnPairwise.__init__ (debruijn_pairwise.py:42-76)."""

    status: int
    k: int
    nondup0: List[bool]          # per k-mer position of seq 0: not a duplicate k-mer
    nondup1: List[bool]
    seq_node_idx: Dict[int, List[int]]
    merge_node_idx: List[int]    # node ids of merge nodes in seq-1 (second sequence) order
    anchors: List[Tuple[int, int]]  # (pos in s0, pos in s1) of merge k-mers, ascending pos in s1
    id_count: int
    seq_last_kmer_idx: List[int]


def build_graph(s0: str, s1: str, k: int) -> Graph:
    """Stages 1-2.  Follows utils.py:129-131 (k range check), utils.py:148-154
    (duplicates = count > 1 within EITHER sequence), debruijn_pairwise.py:94-107 (a
    non-duplicate k-mer already seen in seq 0 is a merge node and reuses its id; only
    middle nodes are registered) and :162-204 (start / end nodes, id order)."""
    for s in (s0, s1):
        if k < 0 or k > len(s):
            return Graph(K_INVALID, k, [], [], {}, [], [], 0, [])
    K0 = [s0[i:i + k] for i in range(len(s0) - k + 1)]
    K1 = [s1[i:i + k] for i in range(len(s1) - k + 1)]
    c0: Dict[str, int] = {}
    c1: Dict[str, int] = {}
    for x in K0:
        c0[x] = c0.get(x, 0) + 1
    for x in K1:
        c1[x] = c1.get(x, 0) + 1

    def dup(x):
        return c0.get(x, 0) > 1 or c1.get(x, 0) > 1

    nd0 = [not dup(x) for x in K0]
    nd1 = [not dup(x) for x in K1]
    ids0 = [0]
    id_of: Dict[str, int] = {}
    pos_of: Dict[str, int] = {}
    nxt = 1
    last0 = -1
    for p, x in enumerate(K0):
        if nd0[p]:
            id_of[x] = nxt
            pos_of[x] = p
            ids0.append(nxt)
            last0 = nxt
            nxt += 1
    # seq_last_kmer_idx is -1 when the sequence ends in a duplicate k-mer (:192-196)
    last0 = last0 if (K0 and nd0[-1]) else -1
    if not K0:
        last0 = -1
    ids0.append(nxt)   # end node of seq 0
    nxt += 1
    ids1 = [nxt]       # start node of seq 1
    nxt += 1
    merge: List[int] = []
    anchors: List[Tuple[int, int]] = []
    last1 = -1
    for q, x in enumerate(K1):
        if nd1[q]:
            if x in id_of:
                ids1.append(id_of[x])
                merge.append(id_of[x])
                anchors.append((pos_of[x], q))
                last1 = id_of[x]
            else:
                ids1.append(nxt)
                last1 = nxt
                nxt += 1
    last1 = last1 if (K1 and nd1[-1]) else -1
    ids1.append(nxt)
    nxt += 1
    status = OK
    # debruijn_merge_correctness (utils.py:266-291): merge ids must appear in the same
    # order in both sequences <=> strictly increasing along seq 1.
    for a, b in zip(merge, merge[1:]):
        if not a < b:
            status = CYCLE
            break
    return Graph(status, k, nd0, nd1, {0: ids0, 1: ids1}, merge, anchors, nxt, [last0, last1])


@dataclass
class PairResult:
    status: int
    anchors: List[Tuple[int, int]]
    score: int                 # sum of the DP scores of every segment handed to the DP
    aln0: str
    aln1: str
    dp_cells: int              # sum |x|*|y| over segments handed to the DP
    dp_segments: int
    graph: Optional[Graph] = None


def segments(s0: str, s1: str, anchors: Sequence[Tuple[int, int]]):
    """Yield (kind, x, y) for every alignment segment in output order.

    kind: 'first' (kept whole), 'mid' (first column dropped, debruijn_pairwise.py:499-500),
    'shortcut' (:384-388, contributes nothing), 'tail' (kept whole, :521-538),
    'whole' (no merge node, :453-457).  'anchor' entries carry the explicit anchor
    character appended at :508-509 (every merge node but the last)."""
    if not anchors:
        yield ("whole", s0, s1)
        return
    M = len(anchors)
    for i, (a, b) in enumerate(anchors):
        if i == 0:
            yield ("first", s0[:a], s1[:b])
        else:
            pa, pb = anchors[i - 1]
            if a - pa == 1 and b - pb == 1:
                yield ("shortcut", "", "")
            else:
                yield ("mid", s0[pa:a], s1[pb:b])
        if i != M - 1:
            yield ("anchor", s0[a], s1[b])
    a, b = anchors[-1]
    yield ("tail", s0[a:], s1[b:])


def align_pair(s0: str, s1: str, k: int, match: int = 10, transition: int = -1,
               transversion: int = -8, d: int = 10, e: int = 2,
               conv: Convention = DEFAULT, s: Optional[dict] = None) -> PairResult:
    """DeBruijnPairwise(...).alignment(...) restated (debruijn_pairwise.py:418-543)."""
    for ch in s0 + s1:
        if ch not in _CODE:
            return PairResult(BAD_CHAR, [], 0, "", "", 0, 0)
    g = build_graph(s0, s1, k)
    if g.status != OK:
        return PairResult(g.status, g.anchors if g.status == CYCLE else [], 0, "", "", 0, 0, g)
    if s is None:
        s = make_dna_scoring_dict(match, transition, transversion)
    out0: List[str] = []
    out1: List[str] = []
    score = 0
    cells = 0
    nseg = 0
    for kind, x, y in segments(s0, s1, g.anchors):
        if kind == "shortcut":
            continue
        if kind == "anchor":
            out0.append(x); out1.append(y)
            continue
        if x and y:
            sc, a, b = gotoh(x, y, s, d, e, conv)
            score += sc
            cells += len(x) * len(y)
            nseg += 1
        else:
            a, b = dna_global_aln(x, y, s, d, e, conv)
        if kind == "mid":
            a, b = a[1:], b[1:]
        out0.append(a); out1.append(b)
    return PairResult(OK, list(g.anchors), score, "".join(out0), "".join(out1), cells, nseg, g)


# ------------------------------------------------------------------ synthetic workloads
def synth_pair(n: int, p_sub: float, p_indel: float, seed: int):
    """SURVEY.md 8d generator: s0 uniform over ACGT; s1 = s0 with per-base substitutions
    (uniform over the 3 other bases) and indel events (half insertions of uniform random
    bases, half deletions; length ~ Geometric(mean 3))."""
    import numpy as np

    rng = np.random.default_rng(seed)
    a = rng.integers(0, 4, size=n, dtype=np.int8)
    out = []
    i = 0
    r_sub = rng.random(n)
    r_ind = rng.random(n)
    while i < n:
        if r_ind[i] < p_indel:
            L = int(rng.geometric(1.0 / 3.0))
            if rng.random() < 0.5:
                out.extend(rng.integers(0, 4, size=L).tolist())
                out.append(int(a[i]))
                i += 1
            else:
                i += L
            continue
        if r_sub[i] < p_sub:
            out.append(int((a[i] + rng.integers(1, 4)) % 4))
        else:
            out.append(int(a[i]))
        i += 1
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    s0 = lut[a].tobytes().decode()
    s1 = lut[np.asarray(out, dtype=np.int64)].tobytes().decode() if out else ""
    return s0, s1


# ---------------------------------------------------------------- k-size heuristics
def kmer_stats(s0: str, s1: str, k: int) -> Tuple[int, int, int]:
    """(|K0|, |K1|, |K0 & K1|) for the k-mer sets of the two sequences: the set arithmetic of
    predict_p (reference src/dbga/utils.py:387-391) on get_kmers (utils.py:108-131)."""
    if k < 0 or k > len(s0) or k > len(s1):
        raise ValueError("Invalid k size for kmers")          # utils.py:129-130
    a = {s0[i:i + k] for i in range(len(s0) - k + 1)}
    b = {s1[i:i + k] for i in range(len(s1) - k + 1)}
    return len(a), len(b), len(a & b)


def predict_p(s0: str, s1: str, k: int) -> float:
    """utils.py:371-395: (|K0 & K1| / |K0 | K1|) ** (1 / k)."""
    d0, d1, sh = kmer_stats(s0, s1, k)
    return (sh / (d0 + d1 - sh)) ** (1 / k)


def predict_final_p(s0: str, s1: str) -> float:
    """utils.py:398-413: the minimum of predict_p over k in [ceil(log4 L), ceil(log2 L) + 10)."""
    import math
    shortest = min(len(s0), len(s1))
    return min(predict_p(s0, s1, k) for k in range(math.ceil(math.log(shortest, 4)), math.ceil(math.log2(shortest)) + 10))


def kmers_larger_than_threshold(seq: str, k: int, threshold: float) -> bool:
    """src/dbga/kmersize.py:10-14: repeated k-mer occurrences / all occurrences > threshold."""
    n = len(seq) - k + 1
    distinct = len({seq[i:i + k] for i in range(n)})
    return (n - distinct) / n > threshold

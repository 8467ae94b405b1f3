"""Range check of the long-segment frames (dbga_b200/csrc/stage3_dp16.cu: dp_wf16s_kernel) on the CPU.

The packed int16 cell stores 4 * (V - base - off) + priority with V = H + ge * (i + j).  The strip kernel measures V
from a frame base per (row strip, block of S16_CB columns): 0 in block 0 and in strip 0, else chosen so that
max(M, X, Y) of the strip's top boundary row at the block's first column sits at S16_F0.  DESIGN.md section 4 argues
that every state of every cell of the strip in that block then lies in [-300, ulimit] (the encoding's valid range)
whatever the segment's size.  Here the three Gotoh matrices are computed exactly (int64, numpy, a row at a time) for
sequences chosen to push the bounds -- random, near-identical, shifted copies (a long terminal gap, then matches),
homopolymers against random, long internal gaps -- and every finite value is checked against its frame, for both strip
heights and for the steepest scoring make_dp_params lets the strip form take."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _const(path, pattern):
    m = re.search(pattern, open(os.path.join(ROOT, path)).read())
    assert m, pattern
    return int(m.group(1))


CB = _const("dbga_b200/csrc/common.cuh", r"constexpr int S16_CB = (\d+);")
F0 = _const("dbga_b200/csrc/common.cuh", r"constexpr int S16_F0 = (\d+);")
OFF = _const("dbga_b200/csrc/api.cu", r"h\.off = (\d+);")
ULIMIT = 8191 + OFF - 41          # api.cu: h.ulimit
FLOOR = -300                      # api.cu: "valid values ... >= -300"
NEG = -(1 << 40)


def gotoh_hat(x, y, score, go, ge):
    """T^ = max(M^, X^, Y^), and the three matrices, in the H^ = H + ge (i + j) frame (no X <-> Y, gap(L) = go + (L-1) ge)."""
    m, n = len(x), len(y)
    G = ge - go
    M = np.full((m + 1, n + 1), NEG, np.int64); X = M.copy(); Y = M.copy()
    M[0, 0] = 0
    Y[0, 1:] = G                                   # Y^[0][j] = -(go + (j-1) ge) + ge j
    X[1:, 0] = G
    S = score[np.asarray(x)[:, None], np.asarray(y)[None, :]] + 2 * ge
    for i in range(1, m + 1):
        T_prev = np.maximum(M[i - 1], np.maximum(X[i - 1], Y[i - 1]))
        M[i, 1:] = np.where(T_prev[:-1] > NEG // 2, S[i - 1] + T_prev[:-1], NEG)
        X[i, 1:] = np.maximum(np.where(M[i - 1, 1:] > NEG // 2, M[i - 1, 1:] + G, NEG), X[i - 1, 1:])
        cand = np.where(M[i, :-1] > NEG // 2, M[i, :-1] + G, NEG)       # Y^[i][j] from M^[i][j-1]
        Y[i, 1:] = np.maximum.accumulate(cand)
    return M, X, Y


def check_frames(x, y, score, go, ge, rows):
    M, X, Y = gotoh_hat(x, y, score, go, ge)
    T = np.maximum(M, np.maximum(X, Y))
    m, n = len(x), len(y)
    worst_lo, worst_hi = 0, 0
    for i0 in range(0, m, rows):
        i1 = min(m, i0 + rows)
        for c0 in range(0, n + 1, CB):
            c1 = min(n + 1, c0 + CB)
            base = 0 if (i0 == 0 or c0 == 0) else int(T[i0, c0]) - F0
            for A in (M, X, Y):
                blk = A[i0 + 1:i1 + 1, c0:c1]
                fin = blk[blk > NEG // 2]
                if fin.size:
                    worst_lo = min(worst_lo, int(fin.min()) - base)
                    worst_hi = max(worst_hi, int(fin.max()) - base)
    assert worst_lo >= FLOOR, worst_lo
    assert worst_hi <= ULIMIT, worst_hi
    return worst_lo, worst_hi


def dna_score(match, transition, transversion):
    s = np.full((4, 4), transversion, np.int64)
    for a, b in ((0, 2), (2, 0), (1, 3), (3, 1)):         # A<->G, C<->T
        s[a, b] = transition
    np.fill_diagonal(s, match)
    return s


def test_strip_frames_hold_the_int16_range():
    rng = np.random.default_rng(7)
    cases = []
    a = rng.integers(0, 4, 1700)
    cases.append((a, rng.integers(0, 4, 1500)))                                        # unrelated: scores sink
    b = a.copy(); mut = rng.random(a.size) < 0.03; b[mut] = (b[mut] + 1) & 3
    cases.append((a, b))                                                                # near-identical: scores climb
    cases.append((a, np.concatenate([rng.integers(0, 4, 600), a])[:1700]))              # shifted copy: gap, then matches
    cases.append((np.zeros(1300, np.int64), rng.integers(0, 4, 1400)))                  # homopolymer against random
    cases.append((np.concatenate([a[:400], a[1100:]]), a))                              # a 700-base internal gap
    cases.append((a[:1100], np.concatenate([a[:300], rng.integers(0, 4, 500), a[300:1100]])))   # a 500-base insertion
    for match, ts, tv, d, e in ((10, -1, -8, 10, 2), (17, -3, -20, 30, 3), (1, 0, -1, 0, 0), (-3, 2, 5, 7, 3)):
        score = dna_score(match, ts, tv)
        per_step = max(0, int(score.max())) + 2 * e
        assert per_step * (max(512, CB) + 80) + F0 + 300 <= ULIMIT        # the condition make_dp_params checks
        for x, y in cases:
            for rows in (256, 512):
                check_frames(x, y, score, d, e, rows)


def test_frame_deltas_chain_to_the_anchor():
    """The bookkeeping lane 0 of a strip does at every block start (stage3_dp16.cu): from the boundary entry's anchor
    a_c (encoded in the PRODUCER's frame) and the producer's own delta carried in the entry it derives its frame step
    delta^s_c = delta^(s-1)_c + a_c - a_(c-1), with a_0 = AF = 4 (S16_F0 - off).  Summed up, the steps must put the
    strip's base exactly S16_F0 below its anchor, T(i0, c CB) - S16_F0, for every strip and block, and every step
    must fit the 16 bits it travels in."""
    rng = np.random.default_rng(11)
    a = rng.integers(0, 4, 2300)
    b = a.copy(); mut = rng.random(a.size) < 0.05; b[mut] = (b[mut] + 1) & 3
    for x, y in ((a, b), (a, rng.integers(0, 4, 2100)), (a, np.concatenate([rng.integers(0, 4, 700), a])[:2300])):
        M, X, Y = gotoh_hat(x, y, dna_score(10, -1, -8), 10, 2)
        T = np.maximum(M, np.maximum(X, Y))
        AF = 4 * (F0 - OFF)
        for rows in (256, 512):
            nblk = (len(y) + 1 + CB - 1) // CB
            delta_prev_strip = np.zeros(nblk, np.int64)          # strip 0: base 0 everywhere
            base_prev_strip = np.zeros(nblk, np.int64)
            for i0 in range(rows, len(x), rows):
                delta, base = np.zeros(nblk, np.int64), np.zeros(nblk, np.int64)
                a_prev, bsum = AF, 0
                for c in range(1, nblk):
                    if c * CB > len(y):
                        break
                    a_c = 4 * (int(T[i0, c * CB]) - int(base_prev_strip[c]) - OFF)      # what the producer stored (tag stripped)
                    d = int(delta_prev_strip[c]) + a_c - a_prev
                    assert -32768 <= d <= 32767 and -32768 <= a_c <= 32767
                    a_prev = a_c
                    bsum += d
                    delta[c], base[c] = d, bsum // 4
                    assert bsum % 4 == 0 and base[c] == int(T[i0, c * CB]) - F0
                delta_prev_strip, base_prev_strip = delta, base

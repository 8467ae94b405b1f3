"""Synthetic closely-related DNA pairs for benchmarks (SURVEY.md 8d generator, batched):
s0 uniform over ACGT; s1 = s0 with per-base substitutions (uniform over the three other
bases) and indel events (half insertions of uniform random bases, half deletions, length
~ Geometric(mean 3)).  Returns the C-ABI layout directly: one uint8 buffer + offsets."""
from __future__ import annotations

import numpy as np

_LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


def synth_batch(P: int, n: int, p_sub: float, p_indel: float, seed: int, out: np.ndarray | None = None):
    """-> (buf uint8[total], offsets int64[2P+1]).  `out`, if given, is a preallocated
    (e.g. pinned) uint8 array of at least P * (2n + 64) bytes that receives the buffer."""
    rng = np.random.default_rng(seed)
    s0 = rng.integers(0, 4, size=(P, n), dtype=np.uint8)
    s1 = s0.copy()
    if p_sub > 0:
        mask = rng.random((P, n)) < p_sub
        s1[mask] = (s1[mask] + rng.integers(1, 4, size=int(mask.sum()), dtype=np.uint8)) & 3
    rows = [None] * P
    if p_indel > 0:
        ev = rng.random((P, n)) < p_indel
        ev_pairs, ev_pos = np.nonzero(ev)
        n_ev = len(ev_pairs)
        ev_len = rng.geometric(1.0 / 3.0, size=n_ev)
        ev_ins = rng.random(n_ev) < 0.5
        starts = np.searchsorted(ev_pairs, np.arange(P + 1))
        for i in range(P):
            lo, hi = starts[i], starts[i + 1]
            if lo == hi:
                continue
            row = s1[i]
            parts, cur = [], 0
            for e in range(lo, hi):
                pos = int(ev_pos[e])
                if pos < cur:
                    continue
                parts.append(row[cur:pos])
                if ev_ins[e]:
                    parts.append(rng.integers(0, 4, size=int(ev_len[e]), dtype=np.uint8))
                    cur = pos
                else:
                    cur = min(n, pos + int(ev_len[e]))
            parts.append(row[cur:])
            rows[i] = np.concatenate(parts)
    lens = np.empty(2 * P, np.int64)
    lens[0::2] = n
    lens[1::2] = [n if r is None else len(r) for r in rows]
    offsets = np.zeros(2 * P + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    buf = out[:total] if out is not None else np.empty(total, np.uint8)
    # fast path: pairs without indels are written with one strided copy
    plain = np.array([r is None for r in rows])
    o = offsets[:-1]
    for i in range(P):
        a = int(o[2 * i])
        buf[a:a + n] = _LUT[s0[i]]
        r = s1[i] if rows[i] is None else rows[i]
        b = int(o[2 * i + 1])
        buf[b:b + len(r)] = _LUT[r]
    return buf, offsets


def pairs_from_buffers(buf: np.ndarray, offsets: np.ndarray, idx):
    out = []
    for i in idx:
        a, b, c = int(offsets[2 * i]), int(offsets[2 * i + 1]), int(offsets[2 * i + 2])
        out.append((buf[a:b].tobytes().decode(), buf[b:c].tobytes().decode()))
    return out


def synth_long_pair(n: int = 1_000_000, n_bubbles: int = 10, seed: int = 1, background: float = 0.0003,
                    bubble_sub: float = 0.30, bubble_indel: float = 0.04, bubble_len=(2000, 5000)):
    """BASELINE config [4] ("single 1 Mb pair with long indel-rich bubbles (multi-kb segments)",
    SURVEY 8d): s0 uniform; s1 = s0 with a very low background substitution rate plus `n_bubbles`
    planted windows of 2-5 kb in which s1 is a heavily diverged copy of s0's window -- per base
    `bubble_sub` substitutions and `bubble_indel` indel events (half insertions of uniform random
    bases, half deletions, length ~ Geometric(mean 3)) -- so that no k=15 anchor survives inside a
    window and both sides of the bubble are multi-kb and of different lengths.
    -> (buf uint8, offsets int64[3]).  Seeds for which the reference raises "Cycles detected" exist
    (SURVEY 8d); callers search for the first seed the oracle accepts."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 4, n, dtype=np.uint8)
    b = a.copy()
    m = rng.random(n) < background
    b[m] = (b[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
    lo_l, hi_l = bubble_len
    stride = max(hi_l + 1000, (n - 2 * hi_l) // max(1, n_bubbles))
    cand = np.arange(hi_l, n - 2 * hi_l, stride)
    pos = np.sort(rng.choice(cand, min(n_bubbles, len(cand)), replace=False))
    parts, cur = [], 0
    for p in pos:
        L = int(rng.integers(lo_l, hi_l))
        parts.append(b[cur:p])
        win = a[p:p + L]
        out, i = [], 0
        sub = rng.random(L) < bubble_sub
        ev = rng.random(L) < bubble_indel
        ins = rng.random(L) < 0.5
        ln = rng.geometric(1.0 / 3.0, size=L)
        newb = rng.integers(1, 4, L, dtype=np.uint8)
        while i < L:
            if ev[i]:
                if ins[i]:
                    out.append(rng.integers(0, 4, int(ln[i]), dtype=np.uint8))
                else:
                    i += int(ln[i])
                    continue
            out.append(np.array([(win[i] + newb[i]) & 3 if sub[i] else win[i]], np.uint8))
            i += 1
        parts.append(np.concatenate(out) if out else np.zeros(0, np.uint8))
        cur = p + L
    parts.append(b[cur:])
    b2 = np.concatenate(parts)
    buf = np.concatenate([_LUT[a], _LUT[b2]])
    return buf, np.array([0, n, n + len(b2)], np.int64)

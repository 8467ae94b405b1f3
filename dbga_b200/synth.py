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

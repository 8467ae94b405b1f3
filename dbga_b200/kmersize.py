"""k-size heuristics of the reference on top of the GPU k-mer set statistics (SURVEY.md 8f,
rank 4): `predict_p` / `predict_final_p` (src/dbga/utils.py:371-413) and the duplicate-proportion
test of src/dbga/kmersize.py:10-27.  The k-mer sets are counted on the device
(`dbga_kmer_stats`: |K0|, |K1|, |K0 & K1| per pair); only the closing arithmetic is here.
There is no host-side set building behind these functions."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from .batch import default_aligner
from .utils import _raise_for_status


def _stats(pairs: Sequence[Tuple[str, str]], k: int, device: int = 0):
    status, d0, d1, sh = default_aligner(device).kmer_stats(pairs, k)
    for st in status:
        if st != 0:
            _raise_for_status(int(st))         # "Invalid k size for kmers" (utils.py:129-130), bad characters
    return d0, d1, sh


def predict_p_batch(pairs: Sequence[Tuple[str, str]], k: int, device: int = 0) -> np.ndarray:
    """predict_p (utils.py:371-395) for many pairs: (|K0 & K1| / |K0 | K1|) ** (1 / k)."""
    d0, d1, sh = _stats(pairs, k, device)
    return np.array([(int(s) / (int(a) + int(b) - int(s))) ** (1 / k) for a, b, s in zip(d0, d1, sh)], dtype=np.float64)


def predict_p(seqs, k: int) -> float:
    """Similarity estimate from the Jaccard index of the two k-mer sets (utils.py:371-395)."""
    return float(predict_p_batch([(str(seqs.seqs[0]), str(seqs.seqs[1]))], k)[0])


def k_range(shortest: int) -> range:
    """The k values predict_final_p minimises over (utils.py:411-412)."""
    return range(math.ceil(math.log(shortest, 4)), math.ceil(math.log2(shortest)) + 10)


def predict_final_p_batch(pairs: Sequence[Tuple[str, str]], device: int = 0) -> np.ndarray:
    """predict_final_p (utils.py:398-413) for many pairs: one device pass per k over the pairs
    whose range contains it."""
    ranges = [k_range(min(len(a), len(b))) for a, b in pairs]
    best = np.full(len(pairs), np.inf)
    for k in sorted({k for r in ranges for k in r}):
        idx = [i for i, r in enumerate(ranges) if k in r]
        p = predict_p_batch([pairs[i] for i in idx], k, device)
        best[idx] = np.minimum(best[idx], p)
    return best


def predict_final_p(seqs) -> float:
    return float(predict_final_p_batch([(str(seqs.seqs[0]), str(seqs.seqs[1]))])[0])


def kmers_larger_than_threshold(seq: str, k: int, threshold: float, device: int = 0) -> bool:
    """kmersize.py:10-14: is the proportion of repeated k-mer occurrences above the threshold?"""
    _, d1, _ = _stats([(seq, seq)], k, device)
    n = len(seq) - k + 1
    return (n - int(d1[0])) / n > threshold


def calculate_k(seqs: List[str], thresh: float, device: int = 0) -> int:
    """kmersize.py:17-27 on already loaded sequences: the smallest k (carried from one sequence to
    the next, as the reference does) at which no sequence exceeds the threshold."""
    k = 1
    for s in seqs:
        while kmers_larger_than_threshold(str(s), k, thresh, device):
            k += 1
    return k

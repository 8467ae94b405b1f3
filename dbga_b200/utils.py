"""Helpers of the hot path, mirroring the hot subset of the reference's src/dbga/utils.py
(same names, argument meaning and error behaviour); the alignment arithmetic goes to the
CUDA library instead of cogent3."""
from __future__ import annotations

from collections import Counter
from enum import Enum, auto
from pathlib import Path
from typing import Any, Dict, List, Set, Tuple, Union

from . import seqcoll
from .batch import default_aligner, make_dna_scoring_dict  # noqa: F401
from .seqcoll import SequenceCollection  # noqa: F401


def read_nucleotide_from_kmers(seq: str, k: int) -> str:
    """First character of every k-mer of an edge's concatenated k-mer text (utils.py:52-70)."""
    assert len(seq) % k == 0
    return seq[::k] if seq else ""


def load_sequences(data: Any, moltype: str):
    """path | list/tuple of str | SequenceCollection -> SequenceCollection (utils.py:73-105).
    list/tuple input gets the integer names 0, 1, ... (utils.py:100)."""
    if isinstance(data, (str, Path)):
        data = seqcoll.load_unaligned_seqs(data, moltype=moltype)
    elif isinstance(data, (list, tuple)) and all(isinstance(x, str) for x in data):
        data = seqcoll.make_unaligned_seqs(dict(enumerate(data)), moltype=moltype)
    if seqcoll.is_collection(data):
        return data
    raise ValueError("Invalid input for sequence argument")


def get_kmers(sequence: str, k: int) -> List[str]:
    """All overlapping k-mers; k outside [0, len] is an error (utils.py:108-131)."""
    if k < 0 or k > len(sequence):
        raise ValueError("Invalid k size for kmers")
    return [sequence[i:i + k] for i in range(len(sequence) - k + 1)]


def duplicate_kmers(kmer_seqs: List[List[str]]) -> Set[str]:
    """k-mers occurring more than once within any one of the sequences (utils.py:134-154)."""
    dup: Set[str] = set()
    for kmers in kmer_seqs:
        dup.update(x for x, c in Counter(kmers).items() if c > 1)
    return dup


def dna_global_aln(seq1: str, seq2: str, s: Dict[Tuple[str, str], int], d: int, e: int,
                   device: int = 0, **conv) -> Tuple[str, str]:
    """Global affine-gap alignment of two strings on the GPU (utils.py:157-192: both
    non-empty -> DP; one empty -> all-gap row; both empty -> ("", ""))."""
    res = default_aligner(device).global_align_pairs([(seq1, seq2)], d=d, e=e, s=s, **conv)
    _raise_for_status(int(res.status[0]))
    return res.alignment(0)


def debruijn_merge_correctness(seqs_expansion: List[List[Union[int, List[int]]]]) -> bool:
    """True iff both expansions have the same shape and the same merge ids slot by slot
    (utils.py:266-291)."""
    first, second = seqs_expansion[0], seqs_expansion[1]
    if any(len(x) != len(first) for x in seqs_expansion):
        return False
    for a, b in zip(first, second):
        if isinstance(a, int) and isinstance(b, int):
            if a != b:
                return False
        elif not (isinstance(a, list) and isinstance(b, list)):
            return False
    return True


def get_closest_odd(num: int, upper: bool) -> int:
    """utils.py:327-349."""
    if num % 2 == 1:
        return num
    return num + 1 if upper else num - 1


CYCLE_MESSAGE = "Cycles detected in de Bruijn graph, usually caused by small kmer sizes"


def _raise_for_status(status: int) -> None:
    """Status code of the C ABI -> the reference's exception (include/dbga_b200.h)."""
    if status == 0:
        return
    if status == 1:
        raise ValueError(CYCLE_MESSAGE)                      # debruijn_pairwise.py:73-76
    if status == 2:
        raise ValueError("Invalid k size for kmers")         # utils.py:129-130
    if status == 3:
        raise ValueError("sequences must consist of A, C, G, T for the pairwise DP "
                         "(the reference aligns with moltype 'dna', utils.py:184)")
    raise MemoryError("pair exceeds a workspace or numeric limit of dbga_b200 (status TOO_LARGE)")


# ---------------------------------------------------------------------------- graph types
class NodeType(Enum):
    start = auto()
    middle = auto()
    end = auto()


class Edge:
    """utils.py:544-582."""

    def __init__(self, out_node: int, in_node: int, duplicate_str: str, multiple_duplicate: bool) -> None:
        self.out_node = out_node
        self.in_node = in_node
        self.duplicate_str = duplicate_str
        self.multiple_duplicate = multiple_duplicate
        self.multiplicity = 1


class Node:
    """utils.py:585-650."""

    def __init__(self, id: int, kmer: str, node_type: NodeType) -> None:
        self.id = id
        self.kmer = kmer
        self.node_type = node_type
        self.out_edges: Dict[int, Edge] = {}
        self.next_nodes: Dict[int, Edge] = {}
        self.count = 1

    def add_out_edge(self, other_id: int, seq_idx: int, duplicate_str: str, multiple_duplicate: bool) -> None:
        if seq_idx in self.out_edges:
            self.out_edges[seq_idx].multiplicity += 1
            return
        # The reference constructs Edge(self.id, seq_idx, ...) -- it stores the sequence
        # index where the terminal node id is expected (utils.py:648); callers such as
        # adaptive_debruijn.py:183 read edge.in_node, so the quirk is kept.
        edge = Edge(self.id, seq_idx, duplicate_str, multiple_duplicate)
        self.out_edges[seq_idx] = edge
        self.next_nodes[other_id] = edge

"""`DeBruijnPairwise(data, k, moltype).alignment()` -- drop-in for the reference class
(src/dbga/debruijn_pairwise.py:10-543) whose three stages run on the GPU:

* the constructor runs stages 1-2 (k-mer hashing, merge-node detection, cycle check) through
  the C ABI and raises the reference's errors;
* `alignment()` runs the whole path (stage 3 included) and returns what the reference
  returns: a plain dict when there is no merge node (:453-457), an alignment object that
  compares equal to `{name: gapped_str}` otherwise (:543);
* the graph attributes the reference exposes (`nodes`, `seq_node_idx`, `merge_node_idx`,
  `id_count`, `exist_kmer`, `seq_last_kmer_idx`, `expansion`) are views materialised lazily
  on the host from the GPU's per-position flags by prefix sums.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Tuple, Union

import numpy as np

from . import seqcoll
from .batch import default_aligner, make_dna_scoring_dict
from .utils import (CYCLE_MESSAGE, Edge, Node, NodeType, _raise_for_status,  # noqa: F401
                    debruijn_merge_correctness, dna_global_aln, duplicate_kmers, get_kmers,
                    load_sequences, read_nucleotide_from_kmers)


_DEFAULT_KEY = (10, -1, -8, 10, 2, ())      # alignment()'s default arguments (debruijn_pairwise.py:420-424)


class DeBruijnPairwise:
    """Pairwise de Bruijn graph aligner.

    Attributes (as in the reference docstring, debruijn_pairwise.py:13-38): id_count, k,
    nodes, exist_kmer, seq_node_idx, merge_node_idx, seq_last_kmer_idx, moltype, names,
    sequences, num_seq, avg_len, sc, expansion.
    """

    def __init__(self, data: Any, k: int, moltype: str, device: int = 0) -> None:
        self.sc = load_sequences(data=data, moltype=moltype)
        self.k = k
        self.moltype: str = moltype
        self.names: Tuple[Any, ...] = tuple(self.sc.names)
        self.sequences: Tuple[str, ...] = tuple(str(seq) for seq in self.sc.seqs)
        self.num_seq: int = len(self.sequences)
        self.avg_len = sum(map(len, self.sequences)) / self.num_seq
        self._device = device
        self._views: Dict[str, Any] = {}
        self._cached = None
        assert self.num_seq == len(self.sequences) == 2          # debruijn_pairwise.py:216
        for seq in self.sequences:                               # utils.py:129-130
            if k < 0 or k > len(seq):
                raise ValueError("Invalid k size for kmers")
        s0, s1 = self.sequences
        if k == 0:
            # every (empty) k-mer is a duplicate: no middle node, no merge node
            self._nd0 = np.zeros(len(s0) + 1, np.uint8)
            self._nd1 = np.zeros(len(s1) + 1, np.uint8)
            self._anchors = np.zeros((0, 2), np.int32)
            return
        # One GPU call for the whole path with the reference's default scoring: the constructor needs stages 1-2
        # (graph flags, merge nodes, cycle check, debruijn_pairwise.py:70-76) and alignment() with its default
        # arguments then returns the rows computed here instead of running stages 1-2 a second time.
        res = default_aligner(device).align_pairs([(s0, s1)], k, stages=3, want_graph=True)
        status = int(res.status[0])
        if status not in (0, 1):
            _raise_for_status(status)
        self._nd0, self._nd1 = res.graph_flags(0)
        self._anchors = res.anchors(0)
        if status == 1:
            raise ValueError(CYCLE_MESSAGE)                      # debruijn_pairwise.py:73-76
        self._cached = (_DEFAULT_KEY, res.alignment(0))

    @classmethod
    def _from_graph(cls, sequences: Tuple[str, str], k: int, nd0: np.ndarray, nd1: np.ndarray, anchors: np.ndarray,
                    device: int = 0) -> "DeBruijnPairwise":
        """An instance over stage 1-2 results that a batched call already produced (adaptive descent: one call
        per level for all bubbles); same attributes and views as the constructor's."""
        self = object.__new__(cls)
        self.sc = load_sequences(data=tuple(sequences), moltype="dna")
        self.k, self.moltype = k, "dna"
        self.names = tuple(self.sc.names)
        self.sequences = tuple(str(x) for x in sequences)
        self.num_seq = 2
        self.avg_len = sum(map(len, self.sequences)) / 2
        self._device, self._views, self._cached = device, {}, None
        self._nd0, self._nd1, self._anchors = nd0, nd1, anchors
        return self

    # ------------------------------------------------------------------ graph views
    def _ids(self):
        """Node ids by prefix sums: seq 1 gets 0 (start), 1..u0 (non-duplicate k-mers in
        position order), u0+1 (end); seq 2 gets u0+2 (start), the seq-1 id for merge nodes
        and fresh ids otherwise, then its end node (debruijn_pairwise.py:78-107, 162-204)."""
        if "ids" in self._views:
            return self._views["ids"]
        nd0 = self._nd0.astype(bool)
        nd1 = self._nd1.astype(bool)
        id0 = np.cumsum(nd0)                       # id of the k-mer at p when nd0[p]
        u0 = int(id0[-1]) if len(id0) else 0
        ids0 = [0] + id0[nd0].tolist() + [u0 + 1]
        q_id = np.zeros(len(nd1), np.int64)
        is_merge = np.zeros(len(nd1), bool)
        if len(self._anchors):
            is_merge[self._anchors[:, 1]] = True
            q_id[self._anchors[:, 1]] = id0[self._anchors[:, 0]]
        fresh = nd1 & ~is_merge
        q_id[fresh] = u0 + 2 + np.cumsum(fresh)[fresh]
        n_fresh = int(fresh.sum())
        ids1 = [u0 + 2] + q_id[nd1].tolist() + [u0 + 3 + n_fresh]
        out = dict(ids0=ids0, ids1=ids1, id0=id0, q_id=q_id, id_count=u0 + 4 + n_fresh)
        self._views["ids"] = out
        return out

    @property
    def seq_node_idx(self) -> Dict[int, List[int]]:
        v = self._ids()
        return {0: list(v["ids0"]), 1: list(v["ids1"])}

    @property
    def merge_node_idx(self) -> List[int]:
        if "merge" not in self._views:
            v = self._ids()
            self._views["merge"] = [int(v["id0"][a]) for a in self._anchors[:, 0]]
        return self._views["merge"]

    @property
    def id_count(self) -> int:
        return self._ids()["id_count"]

    @property
    def seq_last_kmer_idx(self) -> List[int]:
        v = self._ids()
        out = []
        for nd, ids in ((self._nd0, v["ids0"]), (self._nd1, v["ids1"])):
            out.append(int(ids[-2]) if (len(nd) and nd[-1] and self.k > 0) else -1)
        return out

    def _build_nodes(self):
        """Materialise Node / Edge objects exactly as _get_seq_kmer_indices does
        (debruijn_pairwise.py:141-204): duplicate k-mers ride on the out-edge of the previous
        node as concatenated k-mer text."""
        if "nodes" in self._views:
            return
        v = self._ids()
        nodes: Dict[int, Node] = {}
        exist: Dict[str, int] = {}
        k = self.k
        for seq_idx, (seq, nd, ids) in enumerate(((self.sequences[0], self._nd0, v["ids0"]),
                                                  (self.sequences[1], self._nd1, v["ids1"]))):
            start = ids[0]
            nodes[start] = Node(start, "", NodeType.start)
            prev = start
            edge_kmers: List[str] = []
            multiple = False
            it = iter(ids[1:-1])
            for p in range(len(nd)):
                kmer = seq[p:p + k]
                if not nd[p]:
                    if edge_kmers:
                        multiple = True
                    edge_kmers.append(kmer)
                    continue
                cur = next(it)
                if cur in nodes:
                    nodes[cur].count += 1
                else:
                    nodes[cur] = Node(cur, kmer, NodeType.middle)
                    exist[kmer] = cur
                nodes[prev].add_out_edge(cur, seq_idx, "".join(edge_kmers), multiple if edge_kmers else False)
                edge_kmers = []
                prev = cur
            end = ids[-1]
            nodes[end] = Node(end, "", NodeType.end)
            nodes[prev].add_out_edge(end, seq_idx, "".join(edge_kmers), multiple)
        self._views["nodes"] = nodes
        self._views["exist"] = exist

    @property
    def nodes(self) -> Dict[int, Node]:
        self._build_nodes()
        return self._views["nodes"]

    @property
    def exist_kmer(self) -> Dict[str, int]:
        self._build_nodes()
        return self._views["exist"]

    def extract_bubble(self) -> List[List[Union[int, List[int]]]]:
        """[bubble, merge, bubble, ..., bubble] per sequence (debruijn_pairwise.py:277-300),
        with set membership instead of the reference's list scan."""
        merge = set(self.merge_node_idx)
        expansion = []
        for seq_idx in range(self.num_seq):
            bubbles: List[Union[int, List[int]]] = []
            bubble: List[int] = []
            for node_idx in self.seq_node_idx[seq_idx]:
                if node_idx in merge:
                    bubbles.append(bubble)
                    bubbles.append(node_idx)
                    bubble = []
                else:
                    bubble.append(node_idx)
            bubbles.append(bubble)
            expansion.append(bubbles)
        return expansion

    @property
    def expansion(self):
        if "expansion" not in self._views:
            self._views["expansion"] = self.extract_bubble()
        return self._views["expansion"]

    # ------------------------------------------------------------------ text helpers
    def read_from_kmer(self, node_idx: int, seq_idx: int) -> str:
        """debruijn_pairwise.py:257-275."""
        assert seq_idx in {0, 1}
        kmer = self.nodes[node_idx].kmer
        return kmer if node_idx == self.seq_last_kmer_idx[seq_idx] else kmer[0]

    def extract_bubble_seq(self, bubble_idx_seq: List[int], seq_idx: int) -> str:
        """Sequence text spanned by a node list whose last entry is a merge or end node
        (debruijn_pairwise.py:302-343)."""
        assert seq_idx in {0, 1}
        parts: List[str] = []
        for node_idx, nxt in zip(bubble_idx_seq[:-1], bubble_idx_seq[1:]):
            node = self.nodes[node_idx]
            if node.node_type is NodeType.middle:
                parts.append(self.read_from_kmer(node_idx, seq_idx))
            edge = node.out_edges[seq_idx]
            text = edge.duplicate_str
            if self.nodes[nxt].node_type is NodeType.end:
                if edge.multiple_duplicate:
                    parts.append(read_nucleotide_from_kmers(text[:-self.k], self.k))
                    parts.append(text[-self.k:])
                else:
                    parts.append(text)
            else:
                parts.append(read_nucleotide_from_kmers(text, self.k))
        return "".join(parts)

    def get_merge_edge(self, merge_node_idx: int) -> Tuple[str, str]:
        """debruijn_pairwise.py:397-416."""
        node = self.nodes[merge_node_idx]
        return (read_nucleotide_from_kmers(node.out_edges[0].duplicate_str, self.k),
                read_nucleotide_from_kmers(node.out_edges[1].duplicate_str, self.k))

    def bubble_aln(self, bubble_indices_seq1: List[int], bubble_indices_seq2: List[int],
                   s: Dict[Tuple[str, str], int], d: int, e: int, prev_edge_read1: str = "",
                   prev_edge_read2: str = "", prev_merge: str = "") -> Tuple[str, str]:
        """debruijn_pairwise.py:345-395 (kept for callers; alignment() itself batches every
        bubble of the pair into one GPU call instead of looping over this)."""
        if (len(bubble_indices_seq1) == len(bubble_indices_seq2) == 1
                and prev_edge_read1 == prev_edge_read2 == ""):
            return prev_merge, prev_merge
        x = f"{prev_merge}{prev_edge_read1}{self.extract_bubble_seq(bubble_indices_seq1, 0)}"
        y = f"{prev_merge}{prev_edge_read2}{self.extract_bubble_seq(bubble_indices_seq2, 1)}"
        return dna_global_aln(x, y, s=s, d=d, e=e, device=self._device)

    def visualize(self, path: str, save_dot: bool = False) -> None:  # pragma: no cover
        """debruijn_pairwise.py:227-255; needs graphviz, which is optional."""
        file_path = Path(path).expanduser().absolute()
        suffix = file_path.suffix[1:]
        if suffix not in ["pdf", "png", "svg"]:
            raise ValueError("Not supported file format")
        import graphviz
        dot = graphviz.Digraph("de Bruijn graph")
        dot.attr(rankdir="LR")
        for node in self.nodes.values():
            dot.node(str(node.id), f"{node.id}.{node.kmer}")
        for node in self.nodes.values():
            for other, edge in node.next_nodes.items():
                dot.edge(str(node.id), str(other), label=edge.duplicate_str, weight=str(edge.multiplicity))
        dot.render(outfile=file_path, format=suffix, cleanup=True)
        if save_dot:
            dot.save(file_path.with_suffix(".DOT"))

    # ------------------------------------------------------------------ stage 3
    def alignment(self, match: int = 10, transition: int = -1, transversion: int = -8,
                  d: int = 10, e: int = 2, **conv):
        """Align the two sequences (debruijn_pairwise.py:418-543).  `conv` forwards the DP
        convention switches of the C ABI (gap_len_mode, xy_allowed, pred_order, end_order)."""
        s = make_dna_scoring_dict(match=match, transition=transition, transversion=transversion)
        s0, s1 = self.sequences
        al = default_aligner(self._device)
        if self.k == 0 or len(self._anchors) == 0:
            res = al.global_align_pairs([(s0, s1)], d=d, e=e, s=s, **conv)
            _raise_for_status(int(res.status[0]))
            a0, a1 = res.alignment(0)
            return {self.names[0]: a0, self.names[1]: a1}            # plain dict (:453-457)
        key = (match, transition, transversion, d, e, tuple(sorted(conv.items())))
        if self._cached is not None and self._cached[0] == key:
            a0, a1 = self._cached[1]                                  # computed by the constructor's call
        else:
            res = al.align_pairs([(s0, s1)], self.k, d=d, e=e, s=s, **conv)
            _raise_for_status(int(res.status[0]))
            a0, a1 = res.alignment(0)
        return seqcoll.make_aligned_seqs({self.names[0]: a0, self.names[1]: a1}, moltype=self.moltype)

"""Adaptive (descending-k) de Bruijn alignment: the caller of the pairwise hot path that the
reference implements in src/dbga/adaptive_debruijn.py:17-224.

Every level builds a `DeBruijnPairwise` graph on the GPU (stages 1-2), keeps the merge nodes
as fixed columns and re-aligns each bubble with the next smaller k; a bubble that cannot be
split any further goes to the GPU global alignment (`dna_global_aln`).  The control flow is
host-side recursion exactly as in the reference; all sequence arithmetic is on the device.
"""
from __future__ import annotations

import math
from typing import Any, List, Sequence, Tuple

from .batch import make_dna_scoring_dict
from .debruijn_pairwise import DeBruijnPairwise
from .kmersize import predict_final_p, predict_p  # noqa: F401  (GPU k-mer set statistics, utils.py:371-413)
from .utils import NodeType, dna_global_aln, get_closest_odd, load_sequences


def get_seqs_entropy(seqs) -> float:
    """Shannon entropy (bits) of the k-mer spectrum of all sequences concatenated, with
    k = ceil(log4(total length)) (utils.py:352-369)."""
    text = "".join(str(s) for s in seqs.iter_seqs())
    k = math.ceil(math.log(len(text), 4))
    counts: dict = {}
    for i in range(len(text) - k + 1):
        counts[text[i:i + k]] = counts.get(text[i:i + k], 0) + 1
    total = sum(counts.values())
    return -sum(c / total * math.log2(c / total) for c in counts.values())


def _merge_node_text(dbg: DeBruijnPairwise, node_id: int, seq_idx: int) -> str:
    """Characters a merge node contributes to sequence `seq_idx` (adaptive_debruijn.py:170-204):
    its own first character, plus -- when its out-edge carries duplicate k-mers -- the whole
    edge text if the edge leads to the end node and the text's first character otherwise.
    `edge.in_node` keeps the reference's meaning (utils.py:648)."""
    text = dbg.read_from_kmer(node_id, seq_idx)
    edge = dbg.nodes[node_id].out_edges.get(seq_idx)
    if edge is not None and edge.duplicate_str:
        whole = dbg.nodes[edge.in_node].node_type == NodeType.end
        text += edge.duplicate_str if whole else edge.duplicate_str[0]
    return text


class _Task:
    """One (sub)alignment of the descent: its rows are the concatenation of `parts`, each either a finished
    (row0, row1) pair of strings or a child task."""
    __slots__ = ("seqs", "k_index", "parts", "leaf")

    def __init__(self, seqs, k_index):
        self.seqs, self.k_index, self.parts, self.leaf = tuple(seqs), k_index, None, None

    def rows(self) -> Tuple[str, str]:
        if self.leaf is not None:
            return self.leaf
        done = [p.rows() if isinstance(p, _Task) else p for p in self.parts]
        return "".join(x[0] for x in done), "".join(x[1] for x in done)


def adpt_dbg_alignment_batch(pairs: Sequence[Sequence[str]], k_index: int, k_choice: Tuple[int, ...], s: dict,
                             d: int, e: int, device: int = 0) -> List[Tuple[str, str]]:
    """The reference's recursion (adaptive_debruijn.py:91-224) run level by level: every bubble of every
    task that reaches k = k_choice[i] goes into ONE batched stage 1-2 call, the graphs come back
    together, and the bubbles they leave form the next level; everything that cannot be split any
    further is aligned by one batched dna_global_aln call at the end.  Same results as the recursion
    (a cycle at any level raises the reference's ValueError), a C-ABI call per level instead of one
    per bubble per level."""
    from .batch import default_aligner
    from .utils import CYCLE_MESSAGE, _raise_for_status
    al = default_aligner(device)
    roots = [_Task(p, k_index) for p in pairs]
    level, leaves = list(roots), []
    while level:
        todo, nxt = [], []
        for t in level:
            if not 0 <= t.k_index < len(k_choice) or k_choice[t.k_index] <= 0 or k_choice[t.k_index] > min(map(len, t.seqs)):
                leaves.append(t)
            else:
                todo.append(t)
        by_k: dict = {}
        for t in todo:
            by_k.setdefault(k_choice[t.k_index], []).append(t)
        for k, tasks in by_k.items():
            res = al.align_pairs([t.seqs for t in tasks], k, stages=2, want_graph=True)
            for i, t in enumerate(tasks):
                st = int(res.status[i])
                if st == 1:
                    raise ValueError(CYCLE_MESSAGE)                  # debruijn_pairwise.py:73-76
                if st != 0:
                    _raise_for_status(st)
                nd0, nd1 = res.graph_flags(i)
                dbg = DeBruijnPairwise._from_graph(t.seqs, k, nd0, nd1, res.anchors(i), device)
                if len(dbg.merge_node_idx) < 2:
                    child = _Task(t.seqs, t.k_index + 1)            # nothing to anchor on at this k
                    t.parts = [child]
                    nxt.append(child)
                    continue
                exp1 = [list(x) if isinstance(x, list) else x for x in dbg.expansion[0]]
                exp2 = [list(x) if isinstance(x, list) else x for x in dbg.expansion[1]]
                t.parts = []
                for j in range(len(exp1) - 2):
                    part1, part2 = exp1[j], exp2[j]
                    if isinstance(part1, list):
                        # a bubble is always followed by its closing merge node
                        child = _Task((dbg.extract_bubble_seq(part1 + [exp1[j + 1]], 0),
                                       dbg.extract_bubble_seq(part2 + [exp2[j + 1]], 1)), t.k_index + 1)
                        t.parts.append(child)
                        nxt.append(child)
                    else:
                        t.parts.append((_merge_node_text(dbg, part1, 0), _merge_node_text(dbg, part2, 1)))
                # last merge node + last bubble (the last merge node may be the end of a sequence)
                child = _Task((dbg.extract_bubble_seq([exp1[-2]] + exp1[-1], 0),
                               dbg.extract_bubble_seq([exp2[-2]] + exp2[-1], 1)), t.k_index + 1)
                t.parts.append(child)
                nxt.append(child)
        level = nxt
    # ---- leaves: dna_global_aln (utils.py:157-192), one batched call for those with two non-empty sides
    both = [t for t in leaves if t.seqs[0] and t.seqs[1]]
    if both:
        res = al.global_align_pairs([t.seqs for t in both], d=d, e=e, s=s)
        for i, t in enumerate(both):
            _raise_for_status(int(res.status[i]))
            t.leaf = res.alignment(i)
    for t in leaves:
        if t.leaf is None:                                           # an empty side: utils.py:187-192
            x, y = t.seqs
            t.leaf = (x, "-" * len(x)) if x else (("-" * len(y), y) if y else ("", ""))
    return [t.rows() for t in roots]


def adpt_dbg_alignment_recursive(seqs: Sequence[str], k_index: int, k_choice: Tuple[int, ...], s: dict,
                                 d: int, e: int) -> Tuple[str, ...]:
    """Align two strings, descending through `k_choice` (adaptive_debruijn.py:91-224)."""
    return adpt_dbg_alignment_batch([tuple(seqs)], k_index, k_choice, s, d, e)[0]


def adpt_dbg_alignment(data: Any, moltype: str, match: int = 10, transition: int = -1, transversion: int = -8,
                       d: int = 10, e: int = 2) -> str:
    """Top-level adaptive alignment returning FASTA text (adaptive_debruijn.py:17-88): k runs
    over the odd numbers from about log2(len)+10 down to max(13, log4(len)); dissimilar or
    low-complexity inputs go straight to the global alignment."""
    sc = load_sequences(data, moltype=moltype)
    shortest = min(len(x) for x in sc.seqs)
    min_k = max(13, math.ceil(math.log(shortest, 4)))
    max_k = math.ceil(math.log2(shortest)) + 10
    k_choice = tuple(range(get_closest_odd(max_k, True), get_closest_odd(min_k, False) - 1, -2))
    s = make_dna_scoring_dict(match=match, transition=transition, transversion=transversion)
    strings = tuple(str(x) for x in sc.iter_seqs())
    if predict_final_p(sc) < 0.80 or get_seqs_entropy(sc) < 11:
        aln = dna_global_aln(strings[0], strings[1], s, d, e)
    else:
        aln = adpt_dbg_alignment_recursive(strings, 0, k_choice, s, d, e)
    return alignment_to_fasta(dict(zip(sc.names, aln)))


def alignment_to_fasta(alignment_dict, block_size: int = 60) -> str:
    """cogent3.format.fasta.alignment_to_fasta as the reference calls it (adaptive_debruijn.py:86-88,
    default block_size=60): ">label", then the row in blocks of 60 columns, one per line."""
    out = []
    for name, row in alignment_dict.items():
        out.append(f">{name}\n")
        row = str(row)
        for i in range(0, len(row), block_size):
            out.append(row[i:i + block_size] + "\n")
    return "".join(out)

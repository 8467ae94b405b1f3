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


def adpt_dbg_alignment_recursive(seqs: Sequence[str], k_index: int, k_choice: Tuple[int, ...], s: dict,
                                 d: int, e: int) -> Tuple[str, ...]:
    """Align two strings, descending through `k_choice` (adaptive_debruijn.py:91-224)."""
    def leaf():
        return dna_global_aln(*seqs, s, d, e)

    if not 0 <= k_index < len(k_choice):
        return leaf()
    k = k_choice[k_index]
    if k <= 0 or k > min(map(len, seqs)):
        return leaf()
    dbg = DeBruijnPairwise(tuple(seqs), k, moltype="dna")
    if len(dbg.merge_node_idx) < 2:
        # no (or a single) merge node: nothing to anchor on at this k
        return adpt_dbg_alignment_recursive(seqs, k_index + 1, k_choice, s, d, e)

    exp1, exp2 = [list(x) if isinstance(x, list) else x for x in dbg.expansion[0]], \
                 [list(x) if isinstance(x, list) else x for x in dbg.expansion[1]]
    rows: Tuple[List[str], List[str]] = ([], [])
    for i in range(len(exp1) - 2):
        part1, part2 = exp1[i], exp2[i]
        if isinstance(part1, list):
            # a bubble is always followed by its closing merge node
            sub = (dbg.extract_bubble_seq(part1 + [exp1[i + 1]], 0),
                   dbg.extract_bubble_seq(part2 + [exp2[i + 1]], 1))
            aligned = adpt_dbg_alignment_recursive(sub, k_index + 1, k_choice, s, d, e)
            rows[0].append(aligned[0]); rows[1].append(aligned[1])
        else:
            rows[0].append(_merge_node_text(dbg, part1, 0))
            rows[1].append(_merge_node_text(dbg, part2, 1))
    # last merge node + last bubble (the last merge node may be the end of a sequence)
    tail = (dbg.extract_bubble_seq([exp1[-2]] + exp1[-1], 0), dbg.extract_bubble_seq([exp2[-2]] + exp2[-1], 1))
    aligned = adpt_dbg_alignment_recursive(tail, k_index + 1, k_choice, s, d, e)
    rows[0].append(aligned[0]); rows[1].append(aligned[1])
    return "".join(rows[0]), "".join(rows[1])


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

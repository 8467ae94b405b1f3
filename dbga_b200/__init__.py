"""dbga_b200: B200-native implementation of DBGA's pairwise alignment hot path, behind
DBGA's own interface (`DeBruijnPairwise(data, k, moltype).alignment()`), plus a batch API.
"""
__version__ = "0.1.0"

from .batch import Aligner, BatchResult, MultiAligner, align_batch, make_dna_scoring_dict, pack_ascii  # noqa: F401
from .debruijn_pairwise import DeBruijnPairwise  # noqa: F401
from .utils import (NodeType, Node, Edge, load_sequences, get_kmers, duplicate_kmers,  # noqa: F401
                    dna_global_aln, debruijn_merge_correctness, read_nucleotide_from_kmers,
                    get_closest_odd)

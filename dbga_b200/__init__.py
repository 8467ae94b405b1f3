"""dbga_b200: B200-native implementation of DBGA's pairwise alignment hot path, behind
DBGA's own interface (`DeBruijnPairwise(data, k, moltype).alignment()`), plus a batch API.
"""
__version__ = "0.1.0"

from .batch import Aligner, BatchResult, align_batch, make_dna_scoring_dict  # noqa: F401

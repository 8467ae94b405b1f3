import sys
from cogent3 import ArrayAlignment

# the convention the shimmed global_pairwise uses; tests may rebind it
CONVENTION = None


def make_dna_scoring_dict(match, transition, transversion):
    from oracle import dbga_oracle
    return dbga_oracle.make_dna_scoring_dict(match, transition, transversion)


DP_CALLS = []   # (seq1, seq2) of every call, so tests can compare the exact DP call list


def global_pairwise(seq1, seq2, s, d, e, **kw):
    from oracle import dbga_oracle
    conv = CONVENTION or dbga_oracle.DEFAULT
    DP_CALLS.append((str(seq1), str(seq2)))
    _, a, b = dbga_oracle.gotoh(str(seq1), str(seq2), s, d, e, conv)
    return ArrayAlignment({getattr(seq1, "name", 0): a, getattr(seq2, "name", 1): b})

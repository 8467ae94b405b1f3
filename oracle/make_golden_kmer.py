"""Generate tests/golden/ref_kmer_heuristics.json by running the reference's OWN, UNMODIFIED
k-size heuristics: predict_p / predict_final_p (src/dbga/utils.py:371-413) and
kmers_larger_than_threshold (src/dbga/kmersize.py:10-14), under the cogent3 import shim (these
functions use cogent3 only for the SequenceCollection container).

Run in the build container only (needs /root/reference):   python oracle/make_golden_kmer.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "cogent3_shim"), os.path.join(REF, "src"), ROOT]

from dbga.utils import load_sequences, predict_final_p, predict_p  # noqa: E402  (the reference)
from dbga.kmersize import kmers_larger_than_threshold  # noqa: E402
from oracle import dbga_oracle as O  # noqa: E402


def main():
    rng = random.Random(20261020)
    cases = []
    for n, ps, pi in [(40, 0.05, 0.02), (120, 0.02, 0.01), (400, 0.1, 0.02), (1500, 0.018, 0.002), (1500, 0.3, 0.05),
                      (3000, 0.01, 0.001), (60, 0.0, 0.0), (200, 0.5, 0.1)]:
        for rep in range(3):
            s0, s1 = O.synth_pair(n, ps, pi, rng.randrange(1 << 30))
            sc = load_sequences([s0, s1], moltype="dna")
            rec = {"s0": s0, "s1": s1, "predict_p": {}, "dup_prop_gt": {}}
            for k in (1, 2, 3, 5, 9, 13, 15, 16, 21, 31):
                if k <= min(len(s0), len(s1)):
                    rec["predict_p"][str(k)] = predict_p(sc, k)
            rec["predict_final_p"] = predict_final_p(sc)
            for k, th in ((1, 0.5), (3, 0.3), (5, 0.1), (7, 0.01), (9, 0.001)):
                if k <= len(s0):
                    rec["dup_prop_gt"][f"{k}:{th}"] = bool(kmers_larger_than_threshold(s0, k, th))
            cases.append(rec)
    out = os.path.join(ROOT, "tests", "golden", "ref_kmer_heuristics.json")
    json.dump({"source": "reference src/dbga/utils.py:371-413, src/dbga/kmersize.py:10-14 (unmodified, cogent3 shim)",
               "cases": cases}, open(out, "w"))
    print(out, len(cases), "cases")


if __name__ == "__main__":
    main()

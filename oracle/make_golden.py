"""Generate tests/golden/*.json by running the reference's OWN, UNMODIFIED code.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

What it pins
------------
``ref_graph_fuzz.json``  -- for seeded random inputs, everything the reference's
``DeBruijnPairwise`` computes without cogent3 arithmetic: ``seq_node_idx``,
``merge_node_idx``, ``id_count``, ``seq_last_kmer_idx``, the cycle ``ValueError``, and the
exact list of ``(seq1, seq2)`` strings it hands to ``cogent3.align.global_pairwise``
(``src/dbga/utils.py:185``).  cogent3 itself is replaced by ``oracle/cogent3_shim`` whose
``global_pairwise`` is the integer Gotoh oracle under the default convention; the gapped
output recorded here therefore pins the reference's segment orchestration
(``debruijn_pairwise.py:418-543``), not cogent3's tie-breaking.

``ref_fixtures.json``    -- the reference's golden vectors transcribed as data:
``tests/test_debruijn_pairwise.py:50-214`` (10 end-to-end cases + 2 cycle cases, inputs
read from ``tests/data/*.fasta``), ``tests/test_utils.py:82-95`` (dna_global_aln,
including the either/or case), ``:113-155`` (get_kmers / duplicate_kmers).  These are
the only vectors that touch *real* cogent3 behaviour.  The script re-runs each of them
through the shimmed reference and refuses to write the file if any disagrees.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "cogent3_shim"), os.path.join(REF, "src"), ROOT]

from dbga.debruijn_pairwise import DeBruijnPairwise  # noqa: E402  (the reference)
from dbga.utils import dna_global_aln, get_kmers, duplicate_kmers  # noqa: E402
import cogent3.align as shim_align  # noqa: E402
from oracle import dbga_oracle as O  # noqa: E402


def read_fasta(path):
    names, seqs = [], []
    for line in open(path).read().splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            names.append(line[1:]); seqs.append("")
        else:
            seqs[-1] += line
    return names, seqs


# ---- the reference's golden vectors (tests/test_debruijn_pairwise.py:50-190), k = 3
E2E = [
    # fixture, exp_seq1_idx, exp_seq2_idx, exp_merge, exp_aln
    ("substitution_middle", list(range(10)), [10, 1, 2, 3, 11, 12, 13, 7, 8, 14], [1, 2, 3, 7, 8], ("GTACAAGCGA", "GTACACGCGA")),
    ("gap_bubble_duplicate", list(range(9)), [9, 1, 2, 3, 4, 10, 11, 7, 12], [1, 2, 3, 4, 7], ("GTACACGTATG", "GTACACG-ATG")),
    ("gap_at_end", list(range(12)), [12, 1, 2, 3, 4, 5, 6, 7, 8, 13], [1, 2, 3, 4, 5, 6, 7, 8], ("GTACAAGCGATG", "GTACAAGCGA--")),
    ("gap_at_start", list(range(11)), [11, 2, 3, 4, 5, 6, 7, 8, 9, 12], [2, 3, 4, 5, 6, 7, 8, 9], ("TGTACAAGCGA", "-GTACAAGCGA")),
    ("bubble_consecutive_duplicate", list(range(7)), [7, 1, 2, 8, 9, 10, 4, 5, 11], [1, 2, 4, 5], ("TGTACTGTAGA", "TGTACTATAGA")),
    ("duplicate_unbalanced_end", list(range(11)), [11, 1, 12, 13, 14, 5, 6, 7, 8, 15], [1, 5, 6, 7, 8], ("TGTACGTCAATGTCG", "TGTAAGTCAATG---")),
    ("unbalanced_end_w_duplicate", list(range(11)), [11, 1, 12, 13, 14, 5, 6, 7, 8, 15], [1, 5, 6, 7, 8], ("TGTACGTCAATGTCG", "TGTAAGTCAATGT--")),
    ("duplicate_kmer_in_bubble", list(range(9)), [9, 1, 10, 11, 7, 12], [1, 7], ("TAC-ACGTAAT", "TACGACG-AAT")),
    ("both_end_duplicate", list(range(6)), [6, 7, 8, 9, 3, 4, 10], [3, 4], ("ACGTGACG", "ACTTGACG")),
    ("unrelated_sequences", list(range(8)), list(range(8, 18)), [], ("ACGT--AGTATC", "ACCTACAGCA--")),
]
CYCLES = [("TACCACGTAAT", "TACGACCTAAT"), ("TACCGTCCAGACGTAAT", "TACGGTCCAGACCTAAT")]  # :193-214


def run_ref(s0, s1, k, names=None):
    shim_align.DP_CALLS.clear()
    try:
        db = DeBruijnPairwise([s0, s1], k, "dna")
    except ValueError as ex:
        return {"error": str(ex)}
    aln = db.alignment()
    d = aln if isinstance(aln, dict) else aln.to_dict()
    return {
        "seq_node_idx": [db.seq_node_idx[0], db.seq_node_idx[1]],
        "merge_node_idx": list(db.merge_node_idx),
        "id_count": db.id_count,
        "seq_last_kmer_idx": list(db.seq_last_kmer_idx),
        "dp_calls": [list(c) for c in shim_align.DP_CALLS],
        "aln": [d[0], d[1]],
        "returns_dict": isinstance(aln, dict),
    }


def mutate(rng, s, ps, pi):
    out, i = [], 0
    while i < len(s):
        r = rng.random()
        if r < pi:
            if rng.random() < 0.5:
                out.append("".join(rng.choice("ACGT") for _ in range(rng.randint(1, 4))))
                out.append(s[i]); i += 1
            else:
                i += rng.randint(1, 4)
        elif r < pi + ps:
            out.append(rng.choice([c for c in "ACGT" if c != s[i]])); i += 1
        else:
            out.append(s[i]); i += 1
    return "".join(out)


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "tests", "data"), exist_ok=True)
    fixtures = {"e2e": [], "cycles": [], "global_aln": [], "kmers": []}
    for name, i1, i2, merge, aln in E2E:
        path = os.path.join(REF, "tests", "data", name + ".fasta")
        names, seqs = read_fasta(path)
        got = run_ref(seqs[0], seqs[1], 3)
        assert got["seq_node_idx"] == [i1, i2], name
        assert got["merge_node_idx"] == merge, name
        assert tuple(got["aln"]) == aln, (name, got["aln"])
        fixtures["e2e"].append({"fixture": name, "names": names, "seqs": seqs, "k": 3,
                                "seq_node_idx": [i1, i2], "merge_node_idx": merge,
                                "aln": list(aln), "ref": "tests/test_debruijn_pairwise.py:50-190"})
    for s0, s1 in CYCLES:
        got = run_ref(s0, s1, 3)
        assert got.get("error", "").startswith("Cycles detected"), got
        fixtures["cycles"].append({"seqs": [s0, s1], "k": 3, "error": got["error"],
                                   "ref": "tests/test_debruijn_pairwise.py:193-214"})
    s = O.make_dna_scoring_dict(10, -1, -8)
    ga = [("", "ACG", [["---", "ACG"]]), ("CTA", "", [["CTA", "---"]]),
          ("GGATCGA", "GAATTCAGTTA", [["GGA-TC-G--A", "GAATTCAGTTA"], ["GGAT-C-G--A", "GAATTCAGTTA"]])]
    for a, b, accept in ga:   # tests/test_utils.py:82-95
        got = list(dna_global_aln(a, b, s, 10, 2))
        assert got in accept, (a, b, got)
        fixtures["global_aln"].append({"seqs": [a, b], "accept": accept, "ref": "tests/test_utils.py:75-95"})
    assert get_kmers("GTACACGTATG", 3) == ["GTA", "TAC", "ACA", "CAC", "ACG", "CGT", "GTA", "TAT", "ATG"]
    fixtures["kmers"].append({"seq": "GTACACGTATG", "k": 3, "kmers": get_kmers("GTACACGTATG", 3)})
    fixtures["kmers"].append({"seq": "TACCACGTAAT", "k": 4, "kmers": get_kmers("TACCACGTAAT", 4)})
    assert duplicate_kmers([["AG", "CT", "CA", "AC", "CT"], ["CT", "AC", "GT", "AG"]]) == {"CT"}
    # adaptive aligner goldens (reference tests/test_adaptive_debruijn.py:19-48)
    from dbga.adaptive_debruijn import adpt_dbg_alignment_recursive
    ADAPT = [("gap_bubble_duplicate", (3, 2), ("GTACACGTATG", "GTACACG-ATG")),
             ("gap_bubble_duplicate", (7, 5, 3, 2), ("GTACACGTATG", "GTACACG-ATG")),
             ("duplicate_unbalanced_end", (3,), ("TGTACGTCAATGTCG", "TGTAAGTCAATG---")),
             ("duplicate_in_stem", (3,), ("GTACACGTATG", "GTACTCGTATG"))]
    fixtures["adaptive"] = []
    for name, kc, aln in ADAPT:
        _, seqs = read_fasta(os.path.join(REF, "tests", "data", name + ".fasta"))
        got = adpt_dbg_alignment_recursive(seqs, 0, kc, s, 10, 2)
        assert tuple(got) == aln, (name, got)
        fixtures["adaptive"].append({"seqs": seqs, "k_choice": list(kc), "aln": list(aln),
                                     "ref": "tests/test_adaptive_debruijn.py:19-48"})
    # recorded outputs of the reference's adaptive recursion on random related pairs
    arng = random.Random(77)
    fixtures["adaptive_fuzz"] = []
    while len(fixtures["adaptive_fuzz"]) < 40:
        n = arng.choice([60, 150, 400])
        s0 = "".join(arng.choice("ACGT") for _ in range(n))
        s1 = mutate(arng, s0, 0.03, 0.01)
        kc = arng.choice([(9, 5, 3), (7, 4), (11, 7, 5, 3)])
        try:
            got = adpt_dbg_alignment_recursive((s0, s1), 0, kc, s, 10, 2)
        except Exception as ex:  # cycles at some level: the reference raises
            fixtures["adaptive_fuzz"].append({"seqs": [s0, s1], "k_choice": list(kc), "error": type(ex).__name__})
            continue
        fixtures["adaptive_fuzz"].append({"seqs": [s0, s1], "k_choice": list(kc), "aln": list(got)})
    # the top-level adaptive front-end (adaptive_debruijn.py:17-88) on pairs long enough to pass its entropy and
    # similarity gates (k runs from about log2(len) + 10 down to 13); FASTA text with 60-column rows
    from dbga.adaptive_debruijn import adpt_dbg_alignment
    trng = random.Random(99)
    fixtures["adaptive_top"] = []
    for n, ps, pi in ((3000, 0.01, 0.003), (5000, 0.02, 0.004), (4000, 0.005, 0.001), (2500, 0.3, 0.05)):
        s0 = "".join(trng.choice("ACGT") for _ in range(n))
        s1 = mutate(trng, s0, ps, pi)
        try:
            text = adpt_dbg_alignment([s0, s1], "dna")
            fixtures["adaptive_top"].append({"seqs": [s0, s1], "fasta": text})
        except Exception as ex:
            fixtures["adaptive_top"].append({"seqs": [s0, s1], "error": type(ex).__name__})
    fixtures["sop_fuzz"] = sop_fuzz()
    json.dump(fixtures, open(os.path.join(ROOT, "tests", "golden", "ref_fixtures.json"), "w"), indent=1)

    rng = random.Random(20261019)
    cases = []
    while len(cases) < 400:
        n = rng.choice([8, 12, 20, 40, 80, 150, 300])
        k = rng.choice([2, 3, 4, 5, 7, 9, 11])
        alpha = rng.choice(["ACGT", "ACGT", "AC", "ACG"])
        s0 = "".join(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.85:
            s1 = mutate(rng, s0, rng.choice([0, 0.02, 0.1, 0.3]), rng.choice([0, 0.02, 0.1]))
        else:
            s1 = "".join(rng.choice(alpha) for _ in range(rng.randint(1, n)))
        if len(s0) < k or len(s1) < k:
            continue
        cases.append({"seqs": [s0, s1], "k": k, "ref": run_ref(s0, s1, k)})
    ncyc = sum("error" in c["ref"] for c in cases)
    json.dump({"generator": "oracle/make_golden.py", "n_cycle": ncyc, "cases": cases},
              open(os.path.join(ROOT, "tests", "golden", "ref_graph_fuzz.json"), "w"))
    print(f"wrote {len(cases)} fuzz cases ({ncyc} cycle errors) and "
          f"{len(fixtures['e2e'])} e2e fixtures")
    # the FASTA inputs of the reference's e2e goldens are test vectors (8-17 nt);
    # they are re-emitted from the parsed data rather than copied.
    for fx in fixtures["e2e"]:
        with open(os.path.join(ROOT, "tests", "data", fx["fixture"] + ".fasta"), "w") as f:
            for nm, sq in zip(fx["names"], fx["seqs"]):
                f.write(f">{nm}\n{sq}\n")


class _TwoRows:
    """What pairwise_alignment_score reads of a cogent3 ArrayAlignment (src/dbga/utils.py:440-504)."""
    def __init__(self, r0, r1):
        self.seqs = [r0, r1]
        self.num_seqs = 2
        self.seq_len = len(r0)


def sop_fuzz():
    """Outputs of the reference's own pairwise_alignment_score loop (src/dbga/utils.py:440-504) on seeded random
    two-row alignments -- gap runs that switch rows, both-gap columns, runs of mismatches -- recorded as
    ref_fixtures.json: sop_fuzz.  Pins dbga_b200/quality.py and the device counts (DBGA_FLAG_SOP)."""
    from dbga.utils import pairwise_alignment_score
    rng = random.Random(4404)
    out = []
    for _ in range(300):
        n = rng.choice([1, 2, 5, 17, 64, 200])
        pg = rng.choice([0.0, 0.05, 0.3, 0.7])
        rows = []
        for _r in range(2):
            row, c = [], "A"
            for _i in range(n):
                if rng.random() < pg or (row and row[-1] == "-" and rng.random() < 0.5):
                    c = "-"
                else:
                    c = rng.choice("ACGT")
                row.append(c)
            rows.append("".join(row))
        if rng.random() < 0.5:          # mostly matching rows: copy row 0's bases where row 1 has a base
            rows[1] = "".join(a if (b != "-" and a != "-" and rng.random() < 0.8) else b for a, b in zip(rows[0], rows[1]))
        out.append({"rows": rows, "counts": pairwise_alignment_score(_TwoRows(rows[0], rows[1]))})
    return out


if __name__ == "__main__":
    if "--only-sop" in sys.argv:
        path = os.path.join(ROOT, "tests", "golden", "ref_fixtures.json")
        fx = json.load(open(path))
        fx["sop_fuzz"] = sop_fuzz()
        json.dump(fx, open(path, "w"), indent=1)
        print(f"wrote {len(fx['sop_fuzz'])} sop cases")
    else:
        main()

"""Live diff of the oracle against the reference's unmodified code (cogent3 shimmed).
Runs only where /root/reference exists (the build container); the recorded form of the
same check is tests/golden/ref_graph_fuzz.json."""
import os
import random
import sys

import pytest

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def test_live_reference_fuzz():
    here = os.path.dirname(os.path.abspath(__file__))
    shim = os.path.join(os.path.dirname(here), "oracle", "cogent3_shim")
    saved = list(sys.path)
    sys.path[:0] = [shim, REF]
    try:
        from dbga.debruijn_pairwise import DeBruijnPairwise
        from oracle import dbga_oracle as O
        rng = random.Random(99)
        n_ok = n_cyc = 0
        for _ in range(300):
            n = rng.choice([10, 30, 90, 200])
            k = rng.choice([2, 3, 5, 7])
            s0 = "".join(rng.choice("ACGT") for _ in range(n))
            s1 = "".join(c if rng.random() > 0.08 else rng.choice("ACGT") for c in s0)
            if rng.random() < 0.5:
                cut = rng.randrange(len(s1))
                s1 = s1[:cut] + s1[cut + rng.randint(1, 5):]
            if len(s1) < k:
                continue
            r = O.align_pair(s0, s1, k)
            try:
                db = DeBruijnPairwise([s0, s1], k, "dna")
            except ValueError:
                assert r.status == O.CYCLE
                n_cyc += 1
                continue
            assert r.status == O.OK
            assert db.seq_node_idx == r.graph.seq_node_idx
            assert db.merge_node_idx == r.graph.merge_node_idx
            assert db.id_count == r.graph.id_count
            aln = db.alignment()
            d = aln if isinstance(aln, dict) else aln.to_dict()
            assert (d[0], d[1]) == (r.aln0, r.aln1)
            n_ok += 1
        assert n_ok > 100
    finally:
        sys.path[:] = saved
        for m in [m for m in sys.modules if m == "dbga" or m.startswith("dbga.") or m.startswith("cogent3") or m == "graphviz"]:
            del sys.modules[m]


def test_live_reference_sop():
    """dbga_b200.quality.score_rows against the reference's pairwise_alignment_score loop (src/dbga/utils.py:440-504),
    run live on random two-row alignments (the recorded form: tests/golden/ref_fixtures.json sop_fuzz)."""
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    shim = os.path.join(os.path.dirname(here), "oracle", "cogent3_shim")
    saved = list(sys.path)
    sys.path[:0] = [shim, REF]
    try:
        from dbga.utils import pairwise_alignment_score
        from dbga_b200.quality import score_rows

        class TwoRows:
            def __init__(self, r0, r1):
                self.seqs, self.num_seqs, self.seq_len = [r0, r1], 2, len(r0)

        rng = random.Random(5)
        for _ in range(500):
            n = rng.randint(1, 80)
            pg = rng.choice([0.05, 0.3, 0.6])
            r0 = "".join("-" if rng.random() < pg else rng.choice("ACGT") for _ in range(n))
            r1 = "".join("-" if rng.random() < pg else (a if a != "-" and rng.random() < 0.6 else rng.choice("ACGT")) for a in r0)
            want = pairwise_alignment_score(TwoRows(r0, r1))
            got = score_rows(np.frombuffer(r0.encode(), np.uint8), np.frombuffer(r1.encode(), np.uint8))
            assert got == want, (r0, r1)
    finally:
        sys.path[:] = saved
        for m in [m for m in sys.modules if m == "dbga" or m.startswith("dbga.") or m.startswith("cogent3") or m == "graphviz"]:
            del sys.modules[m]

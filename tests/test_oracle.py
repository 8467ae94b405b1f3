"""The oracle (Python and C restatements) against every golden vector the reference's
own tests hold for the path and against outputs of the reference's
unmodified code recorded by oracle/make_golden.py.  CPU only."""
import itertools
import random

import pytest

from oracle import c_oracle as CO
from oracle import dbga_oracle as O


def test_e2e_goldens_python_oracle(ref_fixtures):
    # reference tests/test_debruijn_pairwise.py:50-190 (k=3, default scores)
    for fx in ref_fixtures["e2e"]:
        r = O.align_pair(fx["seqs"][0], fx["seqs"][1], fx["k"])
        assert r.status == O.OK, fx["fixture"]
        assert [r.graph.seq_node_idx[0], r.graph.seq_node_idx[1]] == fx["seq_node_idx"], fx["fixture"]
        assert r.graph.merge_node_idx == fx["merge_node_idx"], fx["fixture"]
        assert [r.aln0, r.aln1] == fx["aln"], fx["fixture"]


def test_e2e_goldens_c_oracle(ref_fixtures):
    for fx in ref_fixtures["e2e"]:
        r = CO.align_pair(fx["seqs"][0], fx["seqs"][1], fx["k"], want_graph=True)
        assert r.status == O.OK
        assert [r.aln0, r.aln1] == fx["aln"], fx["fixture"]
        assert r.id_count == max(fx["seq_node_idx"][1]) + 1


def test_cycle_goldens(ref_fixtures):
    # reference tests/test_debruijn_pairwise.py:193-214
    for fx in ref_fixtures["cycles"]:
        assert O.align_pair(*fx["seqs"], fx["k"]).status == O.CYCLE
        assert CO.align_pair(*fx["seqs"], fx["k"]).status == O.CYCLE


def test_global_aln_goldens(ref_fixtures):
    # reference tests/test_utils.py:75-95 (empty sides; either/or golden)
    s = O.make_dna_scoring_dict(10, -1, -8)
    for fx in ref_fixtures["global_aln"]:
        got = list(O.dna_global_aln(fx["seqs"][0], fx["seqs"][1], s, 10, 2))
        assert got in fx["accept"]


def test_default_is_only_uniform_order_passing_goldens(ref_fixtures):
    # of the six uniform (pred == end) tie orders only X, M, Y
    # reproduces all ten end-to-end goldens.
    passing = []
    for order in O.ORDERS:
        conv = O.Convention(0, False, order, order)
        ok = all([r.aln0, r.aln1] == fx["aln"]
                 for fx in ref_fixtures["e2e"]
                 for r in [O.align_pair(fx["seqs"][0], fx["seqs"][1], fx["k"], conv=conv)])
        if ok:
            passing.append(order)
    assert passing == ["XMY"]


def test_python_oracle_vs_recorded_reference(ref_fuzz):
    """400 seeded inputs run through the reference's own code (cogent3 shimmed)."""
    ncyc = 0
    for c in ref_fuzz:
        s0, s1 = c["seqs"]
        r = O.align_pair(s0, s1, c["k"])
        ref = c["ref"]
        if "error" in ref:
            assert r.status == O.CYCLE
            ncyc += 1
            continue
        assert r.status == O.OK
        g = r.graph
        assert [g.seq_node_idx[0], g.seq_node_idx[1]] == ref["seq_node_idx"]
        assert g.merge_node_idx == ref["merge_node_idx"]
        assert g.id_count == ref["id_count"]
        assert g.seq_last_kmer_idx == ref["seq_last_kmer_idx"]
        assert [r.aln0, r.aln1] == ref["aln"]
        assert r.dp_segments == len(ref["dp_calls"])
        assert r.dp_cells == sum(len(a) * len(b) for a, b in ref["dp_calls"])
        # the exact list of strings handed to global_pairwise (utils.py:185)
        calls = [[x, y] for kind, x, y in O.segments(s0, s1, g.anchors)
                 if kind not in ("shortcut", "anchor") and x and y]
        assert calls == ref["dp_calls"]
    assert ncyc > 20


def test_c_oracle_vs_recorded_reference(ref_fuzz):
    for c in ref_fuzz:
        s0, s1 = c["seqs"]
        r = CO.align_pair(s0, s1, c["k"], want_graph=True)
        ref = c["ref"]
        if "error" in ref:
            assert r.status == O.CYCLE
            continue
        assert r.status == O.OK
        assert [r.aln0, r.aln1] == ref["aln"]
        assert r.id_count == ref["id_count"]
        assert r.dp_segments == len(ref["dp_calls"])
        # merge node ids follow from the anchors: id = 1 + #non-duplicate k-mers before a
        nd0 = r.nondup0
        ids = [1 + int(nd0[:a].sum()) for a, _ in r.anchors.tolist()]
        assert ids == ref["merge_node_idx"]


def test_c_oracle_equals_python_oracle_all_conventions():
    rng = random.Random(7)
    convs = [O.Convention(g, x, p, e) for g, x, p, e in
             itertools.product((0, 1), (False, True), O.ORDERS, O.ORDERS)]
    for it in range(600):
        L = rng.choice([5, 9, 20, 60, 150])
        k = rng.choice([1, 2, 3, 4, 5, 7, 9])
        s0 = "".join(rng.choice("ACGT") for _ in range(L))
        s1 = list(s0)
        for _ in range(rng.randint(0, max(1, L // 8))):
            if not s1:
                break
            i = rng.randrange(len(s1))
            r = rng.random()
            if r < 0.5:
                s1[i] = rng.choice("ACGT")
            elif r < 0.75:
                del s1[i:i + rng.randint(1, 3)]
            else:
                s1[i:i] = [rng.choice("ACGT") for _ in range(rng.randint(1, 3))]
        s1 = "".join(s1)
        conv = convs[it % len(convs)]
        a = O.align_pair(s0, s1, k, conv=conv)
        b = CO.align_pair(s0, s1, k, conv=conv)
        assert a.status == b.status
        if a.status == O.OK:
            assert (a.aln0, a.aln1, a.score, a.dp_cells) == (b.aln0, b.aln1, b.score, b.dp_cells)
            assert a.aln0.replace("-", "") == s0 and a.aln1.replace("-", "") == s1


def test_status_codes():
    assert O.align_pair("ACGT", "ACGT", 5).status == O.K_INVALID      # utils.py:129-130
    assert CO.align_pair("ACGT", "ACGT", 5).status == O.K_INVALID
    assert O.align_pair("ACGN", "ACGT", 2).status == O.BAD_CHAR
    assert CO.align_pair("ACGN", "ACGT", 2).status == O.BAD_CHAR


def test_batch_oracle_matches_single():
    import numpy as np
    pairs = [O.synth_pair(300, 0.02, 0.004, s) for s in range(40)]
    buf = "".join(a + b for a, b in pairs).encode()
    lens = [len(x) for p in pairs for x in p]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = CO.align_batch(np.frombuffer(buf, np.uint8), offs, 9, threads=4)
    for i, (a, b) in enumerate(pairs):
        r = CO.align_pair(a, b, 9)
        assert out["status"][i] == r.status
        if r.status == O.OK:
            assert out["score"][i] == r.score and out["aln_len"][i] == len(r.aln0)
            assert out["dp_cells"][i] == r.dp_cells


def test_kmer_heuristics_vs_reference_goldens():
    """predict_p / predict_final_p / kmers_larger_than_threshold: the oracle's restatement against
    values recorded from the reference's own functions (oracle/make_golden_kmer.py)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_kmer_heuristics.json")))
    assert len(g["cases"]) >= 20
    for c in g["cases"]:
        for k, p in c["predict_p"].items():
            assert O.predict_p(c["s0"], c["s1"], int(k)) == p
        assert O.predict_final_p(c["s0"], c["s1"]) == c["predict_final_p"]
        for key, v in c["dup_prop_gt"].items():
            k, th = key.split(":")
            assert O.kmers_larger_than_threshold(c["s0"], int(k), float(th)) == v
    with pytest.raises(ValueError, match="Invalid k size"):
        O.kmer_stats("ACGT", "ACG", 4)

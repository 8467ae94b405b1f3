"""The reference's own pairwise tests (tests/test_debruijn_pairwise.py:7-214) re-expressed
against the drop-in class; expectations come from tests/golden/ref_fixtures.json, i.e. from
the reference's test file.  Needs a B200."""
import os

import pytest

pytestmark = pytest.mark.gpu


def example_checker(seqs, k, exp_seq1_idx, exp_seq2_idx, exp_merge, exp_aln):
    from dbga_b200 import DeBruijnPairwise, NodeType
    exp_seq1 = exp_aln[0].replace("-", "")
    exp_seq2 = exp_aln[1].replace("-", "")
    db = DeBruijnPairwise(seqs, k, moltype="dna")
    for node in db.nodes.values():
        if node.node_type not in [NodeType.start, NodeType.end]:
            assert len(node.kmer) == db.k
    exp_start_idx = exp_seq1_idx[0], exp_seq2_idx[0]
    exp_end_idx = exp_seq1_idx[-1], exp_seq2_idx[-1]
    assert db.sequences[0] == exp_seq1
    assert db.sequences[1] == exp_seq2
    assert db.seq_node_idx[0] == exp_seq1_idx
    assert db.seq_node_idx[1] == exp_seq2_idx
    for i in range(db.id_count):
        if i in exp_start_idx:
            assert db.nodes[i].node_type is NodeType.start
        elif i in exp_end_idx:
            assert db.nodes[i].node_type is NodeType.end
        else:
            assert db.nodes[i].node_type is NodeType.middle
    aln_result = db.alignment()
    assert aln_result == {f"seq{i + 1}": exp_aln[i] for i in range(db.num_seq)}
    assert db.merge_node_idx == exp_merge
    return db, aln_result


def test_reference_examples(data_dir, ref_fixtures):
    for fx in ref_fixtures["e2e"]:
        db, aln = example_checker(os.path.join(data_dir, fx["fixture"] + ".fasta"), fx["k"],
                                  fx["seq_node_idx"][0], fx["seq_node_idx"][1], fx["merge_node_idx"],
                                  tuple(fx["aln"]))
        if not fx["merge_node_idx"]:
            assert isinstance(aln, dict)          # debruijn_pairwise.py:453-457
        else:
            assert hasattr(aln, "to_fasta") and aln.to_dict() == dict(zip(["seq1", "seq2"], fx["aln"]))


def test_cycle_errors(ref_fixtures):
    from dbga_b200 import DeBruijnPairwise
    for fx in ref_fixtures["cycles"]:
        with pytest.raises(ValueError) as e:
            DeBruijnPairwise(fx["seqs"], k=fx["k"], moltype="dna")
        assert str(e.value) == "Cycles detected in de Bruijn graph, usually caused by small kmer sizes"


def test_error_behaviour():
    from dbga_b200 import DeBruijnPairwise
    with pytest.raises(ValueError, match="Invalid k size for kmers"):
        DeBruijnPairwise(["ACGT", "ACGTAC"], k=5, moltype="dna")
    with pytest.raises(AssertionError):
        DeBruijnPairwise(["ACGT", "ACGT", "ACGT"], k=2, moltype="dna")
    with pytest.raises(ValueError, match="Invalid input for sequence argument"):
        DeBruijnPairwise({1: "ACGT"}, k=2, moltype="dna")


def test_graph_views_match_recorded_reference(ref_fuzz):
    """nodes / expansion / extract_bubble_seq / seq_last_kmer_idx as the reference builds them
    (recorded by oracle/make_golden.py from the reference's own code)."""
    from dbga_b200 import DeBruijnPairwise
    n = 0
    for c in ref_fuzz[:150]:
        ref = c["ref"]
        if "error" in ref:
            with pytest.raises(ValueError):
                DeBruijnPairwise(c["seqs"], c["k"], "dna")
            continue
        db = DeBruijnPairwise(c["seqs"], c["k"], "dna")
        assert [db.seq_node_idx[0], db.seq_node_idx[1]] == ref["seq_node_idx"]
        assert db.merge_node_idx == ref["merge_node_idx"]
        assert db.id_count == ref["id_count"] == len(db.nodes)
        assert db.seq_last_kmer_idx == ref["seq_last_kmer_idx"]
        # the text of the whole node path of each sequence is the sequence itself
        for si in (0, 1):
            assert db.extract_bubble_seq(db.seq_node_idx[si], si) == db.sequences[si]
        aln = db.alignment()
        d = aln if isinstance(aln, dict) else aln.to_dict()
        assert [d[0], d[1]] == ref["aln"]
        assert isinstance(aln, dict) == ref["returns_dict"]
        n += 1
    assert n > 80


def test_dna_global_aln_goldens(ref_fixtures):
    # reference tests/test_utils.py:75-95
    from dbga_b200 import dna_global_aln, make_dna_scoring_dict
    s = make_dna_scoring_dict(10, -1, -8)
    for fx in ref_fixtures["global_aln"]:
        got = list(dna_global_aln(fx["seqs"][0], fx["seqs"][1], s=s, d=10, e=2))
        assert got in fx["accept"]
    assert dna_global_aln("", "", s=s, d=10, e=2) == ("", "")


def test_cli_end_to_end(tmp_path, data_dir):
    from click.testing import CliRunner
    from dbga_b200.cli import main
    out = tmp_path / "out.fasta"
    res = CliRunner().invoke(main, ["-i", os.path.join(data_dir, "gap_bubble_duplicate.fasta"), "-o", str(out),
                                    "-k", "3", "-m", "dna"])
    assert res.exit_code == 0, res.output
    assert "Number of sequences: 2, running de Bruijn pairwise alignment" in res.output
    assert "Alignment task finishes within" in res.output
    assert out.read_text() == ">seq1\nGTACACGTATG\n>seq2\nGTACACG-ATG\n"
    # no shared k-mer: the reference would crash on dict.to_fasta(); we still write FASTA
    res = CliRunner().invoke(main, ["-i", os.path.join(data_dir, "unrelated_sequences.fasta"), "-o", str(out),
                                    "-k", "3", "-m", "dna", "--match", "10", "-d", "10", "-e", "2"])
    assert res.exit_code == 0, res.output
    assert out.read_text() == ">seq1\nACGT--AGTATC\n>seq2\nACCTACAGCA--\n"


def test_baseline_config2_is_a_cycle_error():
    """BASELINE config [1]: single 10 kb pair, 1 % + 0.2 % divergence, k=7: the reference
    raises the cycles ValueError for every seed; at k=11 most seeds align."""
    from dbga_b200 import DeBruijnPairwise
    from oracle import dbga_oracle as O
    from oracle import c_oracle as CO
    s0, s1 = O.synth_pair(10000, 0.01, 0.002, 1)
    with pytest.raises(ValueError, match="Cycles detected"):
        DeBruijnPairwise([s0, s1], 7, "dna")
    for seed in range(1, 6):
        s0, s1 = O.synth_pair(10000, 0.01, 0.002, seed)
        ref = CO.align_pair(s0, s1, 11)
        if ref.status != 0:
            continue
        aln = DeBruijnPairwise([s0, s1], 11, "dna").alignment()
        assert aln == {0: ref.aln0, 1: ref.aln1}
        break
    else:
        pytest.fail("no colinear seed found")


def test_adaptive_recursive_goldens(ref_fixtures):
    # reference tests/test_adaptive_debruijn.py:19-48, plus recorded outputs of the reference's
    # own adpt_dbg_alignment_recursive on random related pairs (oracle/make_golden.py)
    from dbga_b200 import make_dna_scoring_dict
    from dbga_b200.adaptive_debruijn import adpt_dbg_alignment_recursive
    s = make_dna_scoring_dict(10, -1, -8)
    for fx in ref_fixtures["adaptive"]:
        got = adpt_dbg_alignment_recursive(fx["seqs"], 0, tuple(fx["k_choice"]), s, 10, 2)
        assert list(got) == fx["aln"]
    n_ok = 0
    for fx in ref_fixtures["adaptive_fuzz"]:
        if "error" in fx:
            with pytest.raises(Exception):
                adpt_dbg_alignment_recursive(tuple(fx["seqs"]), 0, tuple(fx["k_choice"]), s, 10, 2)
            continue
        got = adpt_dbg_alignment_recursive(tuple(fx["seqs"]), 0, tuple(fx["k_choice"]), s, 10, 2)
        assert list(got) == fx["aln"]
        n_ok += 1
    assert n_ok >= 10


def test_k_zero_means_no_anchors():
    """k = 0 is legal in the reference (utils.py:129: only k < 0 or k > len is rejected): every
    k-mer is the empty string, hence a duplicate; no merge node; plain global alignment."""
    from dbga_b200 import DeBruijnPairwise
    db = DeBruijnPairwise(["ACGTAGTATC", "ACCTACAGCA"], 0, "dna")
    assert db.merge_node_idx == [] and db.seq_node_idx == {0: [0, 1], 1: [2, 3]} and db.id_count == 4
    assert db.alignment() == {0: "ACGT--AGTATC", 1: "ACCTACAGCA--"}


def test_batch_cli_end_to_end(tmp_path):
    """dbga-batch: a FASTA file of many pairs in, the reference CLI's per-pair output (cli.py:131-136: both
    rows, 60-column records) out for every pair it would align; native reader, 2-bit upload, compact result,
    native writer.  One and two shards (two contexts on one GPU box)."""
    import numpy as np
    from click.testing import CliRunner
    from dbga_b200 import Aligner
    from dbga_b200.cli_batch import main as batch_main
    from oracle import dbga_oracle as O
    pairs = [O.synth_pair(700, 0.02, 0.004, 9000 + s) for s in range(120)]
    pairs[5] = ("ACGTTGCAAC", "ACGTAGCAAC")
    src = tmp_path / "in.fasta"
    with open(src, "w") as f:
        for i, (a, b) in enumerate(pairs):
            for nm, s in ((f"s{i}_a some description", a), (f"s{i}_b", b)):
                f.write(f">{nm}\n" + "".join(s[j:j + 70].lower() + "\n" for j in range(0, len(s), 70)))
    ref = Aligner(0).align_pairs(pairs, 9)
    want = ""
    for i in range(len(pairs)):
        if ref.status[i] != 0:
            continue
        a0, a1 = ref.alignment(i)
        for nm, row in ((f"s{i}_a", a0), (f"s{i}_b", a1)):
            want += f">{nm}\n" + "".join(row[j:j + 60] + "\n" for j in range(0, len(row), 60))
    assert (ref.status == 0).sum() > 60 and (ref.status == 1).sum() > 0
    for gpus in (1,):
        out = tmp_path / f"out{gpus}.fasta"
        r = CliRunner().invoke(batch_main, ["-i", str(src), "-o", str(out), "-k", "9", "--gpus", str(gpus)])
        assert r.exit_code == 0, r.output
        assert "pairs written" in r.output and "not aligned: cycles" in r.output
        assert out.read_text() == want
    # two shards through the multi-GPU entry (both contexts on device 0 when the box has one GPU)
    from dbga_b200 import MultiAligner, _lib
    from dbga_b200.fasta import FastaBatch
    import torch
    fa = FastaBatch(src)
    m = MultiAligner([0, 1] if torch.cuda.device_count() > 1 else [0, 0])
    res = m.align_buffers(None, np.ascontiguousarray(fa.offsets), Aligner.params(9, flags=_lib.FLAG_COMPACT_ROWS),
                          packed=fa.packed_ptr, unpack=False)
    out = tmp_path / "out_multi.fasta"
    n = sum(fa.write(out, res.shard[s], first_record=2 * int(res.pair_lo[s]), append=s > 0) for s in range(res.num_shards))
    assert n == int((ref.status == 0).sum()) and out.read_text() == want
    m.close(); fa.close()


def test_adaptive_top_level_goldens_and_call_count(ref_fixtures):
    """adpt_dbg_alignment (reference adaptive_debruijn.py:17-88) on pairs long enough to pass its similarity and
    entropy gates, against FASTA text recorded from the reference's own function (oracle/make_golden.py; rows in
    60-column blocks) -- and the level-synchronous descent makes a C-ABI call per level, not per bubble."""
    from dbga_b200 import adaptive_debruijn as AD, batch
    al = batch.default_aligner(0)
    calls = {"n": 0}
    real = al.align_buffers

    def counting(*a, **kw):
        calls["n"] += 1
        return real(*a, **kw)
    al.align_buffers = counting
    try:
        for fx in ref_fixtures["adaptive_top"]:
            calls["n"] = 0
            if "error" in fx:
                with pytest.raises(Exception):
                    AD.adpt_dbg_alignment(list(fx["seqs"]), "dna")
                continue
            got = AD.adpt_dbg_alignment(list(fx["seqs"]), "dna")
            assert got == fx["fasta"]
            assert max(len(l) for l in got.splitlines()) <= 60
            # k runs over at most ~6 odd values; kmer statistics + one call per level + one for the leaves
            assert calls["n"] <= 16, calls["n"]
    finally:
        al.align_buffers = real


def test_adaptive_batch_of_pairs_equals_one_by_one():
    from dbga_b200 import make_dna_scoring_dict
    from dbga_b200.adaptive_debruijn import adpt_dbg_alignment_batch, adpt_dbg_alignment_recursive
    from oracle import dbga_oracle as O
    s = make_dna_scoring_dict(10, -1, -8)
    pairs = [O.synth_pair(600, 0.03, 0.01, 500 + i) for i in range(24)]
    kc = (13, 9, 5)
    one, ok = [], []
    for p in pairs:
        try:
            one.append(adpt_dbg_alignment_recursive(p, 0, kc, s, 10, 2)); ok.append(p)
        except ValueError:
            pass
    assert len(ok) >= 8
    assert adpt_dbg_alignment_batch(ok, 0, kc, s, 10, 2) == one

"""CPU-side tests: host logic, the C-ABI library's exports, error behaviour without a GPU,
and the 2-rank sharding logic over gloo.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from dbga_b200 import build
    return build.build_library()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "dbga_b200.h")).read()
    declared = set(re.findall(r"\b(dbga_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(lib, name), name
    from dbga_b200 import _lib
    assert declared == set(_lib.EXPORTS)
    assert lib.dbga_version() >= 100


def test_struct_layouts_match_header(built, tmp_path):
    """sizeof / offsetof of the header's structs as gcc sees them == the ctypes mirrors."""
    import subprocess
    from dbga_b200 import _lib
    fields = {"dbga_params": _lib.Params, "dbga_result": _lib.Result, "dbga_multi_result": _lib.MultiResult}
    lines = []
    for cname, ct in fields.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in ct._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dbga_b200.h"\nint main(void){' + "".join(lines) + "return 0;}")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in fields.items():
        assert int(got[cname]) == ctypes.sizeof(ct), cname
        for f, _ in ct._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(ct, f).offset, (cname, f)


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dbga_b200 import Aligner
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Aligner(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dbga_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(from|import)\s+oracle|liboracle|oracle_[a-z]|c_oracle", text), f


def test_utils_mirror_reference_behaviour():
    from dbga_b200 import utils
    # reference tests/test_utils.py:20-37
    assert utils.read_nucleotide_from_kmers("", 3) == ""
    assert utils.read_nucleotide_from_kmers("AOPKFNLSDOEPKPOWSEGH", 4) == "AFDKS"
    assert utils.read_nucleotide_from_kmers("1092831790", 2) == "19819"
    with pytest.raises(AssertionError):
        utils.read_nucleotide_from_kmers("ACMDANFGL", 4)
    # :113-145
    assert utils.get_kmers("GTACACGTATG", 3) == ["GTA", "TAC", "ACA", "CAC", "ACG", "CGT", "GTA", "TAT", "ATG"]
    with pytest.raises(ValueError) as e:
        utils.get_kmers("TACCACGTAAT", 16)
    assert e.value.args[0] == "Invalid k size for kmers"
    # :148-155
    assert utils.duplicate_kmers([["AG", "CT", "CA", "AC", "CT"], ["CT", "AC", "GT", "AG"]]) == {"CT"}
    assert utils.duplicate_kmers([["AG", "CT", "CA", "AG", "CT"], ["CT", "AG", "CT", "AG", "AC"]]) == {"AG", "CT"}
    # :98-110
    assert utils.debruijn_merge_correctness([[0, [1, 2], 3], [0, 3]]) is False
    assert utils.debruijn_merge_correctness([[0, [1, 2], 3], [0, 4, 3]]) is False
    assert utils.debruijn_merge_correctness([[0, [1, 2], 3], [0, [6, 7], 4]]) is False
    assert utils.debruijn_merge_correctness([[0, [1, 2], 3], [0, [6, 7], 3]]) is True
    assert utils.get_closest_odd(3, True) == 3 and utils.get_closest_odd(4, True) == 5 and utils.get_closest_odd(4, False) == 3


def test_load_sequences(data_dir):
    from dbga_b200 import utils, seqcoll
    sc = utils.load_sequences(os.path.join(data_dir, "gap_at_end.fasta"), moltype="dna")
    assert [str(s) for s in sc.seqs] == ["GTACAAGCGATG", "GTACAAGCGA"] and sc.names == ["seq1", "seq2"]
    sc = utils.load_sequences(["ACGTCAG", "GACTGTGA"], moltype="dna")
    assert sc.names == [0, 1]
    assert utils.load_sequences(sc, moltype="dna") is sc
    with pytest.raises(ValueError) as e:
        utils.load_sequences({1: "ACGTGTA"}, moltype="dna")
    assert e.value.args[0] == "Invalid input for sequence argument"
    with pytest.raises(ValueError):
        utils.load_sequences("data/not_a_file", moltype="dna")
    with pytest.raises(seqcoll.AlphabetError):
        utils.load_sequences(["ACTGA", "TACGE"], moltype="dna")


def test_alignment_container():
    from dbga_b200 import seqcoll
    aln = seqcoll.make_aligned_seqs({"seq1": "AC-T", "seq2": "ACGT"}, moltype="dna")
    assert aln == {"seq1": "AC-T", "seq2": "ACGT"}
    assert aln.to_fasta() == ">seq1\nAC-T\n>seq2\nACGT\n"
    assert aln.to_dict() == {"seq1": "AC-T", "seq2": "ACGT"} and aln.seq_len == 4


def test_pack_pairs_and_synth():
    from dbga_b200.batch import pack_pairs
    from dbga_b200.synth import synth_batch, pairs_from_buffers
    buf, offs = pack_pairs([("ACGT", "AC"), ("", "GGG")])
    assert buf.tobytes() == b"ACGTACGGG" and offs.tolist() == [0, 4, 6, 6, 9]
    buf, offs = synth_batch(50, 200, 0.02, 0.01, seed=3)
    assert len(offs) == 101 and offs[-1] == len(buf)
    pairs = pairs_from_buffers(buf, offs, range(50))
    assert all(len(a) == 200 and set(a + b) <= set("ACGT") for a, b in pairs)
    assert any(len(b) != 200 for _, b in pairs) and any(a != b for a, b in pairs)
    b2, o2 = synth_batch(50, 200, 0.02, 0.01, seed=3)
    assert np.array_equal(buf, b2) and np.array_equal(offs, o2)


def test_batch_result_anchor_expansion():
    from dbga_b200.batch import BatchResult
    z = np.zeros(0)
    r = BatchResult(status=np.zeros(2, np.int32), score=z, dp_cells=z, dp_segments=z, n_anchors=z,
                    run_off=np.array([0, 2, 2]), runs=np.array([[3, 1, 3], [9, 8, 1]], np.int32),
                    aln_off=np.array([0, 0, 0]), aln0=np.zeros(0, np.uint8), aln1=np.zeros(0, np.uint8))
    assert r.anchors(0).tolist() == [[3, 1], [4, 2], [5, 3], [9, 8]]
    assert r.anchors(1).shape == (0, 2)


def test_cli_rejects_non_fasta(tmp_path):
    from click.testing import CliRunner
    from dbga_b200.cli import main
    p = tmp_path / "in.txt"
    p.write_text(">a\nACGT\n>b\nACGT\n")
    res = CliRunner().invoke(main, ["-i", str(p), "-o", str(tmp_path / "o.fasta"), "-k", "3", "-m", "dna"])
    assert res.exit_code == 0 and "Expect input and output files to be `fasta` format!" in res.output
    one = tmp_path / "one.fasta"
    one.write_text(">a\nACGT\n")
    res = CliRunner().invoke(main, ["-i", str(one), "-o", str(tmp_path / "o.fasta"), "-k", "3", "-m", "dna"])
    assert "Too few sequences for alignment, should be at least 2!" in res.output


SHARD_SCRIPT = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from dbga_b200.sharding import shard_range, gather_results
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
lo, hi = shard_range(11, rank, 2)
local = [("pair%d" % i, i * i) for i in range(lo, hi)]
allr = gather_results(local)
if rank == 0:
    assert [x[0] for x in allr] == ["pair%d" % i for i in range(11)], allr
    assert [x[1] for x in allr] == [i * i for i in range(11)]
    print("SHARD_OK")
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(tmp_path):
    script = tmp_path / "shard.py"
    script.write_text(SHARD_SCRIPT.format(root=ROOT, port=29611))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0][0]


GATHER_SCRIPT = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from dbga_b200.batch import BatchResult
from dbga_b200.sharding import shard_range, gather_batch_result
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()

def shard(lo, hi):
    # pair i: i % 5 aligned columns of byte 65 + i % 20 / 97 + i % 20, i % 3 runs (i, i + 1, 2 + r), sop row (i, 0, 1, 2)
    lens = np.array([i % 5 for i in range(lo, hi)], np.int64)
    nr = np.array([i % 3 for i in range(lo, hi)], np.int64)
    off = lambda l: np.concatenate([[0], np.cumsum(l)]).astype(np.int64)
    a0 = np.concatenate([np.full(i % 5, 65 + i % 20, np.uint8) for i in range(lo, hi)] + [np.zeros(0, np.uint8)])
    a1 = np.concatenate([np.full(i % 5, 97 + i % 20, np.uint8) for i in range(lo, hi)] + [np.zeros(0, np.uint8)])
    runs = np.array([[i, i + 1, 2 + r] for i in range(lo, hi) for r in range(i % 3)], np.int32).reshape(-1, 3)
    ids = np.arange(lo, hi)
    return BatchResult(status=(ids % 2).astype(np.int32), score=ids * 7, dp_cells=ids * 11, dp_segments=ids % 4, n_anchors=ids * 3,
                       run_off=off(nr), runs=runs, aln_off=off(lens), aln0=a0, aln1=a1,
                       sop=np.stack([ids, 0 * ids, 0 * ids + 1, 0 * ids + 2], 1).astype(np.int32))

P = 23
lo, hi = shard_range(P, rank, 2)
got = gather_batch_result(shard(lo, hi))
if rank == 0:
    want = shard(0, P)
    for f in ("status", "score", "dp_cells", "dp_segments", "n_anchors", "run_off", "runs", "aln_off", "aln0", "aln1", "sop"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    assert got.alignment(9) == want.alignment(9) and np.array_equal(got.anchors(8), want.anchors(8))
    print("GATHER_OK")
else:
    assert got is None
dist.destroy_process_group()
"""


def test_two_rank_array_gather_gloo(tmp_path):
    """gather_batch_result: the shards' BatchResults joined on rank 0 as arrays (offsets rebased), world size 2 over gloo."""
    script = tmp_path / "gather.py"
    script.write_text(GATHER_SCRIPT.format(root=ROOT, port=29613))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0][0]


def test_shard_range_partitions():
    from dbga_b200.sharding import shard_range
    for n in (0, 1, 7, 100_000):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_fasta_batch_reader_and_writer(tmp_path):
    from dbga_b200 import fasta
    from dbga_b200.batch import BatchResult
    p = tmp_path / "pairs.fasta"
    p.write_text(">a1 desc\nACGT\nacg\n>b1\nAC\r\nGT\n\n>a2\n\n>b2\nTTTT")
    names, buf, offs = fasta.read_pairs(p)
    assert names == ["a1", "b1", "a2", "b2"]
    assert buf.tobytes() == b"ACGTACGACGTTTTT" and offs.tolist() == [0, 7, 11, 11, 15]
    odd = tmp_path / "odd.fasta"
    odd.write_text(">x\nAC\n")
    with pytest.raises(ValueError):
        fasta.read_pairs(odd)
    z = np.zeros(0)
    res = BatchResult(status=np.array([0, 1], np.int32), score=z, dp_cells=z, dp_segments=z, n_anchors=z,
                      run_off=np.zeros(3, np.int64), runs=np.zeros((0, 3), np.int32), aln_off=np.array([0, 5, 5]),
                      aln0=np.frombuffer(b"AC-GT", np.uint8), aln1=np.frombuffer(b"ACGGT", np.uint8))
    out = tmp_path / "aln.fasta"
    assert fasta.write_alignments(out, names, res, width=3) == 1
    assert out.read_text() == ">a1\nAC-\nGT\n>b1\nACG\nGT\n"


def test_sop_reference_goldens():
    # reference tests/test_utils.py:158-215
    from dbga_b200.quality import sop, pairwise_alignment_score
    assert sop({"seq1": "AGTCCAGTGA", "seq2": "A--C-ATGGA"}) == {"match": 5, "mismatch": 2, "gap_open": 2, "gap_extend": 1}
    assert sop({"seq1": "AACTTCTTCGTGT", "seq2": "AAC--C--CGTGT"}) == {"match": 9, "mismatch": 0, "gap_open": 2, "gap_extend": 2}
    assert sop({"seq1": "AACTTCTTCGTGT", "seq2": "AAC--C--CGTGT", "seq3": "AAC-----CGTGT"}) == \
        {"match": 25, "mismatch": 0, "gap_open": 4, "gap_extend": 6}
    assert sop({"seq1": "AAGTCGATGGTCA", "seq2": "AAC-CCA---TGA", "seq3": "ATCT---T--CCA"}) == \
        {"match": 14, "mismatch": 9, "gap_open": 7, "gap_extend": 7}
    assert pairwise_alignment_score({"a": "A-C", "b": "AG-"}) == {"match": 1, "mismatch": 0, "gap_open": 2, "gap_extend": 0}


def test_sop_matches_reference_loop():
    """dbga_b200.quality against 300 outputs recorded from the reference's own pairwise_alignment_score loop
    (src/dbga/utils.py:440-504, run unmodified by oracle/make_golden.py: sop_fuzz) on random two-row
    alignments: gap runs that switch rows, both-gap columns, runs of mismatches.  Plus the column identity
    match + mismatch + gap_open + gap_extend + both-gap columns == columns."""
    import json
    from dbga_b200.quality import score_rows
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fixtures.json")))["sop_fuzz"]
    assert len(cases) == 300
    for c in cases:
        r1 = np.frombuffer(c["rows"][0].encode(), np.uint8)
        r2 = np.frombuffer(c["rows"][1].encode(), np.uint8)
        s = score_rows(r1, r2)
        assert s == c["counts"], c["rows"]
        assert sum(s.values()) + int(((r1 == 45) & (r2 == 45)).sum()) == len(r1)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs on the host alone and prints exactly one JSON line with
    the contract's keys (the GPU arm prints the same keys plus roofline / clocks / gpu_launches)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-sample", "300"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_native_fasta_reader_and_writer(tmp_path, built):
    """dbga_fasta_read / dbga_fasta_write (csrc/fasta_io.cu) without a GPU: labels, upper-casing, line breaks,
    the canonical 2-bit words, and 60-column records written from a full and from a COMPACT_ROWS result."""
    import numpy as np
    from dbga_b200 import _lib, fasta, pack_ascii
    p = tmp_path / "pairs.fasta"
    long0, long1 = "ACGT" * 40 + "AC", "ACGT" * 40 + "AG"
    p.write_text(f">a1 desc\nACGT\nacg\n>b1\nAC\r\nGT\n\n>a2\n{long0[:70]}\n{long0[70:]}\n>b2\n{long1}")
    fa = fasta.FastaBatch(p)
    assert fa.names == ["a1", "b1", "a2", "b2"] and fa.num_pairs == 2 and fa.bad_records == 0
    assert fa.seqs.tobytes() == ("ACGTACG" + "ACGT" + long0 + long1).encode()
    assert fa.offsets.tolist() == [0, 7, 11, 11 + len(long0), 11 + len(long0) + len(long1)]
    words = pack_ascii(fa.seqs, fa.offsets)
    got = np.ctypeslib.as_array(ctypes.cast(fa.packed_ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(len(words),))
    assert np.array_equal(got[:-1], words[:-1])
    nbad = tmp_path / "n.fasta"
    nbad.write_text(">x\nACGN\n>y\nAC\n")
    assert fasta.FastaBatch(nbad).bad_records == 1
    # a hand-made result: pair 0 aligned (full rows), pair 1 a cycle
    def result(status, aln_off, a0, a1, runs=None, run_off=None, seg=None):
        r = _lib.Result()
        keep = [np.array(status, np.int32), np.array(aln_off, np.int64), np.frombuffer(a0, np.uint8).copy(), np.frombuffer(a1, np.uint8).copy()]
        r.num_pairs = len(status)
        r.status = keep[0].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)); r.aln_off = keep[1].ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
        r.aln0 = keep[2].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)); r.aln1 = keep[3].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        if runs is not None:
            keep += [np.array(runs, np.int32), np.array(run_off, np.int64), np.array(seg, np.int32)]
            r.runs = keep[4].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)); r.run_off = keep[5].ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
            r.seg_cols = keep[6].ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        return r, keep
    out = tmp_path / "out.fasta"
    r, keep = result([0, 1], [0, 8, 8], b"ACGTACG-", b"AC-GT---")
    assert fa.write(out, r) == 1
    assert out.read_text() == ">a1\nACGTACG-\n>b1\nAC-GT---\n"
    # pair 1 = (long0, long1) as a COMPACT_ROWS result: one run of 161 anchors, then the tail segment AC / AG
    # (the last merge node is left to the tail: 160 run columns + 2 tail columns)
    r, keep = result([1, 0], [0, 0, 2], b"AC", b"AG", runs=[0, 0, 161], run_off=[0, 0, 1], seg=[0, 0, 2])
    assert fa.write(out, r) == 1
    want = "".join(f">{n}\n" + "".join(s[i:i + 60] + "\n" for i in range(0, len(s), 60)) for n, s in (("a2", long0), ("b2", long1)))
    assert out.read_text() == want
    fa.close()

"""What travels over the host link: the NO_RUNS / COMPACT_ROWS result forms, the
2-bit packed input entry, the one-call multi-GPU entry, and the alignment-quality counts
(pairwise_alignment_score, reference src/dbga/utils.py:440-504) the assembler emits.  Every form
must reproduce, byte for byte, what the plain dbga_align_batch call returns.  Needs a B200."""
import numpy as np
import pytest

from oracle import dbga_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aligner():
    from dbga_b200 import Aligner
    a = Aligner(0)
    yield a
    a.close()


@pytest.fixture(scope="module")
def batch():
    """~24 MB: large enough for the chunked pipeline, mixed with a few odd pairs."""
    from dbga_b200.batch import pack_pairs
    pairs = [O.synth_pair(1500, 0.018, 0.002, 31000 + s) for s in range(8000)]
    pairs[17] = ("ACGTACGTAC", "ACGTACGTAC")
    pairs[18] = (pairs[18][0], pairs[18][0][:40])               # very unbalanced
    pairs[4000] = O.synth_pair(9000, 0.01, 0.002, 5)
    buf, offs = pack_pairs(pairs)
    return pairs, buf, offs


def _same_rows(a, b):
    assert np.array_equal(a.status, b.status) and np.array_equal(a.score, b.score)
    assert np.array_equal(a.aln_off, b.aln_off) and np.array_equal(a.aln0, b.aln0) and np.array_equal(a.aln1, b.aln1)
    assert np.array_equal(a.dp_cells, b.dp_cells) and np.array_equal(a.n_anchors, b.n_anchors)


def test_sop_counts_reference_goldens(aligner, ref_fixtures):
    """reference tests/test_utils.py:158-215 (sop) on alignments produced by the device: the counts
    of the rows the assembler wrote equal the host restatement (dbga_b200.quality) of the
    reference's loop; the reference's own sop goldens are checked against that restatement in
    tests/test_host_api.py."""
    from dbga_b200 import quality
    pairs = [O.synth_pair(400, 0.05, 0.03, 100 + s) for s in range(300)] + [tuple(fx["seqs"]) for fx in ref_fixtures["e2e"]]
    from dbga_b200 import _lib
    assert aligner.align_pairs(pairs[:4], 3).sop is None            # only on request
    for k in (3, 7):
        res = aligner.align_pairs(pairs, k, flags=_lib.FLAG_SOP)
        assert res.sop is not None and res.sop.shape == (len(pairs), 4)
        n = 0
        for i in range(len(pairs)):
            if res.status[i] != O.OK:
                assert not res.sop[i].any()
                continue
            lo, hi = int(res.aln_off[i]), int(res.aln_off[i + 1])
            want = quality.score_rows(res.aln0[lo:hi], res.aln1[lo:hi])
            assert res.sop[i].tolist() == [want["match"], want["mismatch"], want["gap_open"], want["gap_extend"]], (k, i)
            n += 1
        assert n > 50


def test_sop_counts_in_a_pipelined_batch(aligner, batch):
    from dbga_b200 import _lib, quality
    pairs, buf, offs = batch
    res = aligner.align_buffers(buf, offs, aligner.params(9, flags=_lib.FLAG_SOP))
    assert res.timing["launches"] > 40                            # several chunks
    for i in list(range(0, len(pairs), 211)) + [17, 18, 4000]:
        if res.status[i] != O.OK:
            continue
        lo, hi = int(res.aln_off[i]), int(res.aln_off[i + 1])
        want = quality.score_rows(res.aln0[lo:hi], res.aln1[lo:hi])
        assert res.sop[i].tolist() == list(want.values()), i
    # whole-batch identity: every column is exactly one of the four kinds
    ok = res.status == O.OK
    assert np.array_equal(res.sop[ok].sum(axis=1), np.diff(res.aln_off)[ok])
    # the same counts come with the compact result form
    comp = aligner.align_buffers(buf, offs, aligner.params(9, flags=_lib.FLAG_SOP | _lib.FLAG_COMPACT_ROWS))
    assert np.array_equal(comp.sop, res.sop)


def test_sop_long_segments_and_conventions(aligner):
    """Slots of more than 64 columns are classified by the whole warp; gap-rich alignments under a
    second convention; a 200 kb pair with multi-kb bubbles."""
    from dbga_b200 import _lib, quality
    from dbga_b200.synth import synth_long_pair
    import random
    rng = random.Random(77)
    pairs = []
    for _ in range(40):
        n = rng.choice([300, 900, 2500])
        x = "".join(rng.choice("ACGT") for _ in range(n))
        y = list(x)
        for _ in range(rng.randint(1, 4)):                                    # long indels and diverged windows
            p = rng.randrange(20, n - 120)
            if rng.random() < 0.5:
                del y[p:p + rng.randint(5, 90)]
            else:
                y[p:p] = [rng.choice("ACGT") for _ in range(rng.randint(5, 90))]
        pairs.append((x, "".join(y)))
    for kw in ({}, dict(gap_len_mode=1, xy_allowed=True, pred_order="MXY", end_order="XYM")):
        res = aligner.align_pairs(pairs, 11, flags=_lib.FLAG_SOP, **kw)
        n = 0
        for i in range(len(pairs)):
            if res.status[i] != O.OK:
                continue
            lo, hi = int(res.aln_off[i]), int(res.aln_off[i + 1])
            assert res.sop[i].tolist() == list(quality.score_rows(res.aln0[lo:hi], res.aln1[lo:hi]).values()), (i, kw)
            n += 1
        assert n > 10
    for seed in range(1, 30):
        buf, offs = synth_long_pair(n=200_000, n_bubbles=4, seed=seed)
        res = aligner.align_buffers(buf, offs, aligner.params(15, flags=_lib.FLAG_SOP))
        if res.status[0] == O.OK:
            assert res.sop[0].tolist() == list(quality.score_rows(res.aln0, res.aln1).values())
            break


def test_no_runs_flag(aligner, batch):
    from dbga_b200 import _lib
    pairs, buf, offs = batch
    full = aligner.align_buffers(buf, offs, aligner.params(9))
    lean = aligner.align_buffers(buf, offs, aligner.params(9, flags=_lib.FLAG_NO_RUNS))
    _same_rows(full, lean)
    assert len(lean.runs) == 0 and not lean.run_off.any() and len(full.runs) > 0
    small = pairs[:50]
    a = aligner.align_pairs(small, 9)
    b = aligner.align_pairs(small, 9, flags=_lib.FLAG_NO_RUNS)          # direct (unchunked) path
    _same_rows(a, b)
    assert len(b.runs) == 0


@pytest.mark.parametrize("count", [60, 8000])
def test_compact_rows_expand_to_the_full_rows(aligner, batch, count):
    from dbga_b200 import _lib
    from dbga_b200.batch import pack_pairs
    pairs, buf, offs = batch
    buf, offs = pack_pairs(pairs[:count]) if count < len(pairs) else (buf, offs)
    full = aligner.align_buffers(buf, offs, aligner.params(9))
    raw = aligner.align_buffers(buf, offs, aligner.params(9, flags=_lib.FLAG_COMPACT_ROWS), unpack=False)
    comp = aligner._unpack(raw, False)
    assert np.array_equal(comp.status, full.status) and np.array_equal(comp.score, full.score)
    assert np.array_equal(comp.runs, full.runs) and np.array_equal(comp.run_off, full.run_off)
    assert comp.aln_off[-1] < full.aln_off[-1] // 2                       # far fewer bytes came back
    off, r0, r1 = aligner.expand_rows(buf, offs, raw)
    assert np.array_equal(off, full.aln_off) and np.array_equal(r0, full.aln0) and np.array_equal(r1, full.aln1)


@pytest.mark.parametrize("count", [60, 8000])
def test_packed_input_equals_ascii_input(aligner, batch, count):
    from dbga_b200 import pack_ascii
    from dbga_b200.batch import pack_pairs
    pairs, buf, offs = batch
    buf, offs = pack_pairs(pairs[:count]) if count < len(pairs) else (buf, offs)
    full = aligner.align_buffers(buf, offs, aligner.params(9))
    words = pack_ascii(buf, offs)
    res = aligner.align_buffers(words, offs, aligner.params(9), packed=True)
    _same_rows(full, res)
    assert np.array_equal(full.runs, res.runs)
    with pytest.raises(ValueError):
        pack_ascii(np.frombuffer(b"ACGTNACGT", np.uint8), np.array([0, 5, 9], np.int64))


def test_multi_entry_two_contexts(aligner, batch):
    """dbga_align_batch_multi with two shards (two contexts on the devices present: on a one-GPU box
    both on device 0): shard results side by side equal the single call, ASCII and packed input."""
    import torch
    from dbga_b200 import MultiAligner, pack_ascii
    pairs, buf, offs = batch
    full = aligner.align_buffers(buf, offs, aligner.params(9))
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    m = MultiAligner(devs)
    try:
        for packed in (None, pack_ascii(buf, offs)):
            shards = m.align_buffers(buf if packed is None else None, offs, aligner.params(9), packed=packed)
            assert len(shards) == 2 and shards[0][0] == 0 and 0 < shards[1][0] < len(pairs)
            status = np.concatenate([r.status for _, r in shards])
            assert np.array_equal(status, full.status)
            assert np.array_equal(np.concatenate([r.score for _, r in shards]), full.score)
            assert np.array_equal(np.concatenate([r.aln0 for _, r in shards]), full.aln0)
            assert np.array_equal(np.concatenate([r.aln1 for _, r in shards]), full.aln1)
            assert np.array_equal(np.concatenate([r.runs for _, r in shards]), full.runs)
            lo1, r1 = shards[1]
            assert r1.alignment(5) == full.alignment(lo1 + 5)
    finally:
        m.close()


def test_wire_forms_edge_cases(aligner):
    """Empty and one-pair batches, sequences shorter than a packed word, pairs without anchors, cycles,
    k larger than a sequence and non-ACGT input in every result form; packed input that does not start
    at offset 0; more shards than pairs."""
    from dbga_b200 import MultiAligner, _lib, pack_ascii
    from dbga_b200.batch import pack_pairs
    rnd = O.synth_pair(300, 0.03, 0.01, 77)
    pairs = [("ACGTTGCAAC", "ACGTAGCAAC"), ("ACG", "ACG"), ("AAAAAAAAAAAA", "AAAAAAAAAAAAAA"), rnd,
             ("ACGTACGTAGCTAGCTAGCATCGATCGATCAGCTACGAT", "TTTTTGGGGGCCCCCAAAAATTTTTGGGGGCCCCCAAAAA"),
             ("AC", "ACGTACGT"), ("ACGTNACGTACGT", "ACGTAACGTACGT"), O.synth_pair(900, 0.02, 0.004, 3)]
    for k in (3, 5):
        buf, offs = pack_pairs(pairs)
        full = aligner.align_buffers(buf, offs, aligner.params(k))
        assert set(full.status.tolist()) >= {O.OK, O.K_INVALID, O.BAD_CHAR}
        lean = aligner.align_buffers(buf, offs, aligner.params(k, flags=_lib.FLAG_NO_RUNS | _lib.FLAG_SOP))
        _same_rows(full, lean)
        raw = aligner.align_buffers(buf, offs, aligner.params(k, flags=_lib.FLAG_COMPACT_ROWS), unpack=False)
        off, r0, r1 = aligner.expand_rows(buf, offs, raw)
        assert np.array_equal(off, full.aln_off) and np.array_equal(r0, full.aln0) and np.array_equal(r1, full.aln1)
        # one pair at a time, every form
        for i, p in enumerate(pairs):
            b1, o1 = pack_pairs([p])
            one = aligner.align_buffers(b1, o1, aligner.params(k))
            assert one.status[0] == full.status[i] and one.alignment(0) == full.alignment(i)
            raw1 = aligner.align_buffers(b1, o1, aligner.params(k, flags=_lib.FLAG_COMPACT_ROWS), unpack=False)
            _, x0, x1 = aligner.expand_rows(b1, o1, raw1)
            assert (x0.tobytes().decode(), x1.tobytes().decode()) == full.alignment(i)
    # packed input: the pairs that can be packed, as a window of a larger batch (offsets[0] != 0)
    good = [p for p in pairs if "N" not in p[0]]
    buf, offs = pack_pairs(good)
    words = pack_ascii(buf, offs)
    ref = aligner.align_buffers(buf, offs, aligner.params(3))
    win = aligner.align_buffers(buf, offs[4:], aligner.params(3))                   # pairs 2.. as their own batch (ASCII)
    assert np.array_equal(win.status, ref.status[2:]) and win.alignment(1) == ref.alignment(3)
    m = MultiAligner([0] * 12)                                                     # more shards than pairs
    try:
        for packed in (None, words):
            shards = m.align_buffers(buf if packed is None else None, offs, aligner.params(3), packed=packed)
            assert np.array_equal(np.concatenate([r.status for _, r in shards]), ref.status)
            assert np.array_equal(np.concatenate([r.aln0 for _, r in shards]), ref.aln0)
            assert np.array_equal(np.concatenate([r.aln1 for _, r in shards]), ref.aln1)
    finally:
        m.close()
    empty = aligner.align_buffers(np.zeros(0, np.uint8), np.zeros(1, np.int64), aligner.params(3, flags=_lib.FLAG_COMPACT_ROWS | _lib.FLAG_SOP))
    assert len(empty) == 0


SHARDED_SCRIPT = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from dbga_b200 import Aligner
from dbga_b200.sharding import align_buffers_sharded
from dbga_b200.synth import synth_batch
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
al = Aligner(0)
buf, offs = synth_batch(301, 600, 0.02, 0.004, seed=5)
prm = al.params(9)
got = align_buffers_sharded(al, buf, offs, prm)
if dist.get_rank() == 0:
    want = al.align_buffers(buf, offs, prm)
    for f in ("status", "score", "dp_cells", "dp_segments", "n_anchors", "run_off", "runs", "aln_off", "aln0", "aln1"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    print("SHARDED_OK", int((want.status == 0).sum()))
else:
    assert got is None
al.close()
dist.destroy_process_group()
"""


def test_two_ranks_shard_one_batch_and_gather_arrays(tmp_path):
    """One process per rank (both on cuda:0 here), pairs sharded by index, the shards' results gathered on rank 0 as
    arrays (dbga_b200.sharding.align_buffers_sharded): identical to the single-process result."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "sharded.py"
    script.write_text(SHARDED_SCRIPT.format(root=root, port=29617))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARDED_OK" in outs[0][0]

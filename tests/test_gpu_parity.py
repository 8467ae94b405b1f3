"""Parity of the CUDA path (through the C ABI) against the oracle.  Needs a B200."""
import itertools
import random

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import dbga_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aligner():
    from dbga_b200 import Aligner
    a = Aligner(0)
    yield a
    a.close()


def rand_seq(rng, n, alpha="ACGT"):
    return "".join(rng.choice(alpha) for _ in range(n))


def mutate(rng, s, ps, pi):
    out, i = [], 0
    while i < len(s):
        r = rng.random()
        if r < pi:
            if rng.random() < 0.5:
                out.append(rand_seq(rng, rng.randint(1, 4)))
                out.append(s[i]); i += 1
            else:
                i += rng.randint(1, 4)
        elif r < pi + ps:
            out.append(rng.choice([c for c in "ACGT" if c != s[i]])); i += 1
        else:
            out.append(s[i]); i += 1
    return "".join(out)


def check_batch(aligner, pairs, k, conv=O.DEFAULT, oracle=CO.align_pair, **score):
    res = aligner.align_pairs(pairs, k, gap_len_mode=conv.gap_len_mode, xy_allowed=conv.xy_allowed,
                              pred_order=conv.pred_order, end_order=conv.end_order, **score)
    n_ok = 0
    for i, (s0, s1) in enumerate(pairs):
        ref = oracle(s0, s1, k, conv=conv, **score)
        assert res.status[i] == ref.status, (i, s0, s1, k, res.status[i], ref.status)
        if ref.status in (O.OK, O.CYCLE):
            anchors = [tuple(x) for x in np.asarray(ref.anchors).reshape(-1, 2).tolist()]
            assert [tuple(x) for x in res.anchors(i).tolist()] == anchors, (i, s0, s1, k)
        if ref.status == O.OK:
            a0, a1 = res.alignment(i)
            assert (a0, a1) == (ref.aln0, ref.aln1), (i, s0, s1, k, conv, a0, a1, ref.aln0, ref.aln1)
            assert res.score[i] == ref.score, (i, s0, s1, k)
            assert res.dp_cells[i] == ref.dp_cells and res.dp_segments[i] == ref.dp_segments
            n_ok += 1
    return n_ok


def test_reference_goldens(aligner, ref_fixtures):
    # reference tests/test_debruijn_pairwise.py:50-214 through the CUDA path
    for fx in ref_fixtures["e2e"]:
        res = aligner.align_pairs([tuple(fx["seqs"])], fx["k"], want_graph=True)
        assert res.status[0] == O.OK
        assert list(res.alignment(0)) == fx["aln"], fx["fixture"]
        nd0, nd1 = res.graph_flags(0)
        ids = [1 + int(nd0[:a].sum()) for a, _ in res.anchors(0).tolist()]
        assert ids == fx["merge_node_idx"], fx["fixture"]
    for fx in ref_fixtures["cycles"]:
        res = aligner.align_pairs([tuple(fx["seqs"])], fx["k"])
        assert res.status[0] == O.CYCLE


def test_recorded_reference_fuzz(aligner, ref_fuzz):
    # 400 inputs recorded from the reference's own code (tests/golden/ref_graph_fuzz.json)
    by_k = {}
    for c in ref_fuzz:
        by_k.setdefault(c["k"], []).append(c)
    for k, cases in by_k.items():
        res = aligner.align_pairs([tuple(c["seqs"]) for c in cases], k, want_graph=True)
        for i, c in enumerate(cases):
            ref = c["ref"]
            if "error" in ref:
                assert res.status[i] == O.CYCLE
                continue
            assert res.status[i] == O.OK
            assert list(res.alignment(i)) == ref["aln"]
            nd0, nd1 = res.graph_flags(i)
            anchors = res.anchors(i).tolist()
            assert [1 + int(nd0[:a].sum()) for a, _ in anchors] == ref["merge_node_idx"]
            # id_count = 4 terminal nodes + non-duplicate k-mers of both - merge nodes
            assert 4 + int(nd0.sum()) + int(nd1.sum()) - len(anchors) == ref["id_count"]
            assert res.dp_segments[i] == len(ref["dp_calls"])


@pytest.mark.parametrize("k", [1, 2, 3, 5, 9, 15, 16, 21, 31])
def test_small_pairs_vs_oracle(aligner, k):
    rng = random.Random(1000 + k)
    pairs = []
    for _ in range(300):
        n = rng.choice([k, k + 1, 40, 90, 200, 600])
        n = max(n, k)
        s0 = rand_seq(rng, n, rng.choice(["ACGT", "ACGT", "AC"]))
        if rng.random() < 0.8:
            s1 = mutate(rng, s0, rng.choice([0, 0.02, 0.1]), rng.choice([0, 0.01, 0.05]))
        else:
            s1 = rand_seq(rng, rng.randint(k, n + 5))
        if len(s1) < k:
            s1 = s0
        pairs.append((s0, s1))
    assert check_batch(aligner, pairs, k) > 50


def test_all_conventions(aligner):
    rng = random.Random(5)
    pairs = []
    for _ in range(60):
        s0 = rand_seq(rng, rng.choice([30, 80, 150]))
        pairs.append((s0, mutate(rng, s0, 0.08, 0.05) or "A"))
    pairs = [(a, b) for a, b in pairs if len(b) >= 5]
    for g, xy, po, eo in itertools.product((0, 1), (False, True), O.ORDERS, ("XMY", "MYX", "YMX")):
        check_batch(aligner, pairs, 5, conv=O.Convention(g, xy, po, eo))
    check_batch(aligner, pairs, 5, match=5, transition=-4, transversion=-4, d=7, e=1)
    check_batch(aligner, pairs, 5, match=1, transition=0, transversion=-1, d=0, e=0)


def test_status_codes(aligner):
    res = aligner.align_pairs([("ACGT", "ACGTA"), ("ACGNACGT", "ACGTACGT"), ("ACGTACGTAC", "ACGTTCGTAC"),
                               ("acgtacgt", "ACGTACGT")], 5)
    assert res.status.tolist() == [O.K_INVALID, O.BAD_CHAR, O.OK, O.BAD_CHAR]


def test_global_align_classes(aligner):
    """dna_global_aln through every DP kernel class: tiny, warp (1 and >1 row strips),
    cluster wavefront (part of a cluster, whole cluster, 2 and 3 row strips), plus empty sides
    (utils.py:187-192)."""
    rng = random.Random(11)
    shapes = [(1, 1), (5, 7), (32, 32), (33, 20), (20, 33), (100, 90), (256, 300), (257, 301), (700, 650),
              (40, 3000), (3000, 40), (600, 600), (2100, 2300), (513, 520), (0, 5), (6, 0),
              (9000, 300), (300, 9000), (16500, 70), (8192, 64), (8193, 64)]
    pairs = []
    for m, n in shapes:
        x = rand_seq(rng, m)
        if m and n and rng.random() < 0.7:
            y = (mutate(rng, x, 0.1, 0.05) + rand_seq(rng, n))[:n]
        else:
            y = rand_seq(rng, n)
        pairs.append((x, y))
    for conv in (O.DEFAULT, O.Convention(1, True, "MYX", "YMX")):
        res = aligner.global_align_pairs(pairs, gap_len_mode=conv.gap_len_mode, xy_allowed=conv.xy_allowed,
                                         pred_order=conv.pred_order, end_order=conv.end_order)
        s = O.make_dna_scoring_dict(10, -1, -8)
        for i, (x, y) in enumerate(pairs):
            assert res.status[i] == O.OK
            if x and y:
                sc, ax, ay = CO.gotoh(x, y, conv=conv)
                assert res.score[i] == sc, (i, len(x), len(y))
            else:
                ax, ay = O.dna_global_aln(x, y, s, 10, 2)
            assert res.alignment(i) == (ax, ay), (i, len(x), len(y), conv)


def test_many_long_segments_take_the_cta_kernel(aligner):
    """More long segments than clusters in flight: the list goes to the CTA-per-segment
    kernel instead of the cluster kernel (same results either way)."""
    rng = random.Random(23)
    pairs = []
    for _ in range(330):
        m = rng.randrange(520, 760)
        x = rand_seq(rng, m)
        pairs.append((x, mutate(rng, x, 0.1, 0.05)))
    res = aligner.global_align_pairs(pairs)
    for i, (x, y) in enumerate(pairs):
        sc, ax, ay = CO.gotoh(x, y)
        assert res.status[i] == O.OK and res.score[i] == sc
        assert res.alignment(i) == (ax, ay), i


def test_large_path_vs_oracle(aligner):
    """Pairs too big for the shared-memory table take the global-memory stage-1 path."""
    pairs = [O.synth_pair(30000, 0.004, 0.001, 3), O.synth_pair(12000, 0.01, 0.002, 4),
             O.synth_pair(60000, 0.002, 0.0005, 5)]
    n_ok = check_batch(aligner, pairs, 11) + check_batch(aligner, pairs, 17)
    assert n_ok >= 1
    # mixed small + large in one batch, and graph flags on the large path
    small = [O.synth_pair(300, 0.02, 0.004, s) for s in range(20)]
    check_batch(aligner, small + pairs[:1] + small, 11)
    res = aligner.align_pairs(pairs[:1], 11, want_graph=True, stages=2)
    ref = CO.align_pair(*pairs[0], 11, want_graph=True, stages=2)
    nd0, nd1 = res.graph_flags(0)
    assert np.array_equal(nd0, ref.nondup0) and np.array_equal(nd1, ref.nondup1)


def test_cfg3_sample_vs_oracle(aligner):
    """BASELINE config 3 shape (1.5 kb, 2 % divergence, k=9): 2000 pairs bit-exact."""
    pairs = [O.synth_pair(1500, 0.018, 0.002, 100 + s) for s in range(2000)]
    n_ok = check_batch(aligner, pairs, 9)
    assert n_ok > 1000


def test_full_size_properties(aligner):
    """Size-independent properties at a larger batch: degapped rows == inputs, equal row
    lengths, no all-gap column, idempotent across two calls."""
    pairs = [O.synth_pair(1500, 0.018, 0.002, 50000 + s) for s in range(20000)]
    r1 = aligner.align_pairs(pairs, 9)
    r2 = aligner.align_pairs(pairs, 9)
    assert np.array_equal(r1.status, r2.status) and np.array_equal(r1.aln0, r2.aln0) and np.array_equal(r1.aln1, r2.aln1)
    assert (r1.status == O.OK).sum() > 10000
    for i in range(0, len(pairs), 37):
        if r1.status[i] != O.OK:
            continue
        a0, a1 = r1.alignment(i)
        assert len(a0) == len(a1)
        assert a0.replace("-", "") == pairs[i][0] and a1.replace("-", "") == pairs[i][1]
        assert not any(x == "-" and y == "-" for x, y in zip(a0, a1))


def test_pipelined_batch_equals_direct(aligner):
    """Batches >= 16 MB go through the chunked 3-lane copy/compute pipeline of
    dbga_align_batch; the result must be identical to aligning the halves directly."""
    pairs = [O.synth_pair(1500, 0.018, 0.002, 7000 + s) for s in range(7000)]     # ~21 MB
    big = aligner.align_pairs(pairs, 9)
    parts = [aligner.align_pairs(pairs[i:i + 3500], 9) for i in (0, 3500)]         # ~10 MB each: direct path
    assert big.timing["launches"] > parts[0].timing["launches"]                     # several chunks ran
    status = np.concatenate([p.status for p in parts])
    assert np.array_equal(big.status, status)
    assert np.array_equal(big.score, np.concatenate([p.score for p in parts]))
    assert np.array_equal(big.n_anchors, np.concatenate([p.n_anchors for p in parts]))
    assert np.array_equal(big.aln0, np.concatenate([p.aln0 for p in parts]))
    assert np.array_equal(big.aln1, np.concatenate([p.aln1 for p in parts]))
    assert np.array_equal(big.runs, np.concatenate([p.runs for p in parts]))
    for i in (0, 1, 3499, 3500, 6999):
        j, part = (i, parts[0]) if i < 3500 else (i - 3500, parts[1])
        assert big.alignment(i) == part.alignment(j)
        assert np.array_equal(big.anchors(i), part.anchors(j))


def test_big_segments_in_big_batch(aligner):
    """With >= 592 pairs in the batch, segments up to 2^24 cells go to the warp-per-segment
    wavefront (many row strips) instead of the CTA kernel."""
    rng = random.Random(21)
    pairs = [(rand_seq(rng, 12), rand_seq(rng, 12)) for _ in range(600)]
    x = rand_seq(rng, 2100)
    pairs += [(x, (mutate(rng, x, 0.1, 0.05) + rand_seq(rng, 2300))[:2300]), (rand_seq(rng, 700), rand_seq(rng, 650)),
              (rand_seq(rng, 3000), rand_seq(rng, 40)), (rand_seq(rng, 40), rand_seq(rng, 3000))]
    res = aligner.global_align_pairs(pairs)
    for i in list(range(0, 600, 97)) + [600, 601, 602, 603]:
        sc, ax, ay = CO.gotoh(*pairs[i])
        assert res.status[i] == O.OK and res.score[i] == sc, i
        assert res.alignment(i) == (ax, ay), i


def test_edge_cases(aligner):
    """Empty batch, single-base sequences, k == len, very unbalanced lengths, homopolymers
    (every k-mer duplicate -> whole-sequence DP), identical pairs (only the tail reaches the DP)."""
    empty = aligner.align_pairs([], 3)
    assert len(empty) == 0 and empty.aln_off.tolist() == [0]
    rng = random.Random(31)
    s = rand_seq(rng, 2000)
    pairs = [("A", "A"), ("A", "C"), ("ACGTACG", "ACGTACG"), (s, s[990:1002]), (s[5:25], s), ("A" * 300, "A" * 280),
             ("AC" * 200, "AC" * 190 + "G"), (s, s), (s[:1200], s[:600] + "T" + s[600:1200]), ("ACGT" * 50, "TGCA" * 50)]
    for k in (1, 3, 7):
        use = [(a, b) for a, b in pairs if len(a) >= k and len(b) >= k]
        check_batch(aligner, use, k)
    check_batch(aligner, [("ACGTACG", "ACGTACG"), ("ACGTACG", "ACGTTCG")], 7)          # k == len


def test_extreme_scores_and_limits(aligner):
    rng = random.Random(32)
    pairs = []
    for _ in range(40):
        a = rand_seq(rng, rng.choice([30, 120, 400]))
        pairs.append((a, mutate(rng, a, 0.1, 0.05) or "ACGT"))
    pairs = [(a, b) for a, b in pairs if len(b) >= 4]
    check_batch(aligner, pairs, 4, match=500, transition=-500, transversion=-499, d=1000, e=999)
    check_batch(aligner, pairs, 4, match=0, transition=0, transversion=0, d=0, e=0)
    check_batch(aligner, pairs, 4, match=-3, transition=2, transversion=5, d=3, e=7)   # unusual but legal
    with pytest.raises(RuntimeError, match="substitution score"):
        aligner.align_pairs(pairs, 4, match=501)
    with pytest.raises(RuntimeError, match="k must be"):
        aligner.align_pairs(pairs, 0)
    res = aligner.align_pairs(pairs, 32)                 # wide k-mers are a supported path (test_wide_kmers_vs_oracle)
    assert set(res.status.tolist()) <= {O.OK, O.CYCLE, O.K_INVALID}


def test_scratch_limit_is_an_explicit_status():
    """A segment whose traceback does not fit the scratch limit is reported as TOO_LARGE,
    never silently dropped; the others in the batch are unaffected."""
    from dbga_b200 import Aligner
    al = Aligner(0)
    al.set_scratch_limit(1 << 20)
    rng = random.Random(33)
    big = (rand_seq(rng, 3000), rand_seq(rng, 3000))          # no shared 15-mer: one 9 M-cell segment
    small = O.synth_pair(400, 0.02, 0.004, 1)
    res = al.align_pairs([small, big, small], 15)
    ref = CO.align_pair(*small, 15)
    assert res.status.tolist() == [ref.status, O.TOO_LARGE, ref.status]
    if ref.status == O.OK:
        assert res.alignment(0) == (ref.aln0, ref.aln1) == res.alignment(2)
    assert res.alignment(1) == ("", "")
    al.close()


def test_graph_flags_in_a_batch(aligner):
    pairs = [O.synth_pair(200, 0.05, 0.01, s) for s in range(30)]
    res = aligner.align_pairs(pairs, 5, want_graph=True, stages=2)
    for i, (a, b) in enumerate(pairs):
        ref = CO.align_pair(a, b, 5, want_graph=True, stages=2)
        assert res.status[i] == ref.status
        nd0, nd1 = res.graph_flags(i)
        assert np.array_equal(nd0, ref.nondup0) and np.array_equal(nd1, ref.nondup1)
        assert [tuple(x) for x in res.anchors(i).tolist()] == [tuple(x) for x in ref.anchors.tolist()]


def test_packed16_wavefront(aligner):
    """The packed 16-bit wavefront (two cells per register, dp_wf16_kernel) against the oracle:
    every warp class (1 / 2 / 4 / 8 rows per half-lane), partial lanes, segments on both sides
    of the encoding's ceiling (per_step * min(m, n) vs ulimit: above it the 32-bit kernel runs),
    every tie-breaking convention, and scoring tables at and beyond the int8 LUT limits."""
    rng = random.Random(41)
    shapes = [(33, 33), (64, 64), (65, 40), (64, 500), (63, 64), (100, 90), (128, 128), (127, 129), (129, 300),
              (200, 35), (256, 256), (255, 257), (257, 300), (300, 2), (2, 300), (500, 512), (512, 512), (511, 37),
              (512, 2000), (40, 3000), (300, 9000), (513, 513), (1000, 1000), (1080, 1090), (1090, 1100),
              (1024, 60), (60, 1024), (33, 1), (1, 33), (450, 450)]
    pairs = []
    for m, n in shapes:
        x = rand_seq(rng, m)
        r = rng.random()
        if r < 0.6:
            y = (mutate(rng, x, 0.1, 0.05) + rand_seq(rng, n))[:n]
        elif r < 0.8:
            y = rand_seq(rng, n)
        else:                                   # shifted copy: long terminal gaps
            y = (rand_seq(rng, n // 3) + x)[:n]
        pairs.append((x, y))

    def run(conv=O.DEFAULT, sub=None, **score):
        res = aligner.global_align_pairs(sub or pairs, gap_len_mode=conv.gap_len_mode, xy_allowed=conv.xy_allowed,
                                         pred_order=conv.pred_order, end_order=conv.end_order, **score)
        for i, (x, y) in enumerate(sub or pairs):
            sc, ax, ay = CO.gotoh(x, y, conv=conv, **score)
            assert res.status[i] == O.OK
            assert res.score[i] == sc, (i, len(x), len(y), conv, score)
            assert res.alignment(i) == (ax, ay), (i, len(x), len(y), conv, score)

    run()
    # with >= 592 pairs in the batch the big shapes stay in the warp classes instead of the CTA class
    filler = [(rand_seq(rng, 40), rand_seq(rng, 45)) for _ in range(600)]
    run(sub=pairs + filler)
    small = [p for p in pairs if len(p[0]) * len(p[1]) <= 70000]
    for g, xy, po, eo in itertools.product((0, 1), (False, True), O.ORDERS, ("XMY", "MYX", "YMX")):
        run(O.Convention(g, xy, po, eo), sub=small)
    run(O.Convention(1, True, "YXM", "MYX"))
    run(match=15, transition=-4, transversion=-48, d=40, e=8)      # s + 2e in [-32, 31], go = 40: the limits
    run(match=16, transition=-4, transversion=-8, d=40, e=8)       # one past the LUT limit: 32-bit kernels
    run(match=5, transition=-4, transversion=-4, d=41, e=1)        # gap open past the limit: 32-bit kernels
    run(match=1, transition=0, transversion=-1, d=0, e=0)
    run(match=31, transition=-30, transversion=-32, d=3, e=0)      # steep: ceiling reached at ~490 rows
    run(match=-3, transition=2, transversion=5, d=7, e=3)          # mismatches score higher than matches
    # gap extension so large that 32 R rows of a lane pair pass the ceiling (growth 80 per diagonal step):
    # the two-segments-per-warp kernels hand their classes to the half-lane kernels (up to ~189 rows) and
    # those hand the rest to the 32-bit kernels
    run(match=-50, transition=-60, transversion=-70, d=40, e=40)


def test_packed16_row_strips(aligner):
    """Long segments as packed row strips (dp_wf16s_kernel: 512-row strips pipelined through boundary rows,
    frames re-anchored every 256 columns): one to many strips, partial last strips, segments whose scores
    grow far past the int16 range (similar sequences), fall far below zero (random ones) or do both
    (shifted copies: a long terminal gap, then matches), few and many segments per batch (strip-major ticket
    order), every predecessor / terminal tie order, and scoring where the strip form is off (32-bit kernels)."""
    rng = random.Random(4242)
    shapes = [(513, 513), (1024, 1024), (1025, 1030), (1100, 1100), (1536, 700), (700, 1536), (2048, 2048), (2049, 255),
              (3000, 5000), (5000, 3000), (4097, 300), (600, 9000), (9000, 600), (1500, 1500), (2600, 2500), (6500, 6400)]
    pairs = []
    for q, (m, n) in enumerate(shapes):
        x = rand_seq(rng, m)
        kind = q % 4
        if kind == 0:
            y = (mutate(rng, x, 0.05, 0.01) + rand_seq(rng, n))[:n]
        elif kind == 1:
            y = (mutate(rng, x, 0.3, 0.04) + rand_seq(rng, n))[:n]       # cfg5's bubbles
        elif kind == 2:
            y = rand_seq(rng, n)
        else:
            y = (rand_seq(rng, n // 3) + x)[:n]
        pairs.append((x, y))

    def run(sub, conv=O.DEFAULT, **score):
        res = aligner.global_align_pairs(sub, gap_len_mode=conv.gap_len_mode, xy_allowed=conv.xy_allowed,
                                         pred_order=conv.pred_order, end_order=conv.end_order, **score)
        for i, (x, y) in enumerate(sub):
            if len(x) < 100:
                continue
            sc, ax, ay = CO.gotoh(x, y, conv=conv, **score)
            assert res.status[i] == O.OK
            assert res.score[i] == sc, (i, len(x), len(y), conv, score)
            assert res.alignment(i) == (ax, ay), (i, len(x), len(y), conv, score)

    run(pairs)                                                           # few pairs: everything above 2^18 cells is long
    filler = [(rand_seq(rng, 40), rand_seq(rng, 45)) for _ in range(600)]
    run(pairs[:12] + filler)                                             # many pairs: only what the one-pass kernel cannot hold
    for po, eo in (("MXY", "YMX"), ("YXM", "XMY"), ("MYX", "MYX")):
        run(pairs[:9], O.Convention(1, False, po, eo))
    run(pairs[:9], match=17, transition=-3, transversion=-20, d=30, e=3)   # per_step 23: the steepest the strip form takes
    run(pairs[:6], match=22, transition=-4, transversion=-8, d=40, e=4)    # per_step 30: 32-bit kernels
    run(pairs[:6], match=1, transition=0, transversion=-1, d=0, e=0)
    # many long segments of one size: tickets of several strips per segment interleave over the whole grid
    many = []
    for _ in range(700):
        x = rand_seq(rng, rng.randint(1030, 1700))
        many.append((x, mutate(rng, x, 0.05, 0.01)))
    res = aligner.global_align_pairs(many)
    assert (res.status == O.OK).all()
    for i in range(0, len(many), 7):
        sc, ax, ay = CO.gotoh(*many[i])
        assert (res.score[i],) + res.alignment(i) == (sc, ax, ay), (i, len(many[i][0]), len(many[i][1]))


def test_packed16_full_size_properties(aligner):
    """The DP microbenchmark's shapes at size (thousands of segments per packed kernel, lists long
    enough that every warp pairs many different neighbours): degapped rows == inputs, the affine
    score recomputed from the returned rows == the reported score (an independent restatement of
    the objective, not of the kernel), and every 50th pair bit-exact against the oracle."""
    def affine_score(a, b, match=10, transition=-1, transversion=-8, d=10, e=2):
        pur = {"A", "G"}
        sc, gx, gy = 0, False, False          # inside a gap in row a / row b
        for x, y in zip(a, b):
            if x == "-":
                sc -= e if gx else d
                gx, gy = True, False
            elif y == "-":
                sc -= e if gy else d
                gx, gy = False, True
            else:
                sc += match if x == y else (transition if (x in pur) == (y in pur) else transversion)
                gx = gy = False
        return sc

    rng = random.Random(77)
    pairs = []
    for lo, hi, cnt in ((33, 64, 1500), (65, 128, 900), (129, 256, 600), (257, 512, 300), (513, 1024, 120)):
        for _ in range(cnt):
            m = rng.randint(lo, hi)
            x = rand_seq(rng, m)
            y = mutate(rng, x, 0.05, 0.01) if rng.random() < 0.8 else rand_seq(rng, rng.randint(lo, hi))
            pairs.append((x, y or "A"))
    res = aligner.global_align_pairs(pairs)
    assert (res.status == O.OK).all()
    for i, (x, y) in enumerate(pairs):
        ax, ay = res.alignment(i)
        assert len(ax) == len(ay) and ax.replace("-", "") == x and ay.replace("-", "") == y, i
        assert affine_score(ax, ay) == res.score[i], (i, len(x), len(y))
        if i % 50 == 0:
            sc, ox, oy = CO.gotoh(x, y)
            assert (res.score[i], ax, ay) == (sc, ox, oy), (i, len(x), len(y))


def test_kmer_stats_and_k_heuristics(aligner):
    """dbga_kmer_stats (|K0|, |K1|, |K0 & K1| on the device) against the oracle's set arithmetic
    for k in 1..31 (32- and 64-bit keys), both table paths (shared memory; one long pair in the
    global-memory table), error statuses; and predict_p / predict_final_p /
    kmers_larger_than_threshold built on it against values recorded from the reference's own
    functions (tests/golden/ref_kmer_heuristics.json)."""
    import json
    import os
    from dbga_b200 import kmersize as KS
    rng = random.Random(91)
    pairs = [O.synth_pair(n, ps, pi, 7000 + i) for i, (n, ps, pi) in enumerate(
        [(40, 0.05, 0.02), (300, 0.1, 0.02), (1500, 0.018, 0.002), (1500, 0.3, 0.05), (5000, 0.01, 0.001)] * 3)]
    pairs += [("A" * 200, "A" * 150), ("ACGT" * 60, "TGCA" * 50), (rand_seq(rng, 64), rand_seq(rng, 900))]
    for k in (1, 2, 3, 5, 9, 12, 15, 16, 17, 24, 31):
        use = [(a, b) for a, b in pairs if len(a) >= k and len(b) >= k]
        st, d0, d1, sh = aligner.kmer_stats(use, k)
        assert (st == O.OK).all()
        for i, (a, b) in enumerate(use):
            assert (int(d0[i]), int(d1[i]), int(sh[i])) == O.kmer_stats(a, b, k), (i, k, len(a), len(b))
    # a pair far beyond the shared-memory table next to small ones, and the error statuses
    big = O.synth_pair(60000, 0.01, 0.002, 5)
    mixed = [pairs[0], big, ("ACG", "ACGTACGT" * 4), ("ACGTNCGT" * 4, "ACGTACGT" * 4), pairs[2]]
    for k in (5, 11, 20):
        st, d0, d1, sh = aligner.kmer_stats(mixed, k)
        assert st.tolist() == [O.OK, O.OK, O.K_INVALID, O.BAD_CHAR, O.OK]
        for i in (0, 1, 4):
            assert (int(d0[i]), int(d1[i]), int(sh[i])) == O.kmer_stats(*mixed[i], k), (i, k)
    st, *_ = aligner.kmer_stats([(big[0], big[1][:30000] + "N" + big[1][30000:])], 11)
    assert st.tolist() == [O.BAD_CHAR]
    # the reference's heuristics, bit for bit (same float expression on the same integers)
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_kmer_heuristics.json")))
    for c in g["cases"]:
        for k, p in c["predict_p"].items():
            assert KS.predict_p_batch([(c["s0"], c["s1"])], int(k))[0] == p
        for key, v in c["dup_prop_gt"].items():
            k, th = key.split(":")
            assert KS.kmers_larger_than_threshold(c["s0"], int(k), float(th)) == v
    finals = KS.predict_final_p_batch([(c["s0"], c["s1"]) for c in g["cases"]])
    assert finals.tolist() == [c["predict_final_p"] for c in g["cases"]]
    with pytest.raises(ValueError, match="Invalid k size"):
        KS.predict_p_batch([("ACGT", "ACG")], 4)


def test_very_long_pairs_take_the_cluster_stage2(aligner):
    """Pairs with more than 65 536 seq-2 positions run stage 2 on a thread-block cluster
    (stage2_cluster_kernel): anchors, status and alignment against the oracle, next to shorter
    pairs in the same batch (CTA stage 2), with and without a cycle."""
    pairs = [O.synth_pair(90000, 0.004, 0.001, 411), O.synth_pair(1500, 0.018, 0.002, 7), O.synth_pair(30000, 0.004, 0.001, 12),
             O.synth_pair(70000, 0.01, 0.002, 413)]
    for k in (15, 11):
        check_batch(aligner, pairs, k)


@pytest.mark.parametrize("k", [32, 33, 40, 64, 101])
def test_wide_kmers_vs_oracle(aligner, k):
    """k >= 32 (the reference accepts any k <= len, utils.py:129-131; adpt_dbg_alignment starts at
    k = 33 beyond 2^21 bases): by-reference keys in the global-memory table, against the oracle."""
    rng = random.Random(4000 + k)
    pairs = []
    for t in range(40):
        n = rng.choice([k, k + 1, 150, 600, 1500, 4000])
        n = max(n, k)
        s0 = rand_seq(rng, n, rng.choice(["ACGT", "ACGT", "AC"]))
        if rng.random() < 0.3 and n > 300:
            s0 = s0[:200] + s0[100:300] + s0[200:]                  # repeated k-mers
        s1 = mutate(rng, s0, rng.choice([0, 0.005, 0.02]), rng.choice([0, 0.002]))
        if len(s1) < k:
            s1 = s0
        pairs.append((s0, s1))
    pairs.append(("A" * (k + 30), "A" * (k + 10)))                # every k-mer duplicate
    pairs.append((pairs[0][0], pairs[0][0][: k - 1] if k > 1 else "A"))   # K_INVALID
    assert check_batch(aligner, pairs, k) >= 10
    res = aligner.align_pairs(pairs[:6], k, want_graph=True, stages=2)
    for i in range(6):
        ref = CO.align_pair(*pairs[i], k, want_graph=True, stages=2)
        if ref.status in (O.OK, O.CYCLE):
            nd0, nd1 = res.graph_flags(i)
            assert np.array_equal(nd0, ref.nondup0) and np.array_equal(nd1, ref.nondup1)
    st, d0, d1, sh = aligner.kmer_stats(pairs[:20], k)
    for i in range(20):
        if st[i] == O.OK:
            assert (d0[i], d1[i], sh[i]) == O.kmer_stats(*pairs[i], k), i


def test_long_homopolymer_does_not_wrap_the_medium_counter(aligner):
    """A k-mer repeated in seq1 and occurring more than 32 768 times in seq2 (a 33 kb homopolymer stretch).
    Round 1's medium-pair kernel counted seq-2 occurrences in 15 bits under its "repeated in seq1" flag (a
    count that large would have carried into it); the by-reference slots keep occurrences as flags
    ("seen", "seen again") that only get OR-ed in, so such pairs (41 kb: two table partitions, in graph
    mode four) stay on the shared-memory path.  Stages 1-2 against the oracle."""
    rng = random.Random(91)
    r1, r2 = rand_seq(rng, 4000), rand_seq(rng, 4000)
    s0 = r1 + "A" * 33000 + r2
    s1 = mutate(rng, r1, 0.01, 0.002) + "A" * 33200 + mutate(rng, r2, 0.01, 0.002)
    for k in (11, 15):
        res = aligner.align_pairs([(s0, s1), (s1, s0)], k, stages=2, want_graph=True)
        for i, (a, b) in enumerate(((s0, s1), (s1, s0))):
            ref = CO.align_pair(a, b, k, stages=2, want_graph=True)
            assert res.status[i] == ref.status
            assert np.array_equal(res.anchors(i), ref.anchors)
            nd0, nd1 = res.graph_flags(i)
            assert np.array_equal(nd0, ref.nondup0) and np.array_equal(nd1, ref.nondup1)

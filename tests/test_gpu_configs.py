"""Every BASELINE.json config at (or near) its stated size, CUDA path vs the oracle: status,
anchors, integer score, DP cells and the full gapped rows of every pair (compared pair by pair in
C by oracle_check_batch).  The cfg3 / cfg4 batches are the very batches bench.py times (same
generator, same seed); cfg5 is the 1 Mb pair with indel-rich multi-kb bubbles.  Needs a B200."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import dbga_oracle as O

pytestmark = pytest.mark.gpu

BENCH_SEED = 20261019          # bench.py --seed default


@pytest.fixture(scope="module")
def aligner():
    from dbga_b200 import Aligner
    a = Aligner(0)
    yield a
    a.close()


def diff_vs_oracle(aligner, buf, offs, k, **kw):
    """-> (gpu result, oracle columns); asserts bit-exact agreement on everything."""
    prm = aligner.params(k, **kw)
    res = aligner.align_buffers(buf, offs, prm)
    ref = CO.check_batch(buf, offs, k, res.aln_off, res.aln0, res.aln1, res.run_off, res.runs.reshape(-1), **kw)
    assert np.array_equal(res.status, ref["status"]), np.nonzero(res.status != ref["status"])[0][:10]
    ok = ref["status"] == O.OK
    rep = ok | (ref["status"] == O.CYCLE)
    assert np.array_equal(res.score[ok], ref["score"][ok])
    assert np.array_equal(res.dp_cells[ok], ref["dp_cells"][ok])
    assert np.array_equal(np.diff(res.aln_off)[ok], ref["aln_len"][ok])
    assert np.array_equal(res.n_anchors[rep], ref["n_anchors"][rep])
    assert not ref["diff"].any(), ("rows/anchors differ", np.nonzero(ref["diff"])[0][:10], ref["diff"][ref["diff"] != 0][:10])
    return res, ref


def test_cfg3_timed_batch_vs_oracle(aligner):
    """The first 20 000 pairs of the batch bench.py times (config [2]: 1.5 kb, 2 %, k=9)."""
    from dbga_b200.synth import synth_batch
    buf, offs = synth_batch(20_000, 1500, 0.018, 0.002, seed=BENCH_SEED)
    res, ref = diff_vs_oracle(aligner, buf, offs, 9)
    assert (ref["status"] == O.OK).sum() > 15_000 and (ref["status"] == O.CYCLE).sum() > 1_000


def test_cfg4_200_pairs_of_30kb_vs_oracle(aligner):
    """Config [3]: 30 kb pairs at 0.5 % divergence, k=11 (the medium-pair path: partitioned
    shared-memory table + generic stage 2), 200 pairs of bench.py --workload cfg4's batch."""
    from dbga_b200.synth import synth_batch
    buf, offs = synth_batch(200, 30_000, 0.004, 0.001, seed=BENCH_SEED)
    res, ref = diff_vs_oracle(aligner, buf, offs, 11)
    assert (ref["status"] == O.OK).sum() > 80 and (ref["status"] == O.CYCLE).sum() > 20


def test_cfg5_one_megabase_pair_with_indel_bubbles_vs_oracle(aligner):
    """Config [4]: 1 Mb pair, k=15, ten planted 2-5 kb bubbles with 30 % substitutions and 4 %
    indel events (both sides multi-kb, different lengths): global-table stage 1, cluster
    stage 2, long-segment wavefront DP.  Seed search as SURVEY 8d: seed 3 is the first the
    oracle accepts, seed 1 is a cycle (the reference's ValueError)."""
    from dbga_b200.synth import synth_long_pair
    buf, offs = synth_long_pair(seed=3)
    res, ref = diff_vs_oracle(aligner, buf, offs, 15)
    assert ref["status"][0] == O.OK and ref["dp_cells"][0] > 50_000_000 and res.dp_segments[0] >= 10
    a0, a1 = res.alignment(0)
    n = int(offs[1])
    assert a0.replace("-", "").encode() == buf[:n].tobytes() and a1.replace("-", "").encode() == buf[n:].tobytes()
    # the planted bubbles reach the DP as segments of more than 1 500 x 1 500 cells (a chance 15-mer may split one)
    runs = res.runs[int(res.run_off[0]):int(res.run_off[1])]
    gaps_a = runs[1:, 0] - (runs[:-1, 0] + runs[:-1, 2] - 1)
    gaps_b = runs[1:, 1] - (runs[:-1, 1] + runs[:-1, 2] - 1)
    assert ((gaps_a > 1500) & (gaps_b > 1500)).sum() >= 8
    assert (gaps_a != gaps_b).sum() >= 5            # indel-rich: the two sides differ in length
    buf, offs = synth_long_pair(seed=1)
    res, ref = diff_vs_oracle(aligner, buf, offs, 15)
    assert ref["status"][0] == O.CYCLE


def test_cfg5_shorter_pairs_other_conventions(aligner):
    """200 kb pairs with indel bubbles under non-default scoring / tie conventions."""
    from dbga_b200.synth import synth_long_pair
    done = 0
    for seed in range(1, 30):
        buf, offs = synth_long_pair(n=200_000, n_bubbles=4, seed=seed)
        if CO.align_batch(buf, offs, 15, threads=1)["status"][0] != O.OK:
            continue
        conv = [O.DEFAULT, O.Convention(1, True, "MXY", "XYM"), O.Convention(0, False, "YMX", "MYX")][done]
        prm_kw = dict(gap_len_mode=conv.gap_len_mode, xy_allowed=conv.xy_allowed, pred_order=conv.pred_order,
                      end_order=conv.end_order)
        prm = aligner.params(15, **prm_kw)
        res = aligner.align_buffers(buf, offs, prm)
        ref = CO.check_batch(buf, offs, 15, res.aln_off, res.aln0, res.aln1, res.run_off, res.runs.reshape(-1), conv=conv)
        assert res.status[0] == ref["status"][0] == O.OK
        assert res.score[0] == ref["score"][0] and res.dp_cells[0] == ref["dp_cells"][0]
        assert not ref["diff"].any(), (seed, conv)
        done += 1
        if done == 3:
            break
    assert done == 3


def test_long_segments_through_the_pipelined_batch(aligner):
    """Long DP segments inside a batch large enough for the chunked pipeline of dbga_align_batch (>= 16 MB:
    four lanes, each with its own pools, streams and side stream for the long class): 900 pairs of 12 kb
    with one planted indel-rich 1.5-2.5 kb bubble each, k = 15.  Every bubble is more than one packed pass
    holds, so it runs as row strips (dp_wf16s_kernel) -- in chunks of a few hundred pairs, i.e. with both
    strip heights in one call.  Everything against the oracle, and the device-resident plan path must give
    byte-identical rows."""
    from dbga_b200.synth import synth_long_pair
    bufs, offs, base = [], [0], 0
    for s in range(900):
        b, o = synth_long_pair(n=12_000, n_bubbles=1, seed=1000 + s, bubble_len=(1500, 2500))
        bufs.append(b)
        offs += [base + int(o[1]), base + int(o[2])]
        base += int(o[2])
    buf, offs = np.concatenate(bufs), np.asarray(offs, np.int64)
    assert buf.nbytes >= 16 << 20
    res, ref = diff_vs_oracle(aligner, buf, offs, 15)
    ok = ref["status"] == O.OK
    assert ok.sum() > 700 and (ref["dp_cells"][ok] > 1_000_000).sum() > 500
    assert res.timing["launches"] > 40                                 # several chunks ran
    import torch
    d = torch.from_numpy(np.concatenate([buf, np.zeros(64, np.uint8)])).cuda()
    aligner.plan(offs, aligner.params(15))
    aligner.run(d.data_ptr())
    direct = aligner.fetch()
    assert np.array_equal(direct.status, res.status) and np.array_equal(direct.score, res.score)
    assert np.array_equal(direct.aln_off, res.aln_off) and np.array_equal(direct.aln0, res.aln0) and np.array_equal(direct.aln1, res.aln1)

"""Host-side cost of dbga_plan (per-pair metadata, stage-1 classing) for the cfg3 batch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbga_b200 import Aligner
from dbga_b200.synth import synth_batch
al = Aligner(0)
buf, offs = synth_batch(100_000, 1500, 0.018, 0.002, seed=1)
prm = al.params(9)
al.plan(offs, prm)
t0 = time.perf_counter()
for _ in range(10):
    al.plan(offs, prm)
print("dbga_plan, 100000 pairs: %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))

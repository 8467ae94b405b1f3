"""Per-stage CUDA-event times of the cfg3 batch (stages 1+2, stage 3, assembly), device-resident: a quick A/B check."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dbga_b200 import Aligner
from dbga_b200.synth import synth_batch
al = Aligner(0)
buf, offs = synth_batch(100000, 1500, 0.018, 0.002, seed=20261019)
d = torch.from_numpy(np.concatenate([buf, np.zeros(64, np.uint8)])).cuda()
al.plan(offs, al.params(9))
t = []
for i in range(8):
    al.run(d.data_ptr()); r = al.fetch(unpack=False)
    if i >= 3: t.append((r.ms_stage1, r.ms_stage3, r.ms_assemble))
print("stage12, stage3, assemble ms:", np.array(t).min(axis=0))

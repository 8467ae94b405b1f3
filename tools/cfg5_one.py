"""One device-resident run of BASELINE config [4] (1 Mb pair, ten indel-rich bubbles, k = 15) -- the command the
round's ncu launch list of the long-segment kernels is taken on (tools/final_measure.sh)."""
import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dbga_b200 import Aligner
from dbga_b200.synth import synth_long_pair
al = Aligner(0)
buf, offs = synth_long_pair(seed=3)
d = torch.from_numpy(np.concatenate([buf, np.zeros(64, np.uint8)])).cuda()
al.plan(offs, al.params(15))
for i in range(3):
    al.run(d.data_ptr()); r = al.fetch(unpack=False)
print(r.ms_stage3)

"""DP microbenchmark: batches of uniform L x L global alignments
through dbga_global_align_batch, device-resident; GCUPS = cells / stage-3 device time."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dbga_b200 import Aligner
from dbga_b200.synth import synth_batch

def main():
    Ls = [int(x) for x in sys.argv[1:]] or [16, 64, 256, 1024, 4096]
    al = Aligner(0)
    if os.environ.get("DP_SCRATCH_GB"):            # more segments than 4 G cells' worth: the traceback pool is sized at a byte per cell
        al.set_scratch_limit(int(os.environ["DP_SCRATCH_GB"]) << 30)
    peak = al.int_peak()
    roof = peak["mix_ops_per_s"] / 9 / 1e9
    out = []
    for L in Ls:
        cells_target = 4e9 if L >= 256 else 1e9
        P = int(os.environ.get("DP_PAIRS", 0)) or int(max(8, min(400_000, cells_target / (L * L))))
        buf, offs = synth_batch(P, L, 0.05, 0.01, seed=L)
        d = torch.from_numpy(np.concatenate([buf, np.zeros(64, np.uint8)])).cuda()
        prm = al.params(1)
        al.plan(offs, prm, global_only=True)
        for _ in range(2):
            al.run(d.data_ptr()); r = al.fetch(unpack=False)
        t3 = []
        for _ in range(3):
            al.run(d.data_ptr()); r = al.fetch(unpack=False)
            t3.append(r.ms_stage3)
        cells = int(r.dp_cells_total)
        ms = min(t3)
        rec = {"L": L, "pairs": P, "cells": cells, "stage3_ms": ms, "gcups": cells / ms / 1e6,
               "frac_of_int_roofline": cells / ms / 1e6 / roof, "assemble_ms": r.ms_assemble}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    print(json.dumps({"roofline_gcups": roof, "int_peak": peak}))

if __name__ == "__main__":
    main()

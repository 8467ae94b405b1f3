"""Small mixed workload for compute-sanitizer (memcheck / racecheck): every kernel class, tiny sizes."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbga_b200 import Aligner
from oracle import dbga_oracle as O
al = Aligner(0)
rng = random.Random(3)
rs = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
pairs = [O.synth_pair(1500, 0.018, 0.002, s) for s in range(6)] + [O.synth_pair(300, 0.1, 0.03, 9)]
pairs += [(rs(40), rs(35)), (rs(300), rs(280)), ("ACGT", "ACGTA"), (rs(700), rs(650))]
r = al.align_pairs(pairs, 9)
print("batch", r.status.tolist())
r = al.align_pairs(pairs[:3], 17, want_graph=True, stages=2)
print("graph", r.status.tolist())
big = [O.synth_pair(12000, 0.01, 0.002, 4)]
r = al.align_pairs(big + pairs[:2], 11)
print("large path", r.status.tolist())
r = al.global_align_pairs([(rs(600), rs(610)), (rs(20), rs(24)), (rs(70), rs(300))])
print("global", r.status.tolist(), r.score.tolist())

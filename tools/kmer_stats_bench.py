"""Throughput of dbga_kmer_stats (host buffers in, counts out) on the cfg3 batch, next to the
reference-style set arithmetic (oracle, one host thread) on a sample of it."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from dbga_b200 import Aligner
from dbga_b200.synth import synth_batch, pairs_from_buffers
from oracle import dbga_oracle as O

al = Aligner(0)
P, n, k = 100_000, 1500, 9
buf, offs = synth_batch(P, n, 0.018, 0.002, seed=3)
status = np.zeros(P, np.int32); d0 = np.zeros(P, np.int64); d1 = np.zeros(P, np.int64); sh = np.zeros(P, np.int64)
i64 = C.POINTER(C.c_int64)
def run():
    rc = al.lib.dbga_kmer_stats(al.h, C.c_void_p(buf.ctypes.data), offs.ctypes.data_as(i64), P, k,
                                status.ctypes.data_as(C.POINTER(C.c_int32)), d0.ctypes.data_as(i64), d1.ctypes.data_as(i64), sh.ctypes.data_as(i64))
    assert rc == 0
for _ in range(2): run()
t0 = time.perf_counter()
for _ in range(5): run()
dt = (time.perf_counter() - t0) / 5
sample = pairs_from_buffers(buf, offs, list(range(0, P, 100)))
t0 = time.perf_counter()
ref = [O.kmer_stats(a, b, k) for a, b in sample]
cdt = time.perf_counter() - t0
for j, i in enumerate(range(0, P, 100)):
    assert (int(d0[i]), int(d1[i]), int(sh[i])) == ref[j]
print(json.dumps({"workload": f"{P} pairs of {n} nt, k={k}", "gpu_ms_host_to_host": dt * 1e3, "gpu_pairs_per_s": P / dt,
                  "bytes_in": int(offs[-1]), "cpu_oracle_pairs_per_s_one_thread": len(sample) / cdt,
                  "checked_against_oracle": len(sample)}))

"""Per-chunk GPU timeline of one pipelined dbga_align_batch call (DBGA_TRACE=1) on the cfg3 batch.
usage: python tools/e2e_trace.py [flags] [packed]     flags: 1 NO_RUNS, 2 COMPACT_ROWS, 4 SOP"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbga_b200 import Aligner, pack_ascii
from dbga_b200.synth import synth_batch
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 1
packed = len(sys.argv) > 2 and sys.argv[2] == "packed"
al = Aligner(0)
P, n = 100_000, 1500
cap = P * (2 * n + 64)
hptr = al.lib.dbga_host_alloc(cap)
hbuf = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_uint8)), shape=(cap,))
buf, offs = synth_batch(P, n, 0.018, 0.002, seed=1, out=hbuf)
src = hptr
if packed:
    words = pack_ascii(buf[:int(offs[-1])], offs)
    src = al.lib.dbga_host_alloc(words.nbytes)
    np.ctypeslib.as_array(C.cast(src, C.POINTER(C.c_uint32)), shape=(len(words),))[:] = words
prm = al.params(9, flags=flags)
ts = []
for _ in range(6):
    t0 = time.perf_counter()
    al.align_buffers(src, offs, prm, unpack=False, packed=packed)
    ts.append((time.perf_counter() - t0) * 1e3)
print("flags", flags, "packed", packed, "ms per call", [round(t, 2) for t in ts])
os.environ["DBGA_TRACE"] = "1"
t0 = time.perf_counter()
al.align_buffers(src, offs, prm, unpack=False, packed=packed)
print("traced call ms", (time.perf_counter() - t0) * 1e3)
r = al.last_raw
print("stage sums ms:", r.ms_stage1, r.ms_stage2, r.ms_stage3, r.ms_assemble, "plan", r.ms_plan, "launches", r.kernel_launches)

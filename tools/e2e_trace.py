import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbga_b200 import Aligner
from dbga_b200.synth import synth_batch
al = Aligner(0)
P, n = 100_000, 1500
cap = P * (2 * n + 64)
hptr = al.lib.dbga_host_alloc(cap)
hbuf = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_uint8)), shape=(cap,))
buf, offs = synth_batch(P, n, 0.018, 0.002, seed=1, out=hbuf)
prm = al.params(9)
for _ in range(3):
    al.align_buffers(hptr, offs, prm, unpack=False)
os.environ["DBGA_TRACE"] = "1"
t0 = time.perf_counter()
al.align_buffers(hptr, offs, prm, unpack=False)
print("ms", (time.perf_counter() - t0) * 1e3)
r = al.last_raw
print("stage sums ms:", r.ms_stage1, r.ms_stage2, r.ms_stage3, r.ms_assemble, "launches", r.kernel_launches)

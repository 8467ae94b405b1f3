"""One host process, N GPUs: dbga_align_batch_multi on the cfg3 batch (100 000 pairs sharded over the devices,
BASELINE config [2] "sharded by pair across 1/2/4/8 GPUs"), host wall clock around the call, pinned host
input, for the wire formats of bench.py.  One JSON line per (N, format)."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dbga_b200 import Aligner, MultiAligner, _lib, pack_ascii
from dbga_b200.synth import synth_batch

ndev = torch.cuda.device_count()
al = Aligner(0)
lib = al.lib
P, n = 100_000, 1500
cap = P * (2 * n + 64)
hptr = lib.dbga_host_alloc(cap)
hbuf = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_uint8)), shape=(cap,))
buf, offs = synth_batch(P, n, 0.018, 0.002, seed=20261019, out=hbuf)
total = int(offs[-1])
words = pack_ascii(buf[:total], offs)
wptr = lib.dbga_host_alloc(words.nbytes)
np.ctypeslib.as_array(C.cast(wptr, C.POINTER(C.c_uint32)), shape=(len(words),))[:] = words
ref = al.align_buffers(hptr, offs, al.params(9, flags=_lib.FLAG_NO_RUNS))
al.close()
for N in [x for x in (1, 2, 4, 8) if x <= ndev]:
    m = MultiAligner(list(range(N)))
    for name, flags, packed in (("ascii in, rows out (NO_RUNS)", 1, False), ("ascii in, compact rows", 2, False),
                                ("2-bit in, rows out (NO_RUNS)", 1, True), ("2-bit in, compact rows", 2, True)):
        prm = Aligner.params(9, flags=flags)
        for _ in range(3):
            res = m.align_buffers(None if packed else hptr, offs, prm, packed=wptr if packed else None, unpack=False)
        t0 = time.perf_counter()
        K = 6
        for _ in range(K):
            res = m.align_buffers(None if packed else hptr, offs, prm, packed=wptr if packed else None, unpack=False)
        dt = (time.perf_counter() - t0) / K
        rec = {"n_gpus": N, "format": name, "ms_per_call": dt * 1e3, "pairs_per_s": P / dt, "scaling": "strong (100 000 pairs sharded)"}
        if flags == 1:      # shards side by side == the single-GPU result
            sh = m.align_buffers(None if packed else hptr, offs, prm, packed=wptr if packed else None)
            rec["equal_to_single_gpu"] = bool(np.array_equal(np.concatenate([r.aln0 for _, r in sh]), ref.aln0) and
                                              np.array_equal(np.concatenate([r.status for _, r in sh]), ref.status))
        print(json.dumps(rec), flush=True)
    m.close()

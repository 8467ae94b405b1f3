"""File to file: a FASTA file of the cfg3 batch (100 000 pairs of 1.5 kb, 200 000 records) -> dbga_fasta_read
(pinned ASCII + 2-bit words) -> dbga_align_batch_packed with COMPACT_ROWS -> dbga_fasta_write (60-column
records).  Prints one JSON line with the three phases (host wall clock) and the size of both files."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbga_b200 import Aligner, _lib
from dbga_b200.fasta import FastaBatch
from dbga_b200.synth import synth_batch

P = int(os.environ.get("PAIRS", "100000"))
d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
src, dst = os.path.join(d, "dbga_cfg3_pairs.fasta"), os.path.join(d, "dbga_cfg3_aligned.fasta")
buf, offs = synth_batch(P, 1500, 0.018, 0.002, seed=20261019)
t0 = time.perf_counter()
with open(src, "wb") as f:                      # 70-column input lines
    for r in range(2 * P):
        row = buf[int(offs[r]):int(offs[r + 1])]
        f.write(b">p%d_%d\n" % (r // 2, r % 2))
        n = len(row)
        body = np.full(n + (n + 69) // 70, 10, np.uint8)
        idx = np.arange(n)
        body[idx + idx // 70] = row
        f.write(body.tobytes())
gen_s = time.perf_counter() - t0
al = Aligner(0)
prm = al.params(9, flags=_lib.FLAG_COMPACT_ROWS)
best = None
for it in range(4):
    t0 = time.perf_counter()
    fa = FastaBatch(src, want_packed=True)
    t1 = time.perf_counter()
    raw = al.align_buffers(fa.packed_ptr, np.ascontiguousarray(fa.offsets), prm, unpack=False, packed=True)
    t2 = time.perf_counter()
    n = fa.write(dst, raw)
    t3 = time.perf_counter()
    rec = {"pairs": P, "pairs_written": n, "read_ms": (t1 - t0) * 1e3, "align_ms": (t2 - t1) * 1e3, "write_ms": (t3 - t2) * 1e3,
           "total_ms": (t3 - t0) * 1e3, "pairs_per_s_file_to_file": P / (t3 - t0), "in_bytes": os.path.getsize(src),
           "out_bytes": os.path.getsize(dst), "host_threads": os.cpu_count()}
    if it and (best is None or rec["total_ms"] < best["total_ms"]):
        best = rec
    fa.close()
# spot check: the first aligned pair's records degap to the input
txt = open(dst, "rb").read(20000).split(b">")[1:3]
rows = [b"".join(t.split(b"\n")[1:]) for t in txt]
i = int(txt[0].split(b"_")[0][1:])
best["first_pair_degaps_to_input"] = bool(rows[0].replace(b"-", b"") == buf[int(offs[2 * i]):int(offs[2 * i + 1])].tobytes())
print(json.dumps(best))
os.remove(src); os.remove(dst)

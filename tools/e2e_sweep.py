"""End-to-end (pinned host -> C ABI -> pinned host) timing sweep over pipeline knobs."""
import ctypes as C, os, subprocess, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KEYS = ("DBGA_FIRST_MB", "DBGA_CHUNK_MB", "DBGA_SUB_MB")
if len(sys.argv) > 1 and sys.argv[1] == "child":
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
    t0 = time.perf_counter()
    for _ in range(8):
        r = al.align_buffers(hptr, offs, prm, unpack=False)
    dt = (time.perf_counter() - t0) / 8 * 1e3
    print(json.dumps({k: os.environ.get(k) for k in KEYS} | {"ms": round(dt, 3), "launches": int(r.kernel_launches)}))
else:
    for first, mb, sub in [(8, 32, 16), (4, 32, 16), (2, 32, 16), (4, 16, 16), (4, 24, 8), (8, 64, 16), (4, 32, 32), (4, 32, 8), (2, 16, 8)]:
        env = dict(os.environ, DBGA_FIRST_MB=str(first), DBGA_CHUNK_MB=str(mb), DBGA_SUB_MB=str(sub))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)

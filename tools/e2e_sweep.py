"""End-to-end (pinned host -> C ABI -> pinned host) timing sweep over pipeline knobs."""
import ctypes as C, os, subprocess, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    print(json.dumps({"chunk_mb": os.environ.get("DBGA_CHUNK_MB"), "lag": os.environ.get("DBGA_LAG"), "ms": round(dt, 3),
                      "launches": int(r.kernel_launches)}))
else:
    for mb in (16, 24, 32, 48, 64):
        for lag in (1, 2, 3):
            env = dict(os.environ, DBGA_CHUNK_MB=str(mb), DBGA_LAG=str(lag))
            out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])

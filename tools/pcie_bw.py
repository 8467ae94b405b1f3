import torch, time, numpy as np
n = 304_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("H2D ms", t(lambda: d.copy_(h, non_blocking=True)), "GB/s", n / t(lambda: d.copy_(h, non_blocking=True)) / 1e6)
print("D2H ms", t(lambda: h2.copy_(d2, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("H2D+D2H concurrent ms", t(both))
def chunked(k=10):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h2[i*c:(i+1)*c].copy_(d2[i*c:(i+1)*c], non_blocking=True)
print("chunked concurrent ms", t(chunked))
print("chunked x20 concurrent ms", t(lambda: chunked(20)))
print("chunked x40 concurrent ms", t(lambda: chunked(40)))
# the same with an SM-heavy kernel running next to the copies
s3 = torch.cuda.Stream()
a = torch.randn(8192, 8192, device="cuda")
def both_busy():
    with torch.cuda.stream(s3):
        for _ in range(6): torch.mm(a, a)
    both()
print("concurrent + GEMMs ms", t(both_busy))
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
def both_membound():
    with torch.cuda.stream(s3):
        for _ in range(12): x.mul_(1.0001)
    both()
print("concurrent + HBM-bound kernels ms", t(both_membound))

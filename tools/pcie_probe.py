"""What the host link of this box can do (development tool): pinned H2D alone, D2H alone, both at once, and the same
in 64 pieces per direction -- the copy pattern of asr/loss/host.py."""
import time, torch
n = 524_000_000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.zeros(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d(pieces=1):
    with torch.cuda.stream(s1):
        for c in range(pieces):
            a, b = c * n // pieces, (c + 1) * n // pieces
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
def d2h(pieces=1):
    with torch.cuda.stream(s2):
        for c in range(pieces):
            a, b = c * n // pieces, (c + 1) * n // pieces
            h_out[a:b].copy_(d_out[a:b], non_blocking=True)
for pieces in (1, 64):
    t = timed(lambda: h2d(pieces)); print("H2D alone   %2d pieces: %6.2f ms  %5.1f GB/s" % (pieces, t * 1e3, n * 4 / t / 1e9))
    t = timed(lambda: d2h(pieces)); print("D2H alone   %2d pieces: %6.2f ms  %5.1f GB/s" % (pieces, t * 1e3, n * 4 / t / 1e9))
    t = timed(lambda: (h2d(pieces), d2h(pieces))); print("both at once %2d pieces: %6.2f ms  %5.1f GB/s per direction" % (pieces, t * 1e3, n * 4 / t / 1e9))
import numpy as np, concurrent.futures
g = h_out.numpy()
pool = concurrent.futures.ThreadPoolExecutor(8)
def zero(i):
    a, b = i * (n // 64), i * (n // 64) + (n // 64) * 27 // 100
    g[a:b] = 0.0
t0 = time.perf_counter(); list(pool.map(zero, range(64))); t = time.perf_counter() - t0
print("host zeroing of 27%% of the buffer, 8 threads: %.2f ms" % (t * 1e3))
t0 = time.perf_counter(); [zero(i) for i in range(64)]; t = time.perf_counter() - t0
print("host zeroing of 27%% of the buffer, 1 thread : %.2f ms" % (t * 1e3))

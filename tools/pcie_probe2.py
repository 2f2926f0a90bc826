"""Host link, second look (development tool): the end-to-end path moves one utterance (~8 MB) per copy, 64 copies per
direction, both directions at once.  Does spreading the copies of one direction over several streams (several copy
engines) close the gap to two single large copies?  Sizes follow the bench workload (valid rows: 524 MB each way)."""
import sys, time, torch
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 0          # 1: short list (for N processes at once)
n = 524_000_000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.zeros(n, dtype=torch.float32, device="cuda")
S_in = [torch.cuda.Stream() for _ in range(8)]; S_out = [torch.cuda.Stream() for _ in range(8)]
def timed(fn, reps=8 if ns else 4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d(pieces, streams):
    for c in range(pieces):
        a, b = c * n // pieces, (c + 1) * n // pieces
        with torch.cuda.stream(S_in[c % streams]):
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
def d2h(pieces, streams):
    for c in range(pieces):
        a, b = c * n // pieces, (c + 1) * n // pieces
        with torch.cuda.stream(S_out[c % streams]):
            h_out[a:b].copy_(d_out[a:b], non_blocking=True)
def both(pieces, streams):          # interleave the enqueues the way the pipeline does
    for c in range(pieces):
        a, b = c * n // pieces, (c + 1) * n // pieces
        with torch.cuda.stream(S_in[c % streams]):
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        with torch.cuda.stream(S_out[c % streams]):
            h_out[a:b].copy_(d_out[a:b], non_blocking=True)
for pieces in ((1, 64) if ns else (1, 16, 64, 256)):
    for streams in ((1,) if ns else (1, 2, 4)):
        if pieces == 1 and streams > 1: continue
        t1 = timed(lambda: h2d(pieces, streams)); t2 = timed(lambda: d2h(pieces, streams)); t3 = timed(lambda: both(pieces, streams))
        print("%3d pieces, %d stream(s)/direction: H2D alone %5.1f GB/s  D2H alone %5.1f GB/s  both at once %6.2f ms = %5.1f GB/s per direction"
              % (pieces, streams, n * 4 / t1 / 1e9, n * 4 / t2 / 1e9, t3 * 1e3, n * 4 / t3 / 1e9), flush=True)

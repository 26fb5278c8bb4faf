"""Developer tool: raw pinned-memory copy rates of the box (one direction alone, both at once) -- the ceiling
of the host-buffer path."""
import time, torch
n = 600_000_000
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_out = torch.empty(n // 2, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.empty(n // 2, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=25):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(chunks):
        a, b = k * (n // chunks), (k + 1) * (n // chunks)
        if h2d:
            with torch.cuda.stream(s1): d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[a // 2:b // 2].copy_(d_out[a // 2:b // 2], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return (n * 4 / dt / 1e9 if h2d else 0.0), (n * 2 / dt / 1e9 if d2h else 0.0)
for _ in range(2):
    print("H2D alone %.1f GB/s" % run(True, False)[0], "| D2H alone %.1f GB/s" % (run(False, True)[1]),
          "| together: H2D %.1f + D2H (half the bytes) %.1f GB/s" % run(True, True))

# the same pair of copies while kernels stream through HBM on a third stream (the pipeline's situation)
s3 = torch.cuda.Stream()
big = torch.empty(1_000_000_000, dtype=torch.float32, device="cuda")
def run_busy(chunks=25):
    stop = torch.zeros(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s3):
        for _ in range(12): big.add_(1.0)          # ~8 GB of traffic each, ~15 ms of kernels in total
    for k in range(chunks):
        a, b = k * (n // chunks), (k + 1) * (n // chunks)
        with torch.cuda.stream(s1): d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        with torch.cuda.stream(s2): h_out[a // 2:b // 2].copy_(d_out[a // 2:b // 2], non_blocking=True)
    s1.synchronize(); dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return n * 4 / dt / 1e9
for _ in range(2):
    print("H2D with a half-size D2H and HBM-streaming kernels beside it: %.1f GB/s" % run_busy())

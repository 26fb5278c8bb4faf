"""Developer probe (GPU box): how fast can spectra be spilled to host memory and read back?
free RAM, cores, cudaHostAlloc rate, pinned D2H / H2D rate, pageable first-touch + copy rates."""
import os, time, sys
import torch
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
print(open("/proc/meminfo").read().split("\n")[0:3])
dev = torch.device("cuda", 0)
x = torch.empty(1 << 30, dtype=torch.uint8, device=dev); x.fill_(3)   # 1 GiB
torch.cuda.synchronize()
for gb in (1, 4, 16):
    t = time.perf_counter(); h = torch.empty(gb << 30, dtype=torch.uint8, pin_memory=True); dt = time.perf_counter() - t
    print(f"pin {gb} GiB: {dt:.3f} s = {gb/dt:.2f} GiB/s", flush=True)
    if gb < 16: del h
# pinned D2H / H2D of 16 GiB in 1 GiB pieces
t = time.perf_counter()
for i in range(16): h[i << 30:(i + 1) << 30].copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t; print(f"D2H pinned 16 GiB {16/dt:.1f} GiB/s")
t = time.perf_counter()
for i in range(16): x.copy_(h[i << 30:(i + 1) << 30], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t; print(f"H2D pinned 16 GiB {16/dt:.1f} GiB/s")
# pinning while a kernel loop runs: does cudaHostAlloc stall launches?
import threading
stop = False; n = [0]
def launcher():
    y = torch.zeros(1 << 20, device=dev)
    while not stop:
        y.add_(1.0); n[0] += 1
    torch.cuda.synchronize()
th = threading.Thread(target=launcher); th.start(); time.sleep(0.5)
n0 = n[0]; t = time.perf_counter(); h2 = torch.empty(8 << 30, dtype=torch.uint8, pin_memory=True); dt = time.perf_counter() - t
n1 = n[0]; time.sleep(dt); n2 = n[0]; stop = True; th.join()
print(f"pin 8 GiB beside a launch loop: {dt:.2f} s; launches during {n1-n0}, same time after {n2-n1}")
del h2
# pageable: first touch and copy from pinned
p = torch.empty(16 << 30, dtype=torch.uint8)
t = time.perf_counter(); p.copy_(h); dt = time.perf_counter() - t; print(f"pinned->pageable first touch 16 GiB: {16/dt:.1f} GiB/s (torch threads {torch.get_num_threads()})")
t = time.perf_counter(); p.copy_(h); dt = time.perf_counter() - t; print(f"pinned->pageable second 16 GiB: {16/dt:.1f} GiB/s")
t = time.perf_counter(); h.copy_(p); dt = time.perf_counter() - t; print(f"pageable->pinned 16 GiB: {16/dt:.1f} GiB/s")
t = time.perf_counter()
for i in range(4): x.copy_(p[i << 30:(i + 1) << 30])
torch.cuda.synchronize(); dt = time.perf_counter() - t; print(f"H2D pageable 4 GiB {4/dt:.1f} GiB/s")
# cudaHostRegister of existing pageable memory
t = time.perf_counter(); r = torch.cuda.cudart().cudaHostRegister(p.data_ptr(), 16 << 30, 0); dt = time.perf_counter() - t
print(f"cudaHostRegister 16 GiB (touched): rc {r} {dt:.2f} s = {16/dt:.1f} GiB/s")

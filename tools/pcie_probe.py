"""Development aid: pinned-memory PCIe copy rates on the GPU box (the floor of the end-to-end call: 6.2 MB up, 24.9 MB down per 4K frame)."""
import time, torch
for mb, name in ((24.8832, "D2H 24.9 MB"), (6.2208, "H2D 6.2 MB"), (4.0, "D2H 4 MB")):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    src, dst = (d, h) if name.startswith("D2H") else (h, d)
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
    print("%-12s %.3f ms  %.1f GB/s" % (name, dt * 1e3, n / dt / 1e9))
# both directions at once
n1, n2 = 24883200, 6220800
h1 = torch.empty(n1, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n1, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n2, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
print("D2H 24.9 MB + H2D 6.2 MB concurrently: %.3f ms" % (dt * 1e3))

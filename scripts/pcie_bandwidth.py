import torch, time
n = 900 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
h2 = torch.empty(n // 2, dtype=torch.float32).pin_memory()
d2 = torch.empty(n // 2, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(name, "%.1f GB/s" % (n * 4 / dt / 1e9))
# both directions at once
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("duplex: D2H %.1f GB/s + H2D %.1f GB/s" % (n * 4 / dt / 1e9, n * 2 / dt / 1e9))
# small chunks 36 MB
c = 36 * 1024 * 1024 // 4
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(25): h[i * c:(i + 1) * c].copy_(d[i * c:(i + 1) * c], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H 25 x 36 MB: %.1f GB/s" % (25 * c * 4 / dt / 1e9))

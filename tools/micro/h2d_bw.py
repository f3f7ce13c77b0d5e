import torch, time
n = 2 * 1024**3
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("1D H2D 2GiB: %.2f GB/s" % (n / dt / 1e9))
# two streams, halves
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): d[n // 2:].copy_(h[n // 2:], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("2-stream H2D 2GiB: %.2f GB/s" % (n / dt / 1e9))

"""Ad-hoc: raw pinned H2D / D2H bandwidth of the box (one big copy, many 1 MB copies, both directions at once)."""
import torch, time
dev = "cuda:0"
n = 295 * 1000 * 1000 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
h2 = torch.empty(n // 2, dtype=torch.float32).pin_memory()
d2 = torch.empty(n // 2, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, K=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3
ms = t(lambda: d.copy_(h, non_blocking=True)); print("H2D 295 MB one copy: %.2f ms %.1f GB/s" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H 147 MB one copy: %.2f ms %.1f GB/s" % (ms, n * 2 / ms / 1e6))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
ms = t(both); print("H2D 295 MB + D2H 147 MB concurrently: %.2f ms" % ms)
c = 288000
def chunks():
    for i in range(0, n - c, c): d[i:i + c].copy_(h[i:i + c], non_blocking=True)
ms = t(chunks); print("H2D 295 MB in 1.15 MB copies: %.2f ms %.1f GB/s" % (ms, n * 4 / ms / 1e6))

"""Pinned-memory copy rates of the box: H2D alone, D2H alone, both at once (the ceiling of any host-buffer pipeline).
    python tools/pcie_probe.py"""
import torch, time
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9
run(1, 1, 1)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once %.1f GB/s each" % (run(1, 0), run(0, 1), run(1, 1)))

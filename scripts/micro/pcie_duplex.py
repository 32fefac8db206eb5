"""PCIe copy bandwidth of the box: H2D alone, D2H alone, both at once (pinned, 1 GiB each)."""
import time
import torch
n = 1 << 27   # doubles: 1 GiB
h1 = torch.empty(n, dtype=torch.float64, pin_memory=True); h1.zero_()
h2 = torch.empty(n, dtype=torch.float64, pin_memory=True); h2.zero_()
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n * 8 / dt / 1e9
run(True, True, 1)
print("H2D alone  %.1f GB/s" % run(True, False))
print("D2H alone  %.1f GB/s" % run(False, True))
print("both: each %.1f GB/s (sum %.1f)" % (run(True, True), 2 * run(True, True)))

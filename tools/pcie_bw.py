import torch, time
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return n * reps / dt / 1e9
run(True, True, 2)
print("H2D only  %.1f GB/s" % run(True, False))
print("D2H only  %.1f GB/s" % run(False, True))
print("both      %.1f GB/s per direction" % run(True, True))

"""What the e2e path can get out of PCIe: the copy pattern of one encode+decode pair (H2D 16.6 MB frame + ~5.5 MB sparse
coefficients, D2H the same two sizes), no kernels, `k` streams per direction, pinned host buffers cycled through a
ring as bench.py does.  Prints the pairs per second the copy engines sustain = the ceiling of bench.py's e2e."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("BIND", "1") == "1":          # pinned buffers on the GPU's own NUMA node, as bench.py places them
    importlib.import_module("cineform-sdk_b200").bind_thread_to_device(0)
FRAME, SPARSE = 3840 * 2160 * 2, int(sys.argv[2]) if len(sys.argv) > 2 else 5_500_000
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ring = 48
hf = [torch.empty(FRAME, dtype=torch.uint8).pin_memory() for _ in range(ring)]
hs = [torch.empty(SPARSE, dtype=torch.uint8).pin_memory() for _ in range(ring)]
ho = [torch.empty(FRAME, dtype=torch.uint8).pin_memory() for _ in range(ring)]
hs2 = [torch.empty(SPARSE, dtype=torch.uint8).pin_memory() for _ in range(ring)]
df = [torch.empty(FRAME, dtype=torch.uint8, device="cuda") for _ in range(8)]
ds = [torch.empty(SPARSE, dtype=torch.uint8, device="cuda") for _ in range(8)]
up = [torch.cuda.Stream() for _ in range(k)]
down = [torch.cuda.Stream() for _ in range(k)]
def run(n):
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(n):
        r, d = i % ring, i % 8
        with torch.cuda.stream(up[i % k]):
            df[d].copy_(hf[r], non_blocking=True)          # encode: frame up
            ds[d].copy_(hs[r], non_blocking=True)          # decode: coefficients up
        with torch.cuda.stream(down[i % k]):
            hs2[r].copy_(ds[d], non_blocking=True)         # encode: coefficients down
            ho[r].copy_(df[d], non_blocking=True)          # decode: frame down
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t)
run(50)
fps = run(600)
print(f"{k} stream(s) per direction, sparse {SPARSE/1e6:.1f} MB: {fps:.0f} pairs/s = {fps * (FRAME + SPARSE) / 1e9:.1f} GB/s per direction")

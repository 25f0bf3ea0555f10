"""Where does the e2e time go?  Measures time blocked in pool.wait() vs the Python submit loop."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module("cineform-sdk_b200")
W, H = 3840, 2160
pkg.bind_thread_to_device(0)        # as bench.py: pinned buffers on the GPU's own NUMA node
desc = pkg.FrameDesc(W, H, pkg.PIXEL_YUYV); quant = pkg.quant_for_quality(desc, 4)
frames = bench.synthetic_frames(16, W, H)
slots, batch = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
pool = pkg.Pool([0], desc, slots=slots, batch=batch, queue_length=64)
lay = pool.layout
ring = 96
h_in = [pkg.pinned_empty((H, lay.frame_pitch)) for _ in range(ring)]
h_cd = [pkg.pinned_empty(pkg.sparse_max_bytes(lay)) for _ in range(ring)]
h_out = [pkg.pinned_empty((H, lay.frame_pitch)) for _ in range(ring)]
for i in range(ring): h_in[i][:] = frames[i % 16]
def run(total, sparse=True):
    sub_f = pool.submit_forward_sparse if sparse else pool.submit_forward
    sub_i = pool.submit_inverse_sparse if sparse else pool.submit_inverse
    INV = 1 << 30; next_f = done = inflight = 0; tw = 0.0
    while done < total:
        while next_f < total and inflight < 40 and next_f - done < ring - 8:
            sub_f(next_f, h_in[next_f % ring], quant, h_cd[next_f % ring]); next_f += 1; inflight += 1
        t = time.perf_counter(); r = pool.wait(); tw += time.perf_counter() - t; inflight -= 1
        if r & INV or mode == "fwd": done += 1
        else:
            sub_i(INV | r, h_cd[r % ring], quant, pkg.PIXEL_YUYV, h_out[r % ring]); inflight += 1
    return tw
run(64)
n = 480
t0 = time.perf_counter(); tw = run(n); dt = time.perf_counter() - t0
print(f"slots {slots} batch {batch} mode {mode}: {n/dt:.0f} fps; blocked in wait {tw/dt*100:.0f}% of wall; python loop {(dt-tw)/n*1e6:.0f} us/frame")
pool.close()

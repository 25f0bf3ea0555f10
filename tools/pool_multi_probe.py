"""ONE frame pool over several GPUs in ONE process (the north star's single EncoderPool queue, cfb_pool_create(devices =
[0..N-1])): end-to-end encode+decode fps with host buffers, sparse hand-over, jobs round-robin over the devices,
in-order delivery.  Usage: python tools/pool_multi_probe.py <ndevices> [slots] [batch]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("cineform-sdk_b200")
W, H = 3840, 2160
nd = int(sys.argv[1]); slots = int(sys.argv[2]) if len(sys.argv) > 2 else 8; batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
desc = pkg.FrameDesc(W, H, pkg.PIXEL_YUYV); quant = pkg.quant_for_quality(desc, 4)
frames = bench.synthetic_frames(16, W, H)
inflight_max = 40 * nd
pool = pkg.Pool(list(range(nd)), desc, slots=slots, batch=batch, queue_length=inflight_max + 16)
lay = pool.layout
ring = ((inflight_max + 24 + nd - 1) // nd) * nd        # a multiple of the device count: ring slot r always serves device r % nd
numa = len(sys.argv) <= 4 or sys.argv[4] != "nonuma"
h_in, h_cd, h_out = [None] * ring, [None] * ring, [None] * ring
for d in range(nd):
    # the pool borrows the caller's buffers (as CEncoderPool does), so their NUMA placement is the caller's: job i goes to device
    # i % nd (round-robin, EncoderPool.cpp:284), hence the buffers of ring slot r are allocated on the node of GPU r % nd
    if numa: pkg.bind_thread_to_device(d)
    for r in range(d, ring, nd):
        h_in[r] = pkg.pinned_empty((H, lay.frame_pitch)); h_in[r][:] = frames[r % 16]
        h_cd[r] = pkg.pinned_empty(pkg.sparse_max_bytes(lay))
        h_out[r] = pkg.pinned_empty((H, lay.frame_pitch)); h_out[r][:] = 0
os.sched_setaffinity(0, range(os.cpu_count()))
def run(total):
    INV = 1 << 30; next_f = done = inflight = 0
    while done < total:
        while next_f < total and inflight < inflight_max and next_f - done < ring - 8:
            pool.submit_forward_sparse(next_f, h_in[next_f % ring], quant, h_cd[next_f % ring]); next_f += 1; inflight += 1
        r = pool.wait(); inflight -= 1
        if r & INV: done += 1
        else:
            pool.submit_inverse_sparse(INV | r, h_cd[r % ring], quant, pkg.PIXEL_YUYV, h_out[r % ring]); inflight += 1
run(64 * nd)
n = 600 * nd
t0 = time.perf_counter(); run(n); dt = time.perf_counter() - t0
ok = bench.psnr(h_out[0][:, 0::2], h_in[0][:, 0::2]) > 40.0
print(f"one pool over {nd} GPU(s), {slots} slots x batch {batch} each, host buffers {'on the node of their GPU' if numa else 'wherever the main thread allocated them'}, one submit/wait thread: {n / dt:.0f} fps ({n / dt / nd:.0f} per GPU), round trip ok {ok}", flush=True)
pool.close()

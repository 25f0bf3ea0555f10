"""Timeline of the pooled e2e path from CUPTI (torch.profiler): how busy each PCIe direction is, at what rate the copies
run, and how much SM time the kernels take.  Development tool: python tools/e2e_trace.py <slots> <batch>"""
import importlib, os, sys, time, json, collections
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("cineform-sdk_b200")
W, H = 3840, 2160
pkg.bind_thread_to_device(0)
desc = pkg.FrameDesc(W, H, pkg.PIXEL_YUYV); quant = pkg.quant_for_quality(desc, 4)
frames = bench.synthetic_frames(16, W, H)
slots, batch = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.init()
pool = pkg.Pool([0], desc, slots=slots, batch=batch, queue_length=80)
lay = pool.layout
ring = 96
h_in = [pkg.pinned_empty((H, lay.frame_pitch)) for _ in range(ring)]
h_cd = [pkg.pinned_empty(pkg.sparse_max_bytes(lay)) for _ in range(ring)]
h_out = [pkg.pinned_empty((H, lay.frame_pitch)) for _ in range(ring)]
for i in range(ring): h_in[i][:] = frames[i % 16]
def run(total):
    INV = 1 << 30; next_f = done = inflight = 0
    while done < total:
        while next_f < total and inflight < 64 and next_f - done < ring - 8:
            pool.submit_forward_sparse(next_f, h_in[next_f % ring], quant, h_cd[next_f % ring]); next_f += 1; inflight += 1
        r = pool.wait(); inflight -= 1
        if r & INV: done += 1
        else:
            pool.submit_inverse_sparse(INV | r, h_cd[r % ring], quant, pkg.PIXEL_YUYV, h_out[r % ring]); inflight += 1
run(64)
n = 320
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); run(n); dt = time.perf_counter() - t0
print(f"{n / dt:.0f} fps under the profiler")
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
groups = collections.defaultdict(list)
for e in ev:
    name = e.name
    key = "H2D" if "HtoD" in name else "D2H" if "DtoH" in name else "kernel" if "Memcpy" not in name and "Memset" not in name else "other"
    groups[key].append((e.time_range.start, e.time_range.end, name))
t_lo = min(s for g in groups.values() for s, _, _ in g); t_hi = max(e for g in groups.values() for _, e, _ in g)
span = t_hi - t_lo
print(f"span {span / 1e3:.1f} ms")
for key, lst in groups.items():
    lst.sort()
    busy = 0; cur_s, cur_e = lst[0][0], lst[0][1]
    for s, e, _ in lst[1:]:
        if s <= cur_e: cur_e = max(cur_e, e)
        else: busy += cur_e - cur_s; cur_s, cur_e = s, e
    busy += cur_e - cur_s
    tot = sum(e - s for s, e, _ in lst)
    print(f"{key}: {len(lst)} ops, busy {busy / span * 100:.1f}% of the span, summed duration {tot / 1e3:.1f} ms")
    if key in ("H2D", "D2H"):
        durs = sorted((e - s) for s, e, _ in lst)
        print(f"   copy durations us: min {durs[0]:.0f} median {durs[len(durs)//2]:.0f} max {durs[-1]:.0f}")
    if key == "kernel":
        per = collections.defaultdict(float)
        for s, e, nme in lst: per[nme.split('(')[0][:50]] += e - s
        for k_, v in sorted(per.items(), key=lambda x: -x[1])[:12]: print(f"   {k_}: {v / n:.1f} us per frame pair")
pool.close()

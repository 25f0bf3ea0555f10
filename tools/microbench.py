"""Quick device-resident timing of the forward/inverse pyramids (development tool, not the judged bench)."""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--inverse", action="store_true")
    ap.add_argument("--interlaced", action="store_true", help="field transform at level 1")
    ap.add_argument("--resolution", type=int, default=1, help="decode resolution: 1 full, 2 half, 3 quarter")
    a = ap.parse_args()
    pkg = importlib.import_module("cineform-sdk_b200")
    torch.cuda.init()
    ctx = pkg.Context(0)
    desc = pkg.FrameDesc(a.width, a.height, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4, interlaced=a.interlaced)
    codec = pkg.Codec(ctx, desc, 1)
    if a.interlaced:
        codec.set_interlaced(True)
    codec.set_decode_resolution(a.resolution)
    lay = codec.layout
    rng = np.random.default_rng(0)
    frame = pu.synthetic_yuyv(rng, a.width, a.height, "natural")
    stream = torch.cuda.ExternalStream(ctx.stream)
    n = a.batch
    with torch.cuda.stream(stream):
        d_frames = [torch.from_numpy(np.roll(frame, i * 7, axis=0)).cuda() for i in range(n)]
        d_pyr = [torch.zeros(lay.total_bytes, dtype=torch.uint8, device="cuda") for _ in range(n)]
        d_out = [torch.zeros(lay.frame_bytes, dtype=torch.uint8, device="cuda") for _ in range(n)]
    fp = [t.data_ptr() for t in d_frames]
    pp = [t.data_ptr() for t in d_pyr]
    op = [t.data_ptr() for t in d_out]
    algo_fwd = (lay.frame_bytes + 2 * a.width * a.height * 2) + a.width * a.height * 2 + a.width * a.height * 2 // 4
    for name in (["forward", "inverse"] if a.inverse else ["forward"]):
        def run():
            if name == "forward":
                codec.forward_device(fp, lay.frame_pitch, quant, pp)
            else:
                codec.inverse_device(pp, quant, pkg.PIXEL_YUYV, op, lay.frame_pitch)
        for _ in range(3):
            run()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(a.iters):
                run()
            e1.record(stream)
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        per_frame_us = ms * 1000 / n
        print(f"{name}: batch {n}: {ms:.3f} ms/batch, {per_frame_us:.1f} us/frame, {n / ms * 1000:.0f} frames/s, "
              f"{algo_fwd * n / ms / 1e6:.0f} GB/s algorithmic (pyramid {algo_fwd/1e6:.1f} MB/frame)")
    print("launches", ctx.stats())


if __name__ == "__main__":
    main()

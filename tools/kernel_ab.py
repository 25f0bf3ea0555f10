"""Per-kernel A/B timing (development + the evidence behind profiles/*_ab.txt): times ONE pyramid level of the forward
or inverse path, device-resident, 16 x 4K frames per launch, CUDA events on the launching stream.  Kernel variants are
selected by environment variables read by the library (CFB_FWD422, CFB_INV422, CFB_TH ...), so each variant runs in its
own process:   python tools/kernel_ab.py --level 1 --dir fwd"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--dir", default="fwd", choices=["fwd", "inv"])
    ap.add_argument("--format", default="YUYV")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    pkg = importlib.import_module("cineform-sdk_b200")
    sys.path.insert(0, ROOT)
    import bench
    torch.cuda.init()
    ctx = pkg.Context(0)
    fmt = getattr(pkg, "PIXEL_" + a.format)
    desc = pkg.FrameDesc(a.width, a.height, fmt)
    quant = pkg.quant_for_quality(desc, 4)
    codec = pkg.Codec(ctx, desc, 1)
    lay = codec.layout
    stream = torch.cuda.ExternalStream(ctx.stream)
    n = a.batch
    rng = np.random.default_rng(0)
    if a.format == "YUYV":
        frames = bench.synthetic_frames(n, a.width, a.height)
    else:
        base = rng.integers(0, 65536, (lay.frame_bytes // 2,), dtype=np.uint16)
        base = (base & 0xfff0).astype(np.uint16)
        frames = [np.roll(base, 1024 * i).view(np.uint8) for i in range(n)]
    with torch.cuda.stream(stream):
        d_frames = [torch.from_numpy(np.ascontiguousarray(f).reshape(-1).view(np.uint8)).cuda() for f in frames]
        d_pyr = [torch.zeros(lay.total_bytes, dtype=torch.uint8, device="cuda") for _ in range(n)]
        d_out = [torch.zeros(max(lay.frame_bytes, 3 * a.width * a.height * 2), dtype=torch.uint8, device="cuda") for _ in range(n)]
    fp, pp, op = [t.data_ptr() for t in d_frames], [t.data_ptr() for t in d_pyr], [t.data_ptr() for t in d_out]
    # a full forward first so that every level has real input
    codec.forward_device(fp, lay.frame_pitch, quant, pp)
    ctx.synchronize()
    bit = 1 << (a.level - 1)
    out_fmt = pkg.PIXEL_YUYV if a.format == "YUYV" else pkg.PIXEL_PLANAR16
    out_pitch = lay.frame_pitch if a.format == "YUYV" else a.width * 2
    if a.dir == "fwd":
        codec.set_level_mask(bit, 0)
        run = lambda: codec.forward_device(fp, lay.frame_pitch, quant, pp)
    else:
        codec.set_level_mask(0, bit)
        run = lambda: codec.inverse_device(pp, quant, out_fmt, op, out_pitch)
    for _ in range(5):
        run()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.iters):
        run()
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    # algorithmic bytes of this level: level 1 = input frame + 2P, level k = 2P / 4^(k-1) read+write ... all channels
    P = sum(lay.band[c][0][0].width * lay.band[c][0][0].height * 4 for c in range(lay.num_channels))     # samples of all channels
    algo = (lay.frame_bytes + 2 * P) if a.level == 1 else (4 * P // (4 ** (a.level - 1)))
    gbs = algo * n / (ms * 1e-3) / 1e9
    env = {k: v for k, v in os.environ.items() if k.startswith("CFB_")}
    print(f"{a.tag or a.dir + str(a.level)} {a.format} {env}: {ms * 1000:.1f} us per {n}-frame launch, {gbs:.0f} GB/s algorithmic "
          f"({gbs / 6572.2:.3f} of measured 6572 GB/s)", flush=True)


if __name__ == "__main__":
    main()

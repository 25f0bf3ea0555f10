import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol, parity_util as pu
pkg = importlib.import_module("cineform-sdk_b200")
w, h = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(1)
frame = pu.synthetic_yuyv(rng, w, h, "random")
desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
quant = pkg.quant_for_quality(desc, 4)
with pkg.Context(0) as ctx, pkg.Codec(ctx, desc, 1) as codec:
    got = codec.unpack_coded(codec.forward_host([frame], quant)[0])
want = pu.oracle_forward_422(ol.oracle(), frame, quant, 0)
for key in sorted(want):
    bad = np.argwhere(got[key] != want[key])
    if len(bad):
        rows = sorted(set(bad[:, 0].tolist())); cols = sorted(set(bad[:, 1].tolist()))
        print(key, want[key].shape, "n", len(bad), "rows", rows[:12], "cols", cols[:12], "...", cols[-3:])
print("done")

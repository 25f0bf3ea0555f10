"""GPU parity tests of the forward path (level 1 packed 4:2:2 + levels 2,3 + fused quantisation),
called through the C ABI (include/cfhd_b200.h) and compared bit for bit with the oracle and with
the golden vectors produced by the reference itself."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as ol
import parity_util as pu
from test_golden import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("cineform-sdk_b200")


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _compare(got, want):
    assert set(want) <= set(got)
    for key in sorted(want):
        if not np.array_equal(got[key], want[key]):
            bad = np.argwhere(got[key] != want[key])
            raise AssertionError(f"band {key}: {bad.shape[0]} mismatches, first at {bad[:5].tolist()} "
                                 f"got {got[key][tuple(bad[0])]} want {want[key][tuple(bad[0])]}")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_vectors(pkg, ctx, path):
    frame, div, prescale, quality, bands = load_golden(path)
    h, w2 = frame.shape
    desc = pkg.FrameDesc(w2 // 2, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, quality)
    assert quant.table(3) == div
    with pkg.Codec(ctx, desc, 1) as codec:
        got = codec.unpack_coded(codec.forward_host([frame], quant)[0])
    want = {k: v for k, v in bands.items() if not (k[2] == "LL" and k[1] != 3)}
    _compare(got, want)


@pytest.mark.parametrize("size", [(192, 48), (256, 64), (320, 56), (704, 96), (1920, 1080)])
@pytest.mark.parametrize("kind", ["natural", "random", "extreme", "constant"])
@pytest.mark.parametrize("fmt", [0, 1])
def test_forward_422_vs_oracle(pkg, ctx, size, kind, fmt):
    w, h = size
    if (w, h) == (1920, 1080) and kind not in ("natural", "random"):
        pytest.skip("large size covered by natural/random")
    rng = np.random.default_rng(w * 31 + h + fmt)
    frame = pu.synthetic_yuyv(rng, w, h, kind)
    if fmt == 1:
        frame = pu.yuyv_to_uyvy(frame)
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_UYVY if fmt else pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4)
    with pkg.Codec(ctx, desc, 1) as codec:
        got = codec.unpack_coded(codec.forward_host([frame], quant)[0])
    _compare(got, pu.oracle_forward_422(ol.oracle(), frame, quant, fmt))


@pytest.mark.parametrize("quality", [1, 2, 3, 5, 6])
def test_forward_422_qualities(pkg, ctx, quality):
    w, h = 512, 128
    rng = np.random.default_rng(quality)
    frame = pu.synthetic_yuyv(rng, w, h, "natural")
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, quality)
    with pkg.Codec(ctx, desc, 1) as codec:
        got = codec.unpack_coded(codec.forward_host([frame], quant)[0])
    _compare(got, pu.oracle_forward_422(ol.oracle(), frame, quant, 0))


def test_forward_batch_and_pitch(pkg, ctx):
    """A batch of different frames in one launch, with a padded host pitch."""
    w, h, n = 640, 96, 5
    rng = np.random.default_rng(99)
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4)
    padded = [np.zeros((h, w * 2 + 64), np.uint8) for _ in range(n)]
    frames = []
    for p in padded:
        f = pu.synthetic_yuyv(rng, w, h, "natural")
        p[:, :w * 2] = f
        frames.append(f)
    with pkg.Codec(ctx, desc, n) as codec:
        views = [p[:, :] for p in padded]
        coded = codec.forward_host(views, quant)
        for f, cbuf in zip(frames, coded):
            _compare(codec.unpack_coded(cbuf), pu.oracle_forward_422(ol.oracle(), f, quant, 0))


def test_forward_4k_full_size(pkg, ctx):
    """BASELINE config 3 size: one 3840x2160 frame against the oracle + a second identical submission
    must give identical bytes (determinism), and a constant frame must give all-zero highpass bands."""
    w, h = 3840, 2160
    rng = np.random.default_rng(4)
    frame = pu.synthetic_yuyv(rng, w, h, "natural")
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4)
    with pkg.Codec(ctx, desc, 2) as codec:
        a, b = codec.forward_host([frame, frame], quant)
        assert np.array_equal(a, b)
        _compare(codec.unpack_coded(a), pu.oracle_forward_422(ol.oracle(), frame, quant, 0))
        const = np.full((h, w * 2), 128, np.uint8)
        bands = codec.unpack_coded(codec.forward_host([const], quant)[0])
        for key, arr in bands.items():
            if key[2] != "LL":
                assert not arr.any(), key
            else:
                assert (arr == arr[0, 0]).all()


def test_invalid_arguments(pkg, ctx):
    desc = pkg.FrameDesc(256, 64, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4)
    with pkg.Codec(ctx, desc, 1) as codec:
        with pytest.raises(pkg.CfbError) as ei:
            codec.forward_host([np.zeros((64, 512), np.uint8)] * 2, quant)      # batch > max_batch
        assert ei.value.code == 1
    with pytest.raises(pkg.CfbError):
        pkg.Codec(ctx, pkg.FrameDesc(250, 64, pkg.PIXEL_YUYV), 1)

"""GPU parity tests of the inverse path (fused dequantisation + 3 inverse levels), through the C ABI."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as ol
import parity_util as pu
from test_golden import GOLDEN, load_golden, load_golden_decoder_side

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("cineform-sdk_b200")


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _planar16(codec, coded, quant, pkg, w, h):
    out = np.zeros((3 * h, w), np.int16)
    codec.inverse_host([coded], quant, pkg.PIXEL_PLANAR16, [out])
    return [out[0:h, :w], out[h:2 * h, :w // 2], out[2 * h:3 * h, :w // 2]]     # Y, V, U


def _check_planes(got, want):
    for c, (g, w_) in enumerate(zip(got, want)):
        if not np.array_equal(g, w_):
            bad = np.argwhere(g != w_)
            raise AssertionError(f"channel {c}: {bad.shape[0]} mismatches, first {bad[:5].tolist()} "
                                 f"got {g[tuple(bad[0])]} want {w_[tuple(bad[0])]}")


@pytest.mark.parametrize("size", [(192, 48), (256, 64), (320, 56), (704, 96), (1920, 1080)])
@pytest.mark.parametrize("kind", ["natural", "random"])
def test_inverse_planar16_vs_oracle(pkg, ctx, size, kind):
    """Quantised bands produced by the oracle's forward -> our inverse (dequant fused) == oracle inverse."""
    w, h = size
    rng = np.random.default_rng(w + h)
    frame = pu.synthetic_yuyv(rng, w, h, kind)
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4)
    orc = ol.oracle()
    coded_bands = pu.oracle_forward_422(orc, frame, quant, 0)
    want = pu.inverse_pyramid(orc, coded_bands, quant.table(3), tuple(quant.prescale))
    with pkg.Codec(ctx, desc, 1) as codec:
        got = _planar16(codec, codec.pack_coded(coded_bands), quant, pkg, w, h)
    _check_planes(got, want)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_inverse_golden_decoder_bands(pkg, ctx, path):
    """Bands exactly as the reference's decoder held them -> our inverse: the 16-bit planes equal the oracle's
    and the 8-bit YUYV output lies inside the reference decoder's dither envelope (and within 1 LSB of the
    frame the reference actually produced)."""
    frame, div, prescale, quality, _ = load_golden(path)
    bands, dec = load_golden_decoder_side(path)
    h, w2 = frame.shape
    w = w2 // 2
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    unit = pkg.make_quant(pu.UNIT_DIVISORS, prescale)
    orc = ol.oracle()
    want = pu.inverse_pyramid(orc, bands, pu.UNIT_DIVISORS, prescale)
    with pkg.Codec(ctx, desc, 1) as codec:
        coded = codec.pack_coded(bands)
        _check_planes(_planar16(codec, coded, unit, pkg, w, h), want)
        out = np.zeros((h, w2), np.uint8)
        codec.inverse_host([coded], unit, pkg.PIXEL_YUYV, [out])
    a, b = pu.yuyv_envelope(want)
    ok = (out == a) | (out == b)
    assert ok.all(), f"{(~ok).sum()} bytes outside the reference's dither envelope"
    assert np.abs(out.astype(int) - dec.astype(int)).max() <= 1


@pytest.mark.parametrize("fmt", [0, 1])
def test_roundtrip_psnr_and_uyvy(pkg, ctx, fmt):
    w, h = 1920, 1080
    rng = np.random.default_rng(12)
    frame = pu.synthetic_yuyv(rng, w, h, "natural")
    if fmt:
        frame = pu.yuyv_to_uyvy(frame)
    pf = pkg.PIXEL_UYVY if fmt else pkg.PIXEL_YUYV
    desc = pkg.FrameDesc(w, h, pf)
    quant = pkg.quant_for_quality(desc, 4)
    with pkg.Codec(ctx, desc, 1) as codec:
        coded = codec.forward_host([frame], quant)[0]
        out = np.zeros_like(frame)
        codec.inverse_host([coded], quant, pf, [out])
    yo = 1 if fmt else 0
    # the synthetic frame carries sigma=2 noise, which FILMSCAN1 does not preserve: ~47.5 dB luma here
    assert pu.psnr(out[:, yo::2], frame[:, yo::2]) > 45.0        # luma PSNR, as TestCFHD reports it
    assert pu.psnr(out, frame) > 44.0


def test_roundtrip_4k_batch(pkg, ctx):
    """BASELINE config 3 size, batch of 3: encode -> decode through host buffers, PSNR + determinism."""
    w, h, n = 3840, 2160, 3
    rng = np.random.default_rng(5)
    base = pu.synthetic_yuyv(rng, w, h, "natural")
    frames = [np.roll(base, 64 * i, axis=1).copy() for i in range(n)]
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4)
    with pkg.Codec(ctx, desc, n) as codec:
        coded = codec.forward_host(frames, quant)
        outs = [np.zeros_like(f) for f in frames]
        codec.inverse_host(coded, quant, pkg.PIXEL_YUYV, outs)
        outs2 = [np.zeros_like(f) for f in frames]
        codec.inverse_host(coded, quant, pkg.PIXEL_YUYV, outs2)
    for f, o, o2 in zip(frames, outs, outs2):
        assert np.array_equal(o, o2)
        assert pu.psnr(o[:, 0::2], f[:, 0::2]) > 45.0


def _reduced(codec, pkg, coded, quant, res, fmt):
    """Decode `coded` at a reduced resolution: returns (packed 8-bit frame, [Y, V, U] int16 lowpass planes)."""
    codec.set_decode_resolution(res)
    try:
        w, h = codec.decoded_size()
        out = np.zeros((h, w * 2), np.uint8)
        codec.inverse_host([coded], quant, fmt, [out])
        pl = np.zeros((3 * h, w), np.int16)
        codec.inverse_host([coded], quant, pkg.PIXEL_PLANAR16, [pl])
    finally:
        codec.set_decode_resolution(pkg.RESOLUTION_FULL)
    return out, [pl[0:h, :w], pl[h:2 * h, :w // 2], pl[2 * h:3 * h, :w // 2]]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_reduced_resolution_golden(pkg, ctx, path):
    """CFHD_DECODED_RESOLUTION_HALF / _QUARTER: lowpass images equal the reference decoder's own LL1 / LL2, the
    half-resolution frame equals what CFHD_DecodeSample returned byte for byte, the quarter-resolution frame equals
    the oracle's CopyQuarterRowToBuffer restatement."""
    z = np.load(path)
    frame, _, prescale, _, _ = load_golden(path)
    bands, _ = load_golden_decoder_side(path)
    h, w2 = frame.shape
    desc = pkg.FrameDesc(w2 // 2, h, pkg.PIXEL_YUYV)
    unit = pkg.make_quant(pu.UNIT_DIVISORS, prescale)
    with pkg.Codec(ctx, desc, 1) as codec:
        coded = codec.pack_coded(bands)
        for res, stop, name in ((pkg.RESOLUTION_HALF, 1, "half"), (pkg.RESOLUTION_QUARTER, 2, "quarter")):
            out, planes = _reduced(codec, pkg, coded, unit, res, pkg.PIXEL_YUYV)
            _check_planes(planes, [z[f"r_{c}_{stop}_LL"] for c in range(3)])
            assert np.array_equal(out, pu.lowpass_to_422(planes, unsigned_shift=(stop == 2)))
            if name == "half":
                assert np.array_equal(out, z["decoded_half_yuy2"])
        # and the codec still decodes at full resolution afterwards
        full = np.zeros((h, w2), np.uint8)
        codec.inverse_host([coded], unit, pkg.PIXEL_YUYV, [full])
        assert np.abs(full.astype(int) - z["decoded_yuy2"].astype(int)).max() <= 1


@pytest.mark.parametrize("size", [(192, 48), (448, 120), (1920, 1080), (3840, 2160)])
@pytest.mark.parametrize("fmt_name", ["YUYV", "UYVY"])
def test_reduced_resolution_vs_oracle(pkg, ctx, size, fmt_name):
    """Random (adversarial: negative and > 4095 lowpass values occur) coefficients: both shift rules and byte orders."""
    w, h = size
    fmt = getattr(pkg, "PIXEL_" + fmt_name)
    rng = np.random.default_rng(w * 3 + h)
    frame = pu.synthetic_yuyv(rng, w, h, "random")
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 3)
    orc = ol.oracle()
    coded_bands = pu.oracle_forward_422(orc, frame, quant, 0)
    # push the lowpass images out of the 8-bit range in places: scale LL3 of every channel
    for c in range(3):
        ll = coded_bands[(c, 3, "LL")].astype(np.int32)
        coded_bands[(c, 3, "LL")] = np.clip((ll - 8000) * 3, -32768, 32767).astype(np.int16)
    with pkg.Codec(ctx, desc, 1) as codec:
        coded = codec.pack_coded(coded_bands)
        for res, stop in ((pkg.RESOLUTION_HALF, 1), (pkg.RESOLUTION_QUARTER, 2)):
            want = pu.inverse_pyramid(orc, coded_bands, quant.table(3), tuple(quant.prescale), stop_level=stop)
            out, planes = _reduced(codec, pkg, coded, quant, res, fmt)
            _check_planes(planes, want)
            assert np.array_equal(out, pu.lowpass_to_422(want, unsigned_shift=(stop == 2), uyvy=(fmt_name == "UYVY")))
            assert out.min() == 0                    # negative lowpass values occur (clamped / wrapped by the two rules)

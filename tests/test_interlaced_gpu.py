"""GPU parity tests of the interlaced (field transform) level 1, through the C ABI:
forward == oracle == the reference's EncodeSample bands (golden), inverse == oracle and inside the reference
decoder's dither envelope."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as ol
import parity_util as pu
from test_golden import GOLDEN_FIELDS, load_golden, load_golden_decoder_side

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("cineform-sdk_b200")


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _assert_bands(got, want):
    assert set(got) == set(want)
    for key in sorted(want):
        if not np.array_equal(got[key], want[key]):
            bad = np.argwhere(got[key] != want[key])
            raise AssertionError(f"band {key}: {bad.shape[0]} mismatches, first {bad[:4].tolist()}, "
                                 f"got {got[key][tuple(bad[0])]} want {want[key][tuple(bad[0])]}")


def _planar16(codec, pkg, coded, quant, w, h):
    out = np.zeros((3 * h, w), np.int16)
    codec.inverse_host([coded], quant, pkg.PIXEL_PLANAR16, [out])
    return [out[0:h, :w], out[h:2 * h, :w // 2], out[2 * h:3 * h, :w // 2]]


@pytest.mark.parametrize("path", GOLDEN_FIELDS, ids=[os.path.basename(p) for p in GOLDEN_FIELDS])
def test_forward_reproduces_reference_encoder_bands(pkg, ctx, path):
    frame, div, prescale, quality, bands = load_golden(path)
    h, w2 = frame.shape
    desc = pkg.FrameDesc(w2 // 2, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, quality, interlaced=True)
    assert quant.table(3) == div
    with pkg.Codec(ctx, desc, 1) as codec:
        codec.set_interlaced(True)
        coded = np.zeros(codec.layout.coded_bytes, np.uint8)
        codec.forward_host([frame], quant, [coded])
        got = codec.unpack_coded(coded)
    _assert_bands(got, {k: v for k, v in bands.items() if not (k[2] == "LL" and k[1] != 3)})


@pytest.mark.parametrize("path", GOLDEN_FIELDS, ids=[os.path.basename(p) for p in GOLDEN_FIELDS])
def test_inverse_of_reference_decoder_bands(pkg, ctx, path):
    """Bands as the reference's decoder held them (HL put back into its coded, differenced form): 16-bit planes equal
    the oracle's, 8-bit output inside the decoder's dither envelope and within 1 LSB of the frame it produced."""
    frame, div, prescale, quality, _ = load_golden(path)
    bands, dec = load_golden_decoder_side(path)
    for c in range(3):
        hl = bands[(c, 1, "HL")].astype(np.int32)
        hl[:, 1:] -= hl[:, :-1].copy()
        bands[(c, 1, "HL")] = hl.astype(np.int16)
    h, w2 = frame.shape
    w = w2 // 2
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    unit = pkg.make_quant(pu.UNIT_DIVISORS, prescale)
    want = pu.inverse_pyramid(ol.oracle(), bands, pu.UNIT_DIVISORS, prescale, interlaced=True)
    with pkg.Codec(ctx, desc, 1) as codec:
        codec.set_interlaced(True)
        coded = codec.pack_coded(bands)
        got = _planar16(codec, pkg, coded, unit, w, h)
        out = np.zeros((h, w2), np.uint8)
        codec.inverse_host([coded], unit, pkg.PIXEL_YUYV, [out])
    for c in range(3):
        assert np.array_equal(got[c], want[c]), f"channel {c}"
    a, b = pu.yuyv_envelope(want)
    assert ((out == a) | (out == b)).all()
    assert np.abs(out.astype(int) - dec.astype(int)).max() <= 1


@pytest.mark.parametrize("size", [(192, 48), (256, 64), (448, 120), (704, 96), (1920, 1080), (3840, 2160)])
@pytest.mark.parametrize("kind,fmt_name", [("natural", "YUYV"), ("random", "YUYV"), ("natural", "UYVY")])
def test_field_transform_vs_oracle(pkg, ctx, size, kind, fmt_name):
    """Forward and inverse against the oracle on synthetic interlaced content, strips and borders of every width class;
    then the round trip through our own forward + inverse."""
    w, h = size
    fmt = getattr(pkg, "PIXEL_" + fmt_name)
    uyvy = fmt_name == "UYVY"
    rng = np.random.default_rng(w * 5 + h + (7 if uyvy else 0))
    frame = pu.synthetic_yuyv(rng, w, h, kind)
    frame[1::2] = np.roll(frame[1::2], 8, axis=1)           # the two fields differ
    if uyvy:
        frame = pu.yuyv_to_uyvy(frame)
    desc = pkg.FrameDesc(w, h, fmt)
    quant = pkg.quant_for_quality(desc, 4 if kind == "natural" else 2, interlaced=True)
    orc = ol.oracle()
    want_bands = pu.oracle_forward_422(orc, frame, quant, 1 if uyvy else 0, interlaced=True)
    want_planes = pu.inverse_pyramid(orc, want_bands, quant.table(3), tuple(quant.prescale), interlaced=True)
    with pkg.Codec(ctx, desc, 1) as codec:
        codec.set_interlaced(True)
        coded = np.zeros(codec.layout.coded_bytes, np.uint8)
        codec.forward_host([frame], quant, [coded])
        _assert_bands(codec.unpack_coded(coded), want_bands)
        got = _planar16(codec, pkg, coded, quant, w, h)
        for c in range(3):
            assert np.array_equal(got[c], want_planes[c]), f"inverse channel {c}"
        out = np.zeros_like(frame)
        codec.inverse_host([coded], quant, fmt, [out])
        a, b = pu.yuyv_envelope(want_planes, uyvy=uyvy)
        assert ((out == a) | (out == b)).all()
        if kind == "natural":
            yo = 1 if uyvy else 0
            assert pu.psnr(out[:, yo::2], frame[:, yo::2]) > 40.0
        # switching the flag off restores the progressive transform on the same codec
        codec.set_interlaced(False)
        q2 = pkg.quant_for_quality(desc, 4)
        codec.forward_host([frame], q2, [coded])
        _assert_bands(codec.unpack_coded(coded), pu.oracle_forward_422(orc, frame, q2, 1 if uyvy else 0))


def test_interlaced_batch_device_resident(pkg, ctx):
    """Batch of 4 different frames in one launch == 4 single-frame results (frame index plumbing of the carries)."""
    w, h = 704, 96
    rng = np.random.default_rng(99)
    frames = []
    for i in range(4):
        f = pu.synthetic_yuyv(rng, w, h, "natural")
        f[1::2] = np.roll(f[1::2], 4 + 2 * i, axis=1)
        frames.append(f)
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4, interlaced=True)
    with pkg.Codec(ctx, desc, 4) as codec:
        codec.set_interlaced(True)
        coded = [np.zeros(codec.layout.coded_bytes, np.uint8) for _ in range(4)]
        codec.forward_host(frames, quant, coded)
        outs = [np.zeros_like(frames[0]) for _ in range(4)]
        codec.inverse_host(coded, quant, pkg.PIXEL_YUYV, outs)
        for i in range(4):
            one = np.zeros(codec.layout.coded_bytes, np.uint8)
            codec.forward_host([frames[i]], quant, [one])
            assert np.array_equal(one, coded[i])
            o1 = np.zeros_like(frames[0])
            codec.inverse_host([one], quant, pkg.PIXEL_YUYV, [o1])
            assert np.array_equal(o1, outs[i])


def test_interlaced_rejected_for_non_422(pkg, ctx):
    with pkg.Codec(ctx, pkg.FrameDesc(256, 64, pkg.PIXEL_RG48), 1) as codec:
        with pytest.raises(pkg.CfbError):
            codec.set_interlaced(True)

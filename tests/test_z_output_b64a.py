"""B64A output of the final inverse level for RGB 4:4:4 codecs ("decode to RG48 / B64A") on the GPU.
The rule (parity_util.pack_b64a) is pinned to the reference's decoder in test_output16.py; here the CUDA path is compared
with the oracle and, where oracle/_ref travelled, with the reference decoder's own frame."""
import importlib

import numpy as np
import pytest

import oracle_lib as ol
import parity_util as pu

needs_ref = pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref not built (reference absent)")
DECODED_FORMAT_B64A = 30


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("cineform-sdk_b200")


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(256, 64), (328, 48), (200, 48), (640, 96), (720, 480), (1920, 1080)])
@pytest.mark.parametrize("kind", ["natural", "extreme"])
def test_gpu_b64a_output_vs_oracle(pkg, size, kind):
    w, h = size
    rng = np.random.default_rng(w + h)
    frame = pu.synthetic_rg48(rng, w, h, kind)
    desc = pkg.FrameDesc(w, h, pkg.PIXEL_RG48)
    quant = pkg.quant_for_quality(desc, 4)
    orc = ol.oracle()
    pyr = pu.forward_pyramid_planes(orc, pu.unpack_rg48(frame), quant.table(3), tuple(quant.prescale))
    coded_bands = {k: v for k, v in pyr.items() if not (k[2] == "LL" and k[1] != 3)}
    planes = pu.inverse_pyramid(orc, coded_bands, quant.table(3), tuple(quant.prescale))
    want = pu.pack_b64a(planes)
    with pkg.Context(0) as ctx, pkg.Codec(ctx, desc, 2) as codec:
        coded = codec.pack_coded(coded_bands)
        outs = [np.zeros((h, 4 * w), np.uint16) for _ in range(2)]
        codec.inverse_host([coded, coded], quant, pkg.PIXEL_B64A, outs)
        assert np.array_equal(outs[0], want), np.argwhere(outs[0] != want)[:5].tolist()
        assert np.array_equal(outs[1], want)
        # the RG48 output of the same codec afterwards: the two stagings do not disturb each other
        rg = np.zeros((h, 3 * w), np.uint16)
        codec.inverse_host([coded], quant, pkg.PIXEL_RG48, [rg])
        assert np.array_equal(rg, pu.pack_rg48(planes))
        # a padded output pitch
        wide = np.zeros((h, 4 * w + 8), np.uint16)
        codec.inverse_host([coded], quant, pkg.PIXEL_B64A, [wide])
        assert np.array_equal(wide[:, :4 * w], want) and not wide[:, 4 * w:].any()


@pytest.mark.gpu
def test_gpu_b64a_needs_a_444_codec(pkg):
    desc = pkg.FrameDesc(256, 64, pkg.PIXEL_YUYV)
    quant = pkg.quant_for_quality(desc, 4)
    with pkg.Context(0) as ctx, pkg.Codec(ctx, desc, 1) as codec:
        coded = np.zeros(codec.layout.coded_bytes, np.uint8)
        with pytest.raises(pkg.CfbError):
            codec.inverse_host([coded], quant, pkg.PIXEL_B64A, [np.zeros((64, 4 * 256), np.uint16)])


@needs_ref
@pytest.mark.gpu
def test_gpu_b64a_vs_reference_decoder(pkg):
    w, h = 640, 96
    ref_lib = ol.load_ref()
    frame = pu.qbist_rg48(ref_lib, w, h, 1)
    _, _, prescale, sample = pu.ref_encode_frame(ref_lib, frame.view(np.uint8), w, h, pu.COLOR_FORMAT_RG48, 1, 3, 4)
    ref_out, bands = pu.ref_decode_sample_raw(ref_lib, sample, w, h, DECODED_FORMAT_B64A, 3, w * 8)
    bands = {k: v for k, v in bands.items() if not (k[2] == "LL" and k[1] != 3)}
    unit = pkg.make_quant(pu.UNIT_DIVISORS, prescale[0])
    with pkg.Context(0) as ctx, pkg.Codec(ctx, pkg.FrameDesc(w, h, pkg.PIXEL_RG48), 1) as codec:
        out = np.zeros((h, 4 * w), np.uint16)
        codec.inverse_host([codec.pack_coded(bands)], unit, pkg.PIXEL_B64A, [out])
    assert np.array_equal(out.view(np.uint8).reshape(h, -1), ref_out)

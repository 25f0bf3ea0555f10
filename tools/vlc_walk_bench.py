#!/usr/bin/env python
"""Host run-length / VLC coder: time per frame on one core (no GPU needed).

One Qbist frame through the reference's encoder (oracle/_ref), its coded region compacted to the sparse transfer format,
then every coded band through cfb_sparse_vlc_band / cfb_dense_vlc_band / the reference's EncodeQuantLongRuns (probe)."""
import argparse
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle_lib as ol      # noqa: E402
import parity_util as pu     # noqa: E402
from test_vlc import all_bands, decode_book, ref_decode_band, ref_encode_band, ref_encode_band_finished, ref_tables     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--codebook", type=int, default=0, help="code set of the reference encoder (Codec/codebooks.c)")
    ap.add_argument("--frame", type=int, default=3, help="Qbist frame number")
    a = ap.parse_args()
    pkg = importlib.import_module("cineform-sdk_b200")
    ref = ol.load_ref()
    w, h = a.width, a.height
    frame = pu.qbist_yuy2(ref, w, h, a.frame)
    bands, _, _, sample = pu.ref_encode_frame(ref, frame, w, h, pu.COLOR_FORMAT_YUYV, 0, 3, 4)
    lay = pkg.layout_for(pkg.FrameDesc(w, h, pkg.PIXEL_YUYV))
    coded = pkg.pack_coded(lay, bands)
    sparse = pkg.sparse_compact(lay, coded)
    book = pkg.VlcCodebook.from_arrays(*ref_tables(ref, a.codebook))
    views = []
    for c, k, b in all_bands(lay):
        bl = lay.band[c][k][b]
        views.append((c, k, b, bl, coded[bl.offset: bl.offset + bl.pitch * bl.height].view(np.int16).reshape(bl.height, bl.pitch // 2)))
    nz = sum(int(np.count_nonzero(v[4][:, :v[3].width])) for v in views)
    cap = lay.coded_bytes

    def best(fn):
        ts = []
        for _ in range(a.reps):
            t = time.perf_counter()
            n = fn()
            ts.append(time.perf_counter() - t)
        return min(ts) * 1e3, n

    # the C entry points with one preallocated stream buffer (the Python wrappers allocate per call)
    lib, out = pkg.lib(), np.zeros(cap, np.uint8)
    if os.environ.get("CFB_VLC_LIB"):     # a standalone build of cfb_vlc.cu (development)
        alt = C.CDLL(os.environ["CFB_VLC_LIB"])
        for name in ("cfb_sparse_vlc_band", "cfb_dense_vlc_band", "cfb_sparse_band_nonzeros"):
            getattr(alt, name).argtypes = getattr(lib, name).argtypes
            getattr(alt, name).restype = getattr(lib, name).restype
        lib = alt

    def writer():
        bw = pkg.BitWriter()
        bw.cur, bw.end, bw.buffer, bw.bits_free, bw.bytes = out.ctypes.data, out.ctypes.data + out.size, 0, 32, 0
        return bw

    def run_sparse():
        n = 0
        for c, k, b, _, _ in views:
            bw = writer()
            assert lib.cfb_sparse_vlc_band(C.byref(lay), sparse.ctypes.data, c, k, b, C.byref(book), C.byref(bw)) == 0
            n += bw.bytes
        return n

    def run_dense():
        n = 0
        for _, _, _, bl, p in views:
            bw = writer()
            assert lib.cfb_dense_vlc_band(p.ctypes.data, bl.width, bl.height, bl.pitch, C.byref(book), C.byref(bw)) == 0
            n += bw.bytes
        return n

    t_sparse, n1 = best(run_sparse)
    def run_count():
        n, one = 0, C.c_uint32()
        for c, k, b, _, _ in views:
            assert lib.cfb_sparse_band_nonzeros(C.byref(lay), sparse.ctypes.data, c, k, b, C.byref(one)) == 0
            n += one.value
        return n

    t_count, n0 = best(run_count)
    assert n0 == nz
    t_dense, n2 = best(run_dense)
    t_ref, n3 = best(lambda: sum(ref_encode_band(ref, p, bl.width, a.codebook, 0)[0].size for _, _, _, bl, p in views))
    assert n1 == n2 == n3
    print(f"{w}x{h} Qbist frame: {nz} non-zero coefficients, {n1} bytes of band streams, sparse buffer {sparse.size} bytes, sample {sample.size} bytes")
    print(f"  cfb_sparse_vlc_band    {t_sparse:7.2f} ms per frame ({t_sparse * 1e6 / nz:5.2f} ns per coefficient); walk only {t_count:.2f} ms")
    print(f"  cfb_dense_vlc_band     {t_dense:7.2f} ms")
    print(f"  reference coder        {t_ref:7.2f} ms (EncodeQuantLongRuns band by band through the probe)")

    # decoder side: the band streams (with end-of-band code) parsed straight into the sparse format vs the reference's FSM
    # decoder writing dense bands
    streams = [(c, k, b, bl, ref_encode_band_finished(ref, p, bl.width, a.codebook)) for c, k, b, bl, p in views if b != 0]
    ll = [(c, k, b, bl, p) for c, k, b, bl, p in views if b == 0]
    dec = pkg.VlcDecoder(lay, decode_book(pkg, ref, a.codebook))

    sparse_out = np.zeros(pkg.sparse_max_bytes(lay), np.uint8)          # one buffer for the frame loop, as a decoder would keep

    def run_decode():
        dec.begin(sparse_out)
        it = iter(streams)
        for c in range(lay.num_channels):
            for item in ll:
                if item[0] == c:
                    dec.dense_band(item[0], item[1], item[2], item[4][:, :item[3].width])
            for _ in range(9):
                cc, k, b, bl, st = next(it)
                dec.band(cc, k, b, st, 1)
        return dec.end().size

    t_dec, nsp = best(run_decode)
    t_fsm, _ = best(lambda: sum(ref_decode_band(ref, st, bl.width, bl.height, bl.pitch, a.codebook, 1).size for _, _, _, bl, st in streams))
    print(f"  cfb_vlc_decode_band    {t_dec:7.2f} ms per frame, band streams -> sparse buffer ({nsp} bytes)")
    print(f"  reference FSM decoder  {t_fsm:7.2f} ms (DecodeBandFSM16sNoGap band by band through the probe, dense bands)")
    dec.close()


if __name__ == "__main__":
    main()

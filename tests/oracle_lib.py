"""ctypes loaders for the CHECKERS (test infrastructure only):

  * oracle/liboracle.so          -- our C restatement of the reference algorithm
  * oracle/_ref/libcfhd_ref.so   -- the unmodified reference + ref_probe.cpp shim

Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _declare(lib, prefix):
    g = lambda n: getattr(lib, prefix + n)
    g("fwd_row").argtypes = [_i16p, _i16p, _i16p, C.c_int, C.c_int]
    g("fwd_row").restype = None
    g("quantize_row").argtypes = [_i16p, _i16p, C.c_int, C.c_int, C.c_int]
    g("quantize_row").restype = None
    g("fwd_level").argtypes = [_i16p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int,
                               _i16p, _i16p, _i16p, _i16p, C.c_int]
    g("fwd_level").restype = None
    g("fwd_level_422").argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int,
                                   _i16p, _i16p, _i16p, _i16p, C.c_int]
    g("fwd_level_422").restype = None
    g("inv_level").argtypes = [_i16p, _i16p, _i16p, _i16p, C.c_int, C.c_int, C.c_int, C.c_int, _i16p, C.c_int]
    g("inv_level").restype = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def load_oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    _declare(lib, "orc_")
    return lib


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libcfhd_ref.so"))


def load_ref():
    lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libcfhd_ref.so"), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
    _declare(lib, "ref_")
    lib.ref_qbist_frames.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, _u8p]
    lib.ref_qbist_frames.restype = None
    return lib


class Impl:
    """Uniform numpy-level wrapper over either checker library."""

    def __init__(self, lib, prefix):
        self.lib, self.p = lib, prefix

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    def fwd_row(self, x, prescale):
        x = np.ascontiguousarray(x, dtype=np.int16)
        m = x.size // 2
        lo, hi = np.zeros(m, np.int16), np.zeros(m, np.int16)
        self._f("fwd_row")(x, lo, hi, x.size, prescale)
        return lo, hi

    def quantize_row(self, x, divisor, midpoint=2):
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.zeros_like(x)
        self._f("quantize_row")(x, out, x.size, divisor, midpoint)
        return out

    def fwd_level(self, plane, variant, quant, midpoint=2):
        plane = np.ascontiguousarray(plane, dtype=np.int16)
        h, w = plane.shape
        bands = [np.zeros((h // 2, w // 2), np.int16) for _ in range(4)]
        q = np.asarray(quant, dtype=np.int32)
        self._f("fwd_level")(plane, w * 2, w, h, variant, q, midpoint, *bands, (w // 2) * 2)
        return bands

    def fwd_level_422(self, frame, channel, fmt, quant, precision=10, midpoint=2):
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        h, pitch = frame.shape
        fw = pitch // 2
        w = fw if channel == 0 else fw // 2
        bands = [np.zeros((h // 2, w // 2), np.int16) for _ in range(4)]
        q = np.asarray(quant, dtype=np.int32)
        self._f("fwd_level_422")(frame, pitch, w, h, channel, fmt, precision, q, midpoint, *bands, (w // 2) * 2)
        return bands

    def fwd_fields_422(self, frame, channel, fmt, quant, precision=10, midpoint=2):
        """Interlaced (field) transform of level 1, oracle only (orc_fwd_fields_422)."""
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        h, pitch = frame.shape
        fw = pitch // 2
        w = fw if channel == 0 else fw // 2
        bands = [np.zeros((h // 2, w // 2), np.int16) for _ in range(4)]
        q = np.asarray(quant, dtype=np.int32)
        f = self._f("fwd_fields_422")
        f.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int,
                      _i16p, _i16p, _i16p, _i16p, C.c_int]
        f.restype = None
        f(frame, pitch, w, h, channel, fmt, precision, q, midpoint, *bands, (w // 2) * 2)
        return bands

    def fwd_fields_plane(self, plane, quant, midpoint=2):
        """Planar interlaced (field) transform of level 1, oracle only (orc_fwd_fields_plane)."""
        plane = np.ascontiguousarray(plane, dtype=np.int16)
        h, w = plane.shape
        bands = [np.zeros((h // 2, w // 2), np.int16) for _ in range(4)]
        q = np.asarray(quant, dtype=np.int32)
        f = self._f("fwd_fields_plane")
        f.argtypes = [_i16p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i16p, _i16p, _i16p, _i16p, C.c_int]
        f.restype = None
        f(plane, w * 2, w, h, q, midpoint, *bands, (w // 2) * 2)
        return bands

    def inv_fields(self, ll, lh, hl, hh):
        """Inverse field transform on decoder-side (dequantised, HL integrated) bands (orc_inv_fields)."""
        bands = [np.ascontiguousarray(b, dtype=np.int16) for b in (ll, lh, hl, hh)]
        h, w = bands[0].shape
        out = np.zeros((2 * h, 2 * w), np.int16)
        f = self._f("inv_fields")
        f.argtypes = [_i16p, _i16p, _i16p, _i16p, C.c_int, C.c_int, C.c_int, _i16p, C.c_int]
        f.restype = None
        f(*bands, w * 2, w, h, out, w * 4)
        return out

    def inv_level(self, ll, lh, hl, hh, descale):
        bands = [np.ascontiguousarray(b, dtype=np.int16) for b in (ll, lh, hl, hh)]
        h, w = bands[0].shape
        out = np.zeros((2 * h, 2 * w), np.int16)
        self._f("inv_level")(*bands, w * 2, w, h, descale, out, w * 4)
        return out


def oracle():
    return Impl(load_oracle(), "orc_")


def ref():
    return Impl(load_ref(), "ref_")


CFHD_PIXEL_FORMAT_YUY2 = (ord("Y") << 24) | (ord("U") << 16) | (ord("Y") << 8) | ord("2")  # FOUR_CHAR_CODE, Common/CFHDTypes.h:122


def fourcc(s):
    return (ord(s[0]) << 24) | (ord(s[1]) << 16) | (ord(s[2]) << 8) | ord(s[3])

"""ctypes access to oracle/_ref/libkat.so: frame-level drivers around the UNMODIFIED reference's
loop filter and predictors (oracle/refbuild/katharness.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libkat.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.kat_loop_filter_frame.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kat_inter_frame.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.kat_intra_frame.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _rec(fr):
    from vp8b200.recfile import HDR_DTYPE
    hdr = np.asarray(fr.hdr, HDR_DTYPE).tobytes()
    mb = np.ascontiguousarray(fr.mb)
    aux = np.ascontiguousarray(fr.aux) if fr.aux.size else np.zeros((1, 64), np.uint8)
    return hdr, mb, aux


def loop_filter_frame(w, h, fr, frame):
    """vp8_loop_filter_frame (loopfilter.c:203) in place on `frame` (whole YV12 allocation)."""
    hdr, mb, aux = _rec(fr)
    assert frame.flags["C_CONTIGUOUS"] and frame.dtype == np.uint8
    assert lib().kat_loop_filter_frame(w, h, hdr, mb.ctypes.data, aux.ctypes.data, frame.ctypes.data) == 0
    return frame


def inter_frame(w, h, fr, fbs):
    """vp8_build_inter_predictors_mb (reconinter.c:560) for every inter MB; fbs = [dst, last, golden, altref]."""
    hdr, mb, aux = _rec(fr)
    arr = (C.c_void_p * 4)(*[f.ctypes.data for f in fbs])
    assert lib().kat_inter_frame(w, h, hdr, mb.ctypes.data, aux.ctypes.data, arr) == 0
    return fbs[0]


def intra_frame(w, h, fr, frame):
    """the reference's intra predictors for every MB in raster order, no residual."""
    hdr, mb, aux = _rec(fr)
    assert lib().kat_intra_frame(w, h, mb.ctypes.data, aux.ctypes.data, frame.ctypes.data) == 0
    return frame

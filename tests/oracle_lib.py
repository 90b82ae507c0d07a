"""ctypes access to the CPU oracle (oracle/recon_oracle.c) - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200"))
from vp8b200.recfile import HDR_DTYPE  # noqa: E402

SRC = os.path.join(ROOT, "oracle", "recon_oracle.c")
LIB = os.path.join(ROOT, "oracle", "_build", "librecon_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
                               "-o", LIB, SRC])
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_frame_size.restype = C.c_size_t
        L.oracle_frame_size.argtypes = [C.c_void_p]
        L.oracle_fb.restype = C.POINTER(C.c_uint8)
        L.oracle_fb.argtypes = [C.c_void_p, C.c_int]
        L.oracle_frame.argtypes = [C.c_void_p] * 5
        L.oracle_frame_stages.argtypes = [C.c_void_p] * 5 + [C.c_int]
        L.oracle_idct_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_dc_add.argtypes = [C.c_int16, C.c_void_p, C.c_int]
        L.oracle_iwalsh.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_sixtap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_bilinear.argtypes = L.oracle_sixtap.argtypes
        L.oracle_intra4x4.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_extend_borders.argtypes = [C.c_void_p, C.c_int]
        L.oracle_loop_filter.argtypes = [C.c_void_p] * 3
        _lib = L
    return _lib


class OracleDecoder:
    """CPU restatement of the reconstruction path, same records in, same buffers out."""

    def __init__(self, coded_w, coded_h, n_fb=4):
        self.L = lib()
        self.h = self.L.oracle_create(coded_w, coded_h, n_fb)
        assert self.h
        self.frame_size = self.L.oracle_frame_size(self.h)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def fb(self, i):
        """Writable numpy view of frame buffer i (whole allocation)."""
        return np.ctypeslib.as_array(self.L.oracle_fb(self.h, i), shape=(self.frame_size,))

    def frame(self, fr, stages=7):
        hdr = np.asarray(fr.hdr, HDR_DTYPE).tobytes()
        mb = np.ascontiguousarray(fr.mb)
        aux = np.ascontiguousarray(fr.aux)
        coef = np.ascontiguousarray(fr.coef)
        self.L.oracle_frame_stages(self.h, hdr, mb.ctypes.data, aux.ctypes.data if aux.size else None,
                                   coef.ctypes.data if coef.size else None, stages)

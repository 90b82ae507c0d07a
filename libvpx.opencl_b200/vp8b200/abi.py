"""ctypes binding of include/vp8b200.h.  There is no fallback: importing the library fails
loudly when libvp8b200.so has not been built, and Context() raises when no CUDA device is
usable."""
import ctypes as C
import os

import numpy as np

from .recfile import HDR_DTYPE, MB_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
# VP8B200_LIB selects another build of the same library (A/B experiments); default = in-tree build
LIB_PATH = os.environ.get("VP8B200_LIB") or os.path.join(os.path.dirname(_HERE), "libvp8b200.so")

EXPORTS = [
    "vp8b200_abi_version", "vp8b200_strerror", "vp8b200_last_error", "vp8b200_device_count",
    "vp8b200_create", "vp8b200_destroy", "vp8b200_frame_size", "vp8b200_y_stride",
    "vp8b200_host_alloc", "vp8b200_host_alloc_on", "vp8b200_host_free", "vp8b200_frame_fetch_begin",
    "vp8b200_frame_fetch_wait", "vp8b200_frame_begin", "vp8b200_frame_submit",
    "vp8b200_frame_abort", "vp8b200_frame_submit_show", "vp8b200_engine_stats", "vp8b200_frame_fetch", "vp8b200_frame_upload", "vp8b200_frame_copy",
    "vp8b200_sync", "vp8b200_stage_frame", "vp8b200_staged_free", "vp8b200_batch_run",
    "vp8b200_launch_count", "vp8b200_stream", "vp8b200_global_stats", "vp8b200_profile_enable",
    "vp8b200_profile_read",
]


class FrameBufs(C.Structure):
    _fields_ = [("mb", C.c_void_p), ("aux", C.c_void_p), ("coef", C.c_void_p),
                ("aux_capacity", C.c_uint32), ("coef_capacity", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libvp8b200.so is not built (%s); run __graft_entry__.build()" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.vp8b200_strerror.restype = C.c_char_p
        L.vp8b200_last_error.restype = C.c_char_p
        L.vp8b200_last_error.argtypes = [C.c_void_p]
        L.vp8b200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
        L.vp8b200_destroy.argtypes = [C.c_void_p]
        L.vp8b200_frame_size.restype = C.c_size_t
        L.vp8b200_frame_size.argtypes = [C.c_void_p]
        L.vp8b200_y_stride.argtypes = [C.c_void_p]
        L.vp8b200_host_alloc.restype = C.c_void_p
        L.vp8b200_host_alloc.argtypes = [C.c_size_t]
        L.vp8b200_host_free.argtypes = [C.c_void_p]
        L.vp8b200_frame_begin.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(FrameBufs)]
        L.vp8b200_frame_submit.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.vp8b200_frame_abort.argtypes = [C.c_void_p]
        L.vp8b200_frame_submit_show.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.vp8b200_engine_stats.restype = None
        L.vp8b200_engine_stats.argtypes = [C.c_int, C.POINTER(C.c_uint64)]
        L.vp8b200_frame_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.vp8b200_frame_fetch_begin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.vp8b200_frame_fetch_wait.argtypes = [C.c_void_p]
        L.vp8b200_host_alloc_on.restype = C.c_void_p
        L.vp8b200_host_alloc_on.argtypes = [C.c_int, C.c_size_t]
        L.vp8b200_frame_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.vp8b200_frame_copy.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.vp8b200_sync.argtypes = [C.c_void_p]
        L.vp8b200_stage_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                          C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.vp8b200_staged_free.argtypes = [C.c_void_p, C.c_void_p]
        L.vp8b200_batch_run.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
        L.vp8b200_launch_count.restype = C.c_uint64
        L.vp8b200_launch_count.argtypes = [C.c_void_p]
        L.vp8b200_stream.restype = C.c_void_p
        L.vp8b200_stream.argtypes = [C.c_void_p]
        L.vp8b200_global_stats.restype = None
        L.vp8b200_global_stats.argtypes = [C.POINTER(C.c_uint64)]
        L.vp8b200_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.vp8b200_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


class Vp8b200Error(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class Context:
    """One decoder instance's device state (vp8b200_ctx)."""

    def __init__(self, coded_w, coded_h, n_fb=4, device=0):
        self.L = lib()
        h = C.c_void_p()
        st = self.L.vp8b200_create(C.byref(h), device, coded_w, coded_h, n_fb)
        if st:
            raise Vp8b200Error("vp8b200_create: " + self.L.vp8b200_strerror(st).decode())
        self.h = h
        self.frame_size = self.L.vp8b200_frame_size(h)
        self.n_mb = (coded_w // 16) * (coded_h // 16)
        self._staged = []

    def _ck(self, st, what):
        if st:
            raise Vp8b200Error("%s: %s (%s)" % (what, self.L.vp8b200_strerror(st).decode(),
                                                self.L.vp8b200_last_error(self.h).decode()))

    def close(self):
        if self.h:
            for s in self._staged:
                self.L.vp8b200_staged_free(self.h, s)
            self._staged = []
            self.L.vp8b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, fr):
        """frame_begin + fill the pinned buffers + frame_submit (host buffers -> device)."""
        hdr = np.asarray(fr.hdr, HDR_DTYPE).tobytes()
        bufs = FrameBufs()
        self._ck(self.L.vp8b200_frame_begin(self.h, hdr, C.byref(bufs)), "frame_begin")
        C.memmove(bufs.mb, fr.mb.ctypes.data, 16 * self.n_mb)
        if fr.n_aux:
            C.memmove(bufs.aux, fr.aux.ctypes.data, 64 * fr.n_aux)
        if fr.n_coef:
            C.memmove(bufs.coef, fr.coef.ctypes.data, 32 * fr.n_coef)
        self._ck(self.L.vp8b200_frame_submit(self.h, fr.n_aux, fr.n_coef), "frame_submit")

    def submit_show(self, fr, show_fb=-1, out=None, display=(0, 0)):
        """frame_begin + fill + frame_submit_show: the coalesced path (engine thread issues)."""
        hdr = np.asarray(fr.hdr, HDR_DTYPE).tobytes()
        bufs = FrameBufs()
        self._ck(self.L.vp8b200_frame_begin(self.h, hdr, C.byref(bufs)), "frame_begin")
        C.memmove(bufs.mb, fr.mb.ctypes.data, 16 * self.n_mb)
        if fr.n_aux:
            C.memmove(bufs.aux, fr.aux.ctypes.data, 64 * fr.n_aux)
        if fr.n_coef:
            C.memmove(bufs.coef, fr.coef.ctypes.data, 32 * fr.n_coef)
        dst = out.ctypes.data_as(C.c_void_p) if out is not None else None
        self._ck(self.L.vp8b200_frame_submit_show(self.h, fr.n_aux, fr.n_coef, show_fb, dst, display[0], display[1]),
                 "frame_submit_show")

    def fetch_wait(self):
        self._ck(self.L.vp8b200_frame_fetch_wait(self.h), "frame_fetch_wait")

    def fetch(self, fb, out=None):
        if out is None:
            out = np.empty(self.frame_size, np.uint8)
        self._ck(self.L.vp8b200_frame_fetch(self.h, fb, out.ctypes.data_as(C.c_void_p), out.size), "frame_fetch")
        return out

    def fetch_visible(self, fb, display_w, display_h, out=None):
        """Lazy fetch of the visible samples only; the rest of `out` is left untouched."""
        if out is None:
            out = np.zeros(self.frame_size, np.uint8)
        self._ck(self.L.vp8b200_frame_fetch_begin(self.h, fb, out.ctypes.data_as(C.c_void_p), display_w, display_h),
                 "frame_fetch_begin")
        self._ck(self.L.vp8b200_frame_fetch_wait(self.h), "frame_fetch_wait")
        return out

    def upload(self, fb, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        self._ck(self.L.vp8b200_frame_upload(self.h, fb, buf.ctypes.data_as(C.c_void_p), buf.size), "frame_upload")

    def sync(self):
        self._ck(self.L.vp8b200_sync(self.h), "sync")

    def stage(self, fr):
        s = C.c_void_p()
        hdr = np.asarray(fr.hdr, HDR_DTYPE).tobytes()
        mb = np.ascontiguousarray(fr.mb)
        aux = np.ascontiguousarray(fr.aux)
        coef = np.ascontiguousarray(fr.coef)
        self._ck(self.L.vp8b200_stage_frame(self.h, hdr, _ptr(mb), _ptr(aux), fr.n_aux, _ptr(coef),
                                            fr.n_coef, C.byref(s)), "stage_frame")
        self._staged.append(s)
        return s

    def profile(self, enable=True):
        self._ck(self.L.vp8b200_profile_enable(self.h, int(enable)), "profile_enable")

    def profile_read(self):
        """{kernel: (total ms, launches)} since the last read; waits for the stream."""
        ms = (C.c_double * 4)()
        n = (C.c_uint64 * 4)()
        self._ck(self.L.vp8b200_profile_read(self.h, ms, n), "profile_read")
        return {k: (ms[i], int(n[i])) for i, k in enumerate(("inter", "intra", "loopfilter", "border"))}

    def launch_count(self):
        return int(self.L.vp8b200_launch_count(self.h))

    def stream(self):
        return self.L.vp8b200_stream(self.h)


def global_stats():
    out = (C.c_uint64 * 4)()
    lib().vp8b200_global_stats(out)
    return {"h2d_bytes": int(out[0]), "d2h_bytes": int(out[1]), "launches": int(out[2]), "frames": int(out[3])}


def batch_run(ctxs, staged):
    """One launch of each kernel covering frame staged[i] of context ctxs[i]."""
    n = len(ctxs)
    ca = (C.c_void_p * n)(*[c.h for c in ctxs])
    sa = (C.c_void_p * n)(*staged)
    st = lib().vp8b200_batch_run(ca, sa, n)
    if st:
        ctxs[0]._ck(st, "batch_run")

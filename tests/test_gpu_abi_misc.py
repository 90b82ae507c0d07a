"""GPU: the remaining C-ABI entry points and the error behaviour of the boundary."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT, load_case
from oracle_lib import OracleDecoder
from vp8b200 import abi, frames, recfile
import randrec

pytestmark = pytest.mark.gpu


def test_upload_fetch_copy_roundtrip(gpu_lib):
    geo = frames.Geometry(64, 48)
    ctx = abi.Context(64, 48, 4)
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, geo.frame_size, dtype=np.uint8)
    ctx.upload(2, a)
    assert np.array_equal(ctx.fetch(2), a)
    assert gpu_lib.vp8b200_frame_copy(ctx.h, 0, 2) == 0          # onyxd_if.c:390 path
    assert np.array_equal(ctx.fetch(0), a)
    assert gpu_lib.vp8b200_frame_copy(ctx.h, 9, 0) == -1         # bad index
    assert gpu_lib.vp8b200_frame_fetch(ctx.h, 0, None, 16) == -1
    ctx.close()


def test_lazy_fetch_of_the_visible_samples(gpu_lib):
    """N2/N3: fetch_begin queues the copy, fetch_wait delivers it; with a display size only the
    visible luma / chroma samples move and every other byte of the host image stays as it was."""
    geo = frames.Geometry(64, 48)
    ctx = abi.Context(64, 48, 4)
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, geo.frame_size, dtype=np.uint8)
    ctx.upload(1, a)
    dw, dh = 61, 43                                               # odd display size: chroma is (w+1)/2 x (h+1)/2
    out = ctx.fetch_visible(1, dw, dh, np.full(geo.frame_size, 0xEE, np.uint8))
    assert np.array_equal(geo.i420(out, dw, dh), geo.i420(a, dw, dh))
    want = np.full(geo.frame_size, 0xEE, np.uint8)
    for off, stride, w, h in ((geo.y_off, geo.y_stride, dw, dh), (geo.u_off, geo.uv_stride, (dw + 1) // 2, (dh + 1) // 2),
                              (geo.v_off, geo.uv_stride, (dw + 1) // 2, (dh + 1) // 2)):
        for y in range(h):
            want[off + y * stride: off + y * stride + w] = a[off + y * stride: off + y * stride + w]
    assert np.array_equal(out, want)
    # whole allocation through the same lazy pair; two copies may be in flight, fetch_wait
    # collects the oldest; a third begin first waits for the oldest itself
    buf1, buf2, buf3 = (np.zeros(geo.frame_size, np.uint8) for _ in range(3))
    assert gpu_lib.vp8b200_frame_fetch_begin(ctx.h, 1, buf1.ctypes.data_as(C.c_void_p), 0, 0) == 0
    assert gpu_lib.vp8b200_frame_fetch_begin(ctx.h, 1, buf2.ctypes.data_as(C.c_void_p), 0, 0) == 0
    assert gpu_lib.vp8b200_frame_fetch_wait(ctx.h) == 0
    assert np.array_equal(buf1, a)
    assert gpu_lib.vp8b200_frame_fetch_begin(ctx.h, 1, buf3.ctypes.data_as(C.c_void_p), 0, 0) == 0
    assert gpu_lib.vp8b200_frame_fetch_begin(ctx.h, 1, buf1.ctypes.data_as(C.c_void_p), 0, 0) == 0   # third in flight
    assert np.array_equal(buf2, a)                                # ... so the oldest was waited for
    assert gpu_lib.vp8b200_frame_fetch_wait(ctx.h) == 0
    assert np.array_equal(buf3, a)
    assert gpu_lib.vp8b200_frame_fetch_wait(ctx.h) == 0
    assert gpu_lib.vp8b200_frame_fetch_wait(ctx.h) == 0           # nothing in flight: a no-op
    assert gpu_lib.vp8b200_frame_fetch_begin(ctx.h, 1, buf1.ctypes.data_as(C.c_void_p), 65, 48) == -1   # wider than coded
    assert gpu_lib.vp8b200_frame_fetch_begin(ctx.h, 1, buf1.ctypes.data_as(C.c_void_p), 64, 0) == -1
    ctx.close()


def test_invalid_records_are_rejected_not_executed(gpu_lib):
    """A corrupt record must come back as an error code, never as a wild device read."""
    mb_cols, mb_rows = 6, 4
    ctx = abi.Context(mb_cols * 16, mb_rows * 16, 4)
    rng = np.random.default_rng(9)
    good = randrec.random_frame(rng, mb_cols, mb_rows, p_intra=0.0, p_split=0.0)

    def submit(fr):
        hdr = np.asarray(fr.hdr, recfile.HDR_DTYPE).tobytes()
        bufs = abi.FrameBufs()
        st = gpu_lib.vp8b200_frame_begin(ctx.h, hdr, C.byref(bufs))
        if st:
            return st                                             # the header itself was refused
        C.memmove(bufs.mb, fr.mb.ctypes.data, 16 * fr.mb.shape[0])
        if fr.n_aux:
            C.memmove(bufs.aux, fr.aux.ctypes.data, 64 * fr.n_aux)
        if fr.n_coef:
            C.memmove(bufs.coef, fr.coef.ctypes.data, 32 * fr.n_coef)
        return gpu_lib.vp8b200_frame_submit(ctx.h, fr.n_aux, fr.n_coef)

    assert submit(good) == 0
    bad = recfile.Frame(good.hdr.copy(), good.mb.copy(), good.aux.copy(), good.coef.copy(), 1, 0)
    bad.mb["flags"][5] &= 0xff ^ recfile.MBF_CLAMP                     # huge MV without the clamp flag
    bad.mb["mv_col"][5] = 30000
    assert submit(bad) == -1
    bad = recfile.Frame(good.hdr.copy(), good.mb.copy(), good.aux.copy(), good.coef.copy(), 1, 0)
    bad.mb["coef_mask"][3] = 0x1ffffff
    bad.mb["coef_off"][3] = 10 ** 6                               # arena offset out of range
    bad.mb["flags"][3] &= 0xff ^ recfile.MBF_SKIP
    assert submit(bad) == -1
    bad = recfile.Frame(good.hdr.copy(), good.mb.copy(), good.aux.copy(), good.coef.copy(), 1, 0)
    bad.mb["y_mode"][0] = 9                                       # SPLITMV pointing at a missing aux entry
    bad.mb["mv_row"][0], bad.mb["mv_col"][0] = np.array([12345], "<u4").view("<i2")
    assert submit(bad) == -1
    bad = recfile.Frame(good.hdr.copy(), good.mb.copy(), good.aux.copy(), good.coef.copy(), 1, 0)
    bad.hdr["filter_type"] = 2
    assert submit(bad) == -1
    bad = recfile.Frame(good.hdr.copy(), good.mb.copy(), good.aux.copy(), good.coef.copy(), 1, 0)
    bad.hdr["fb_new"] = bad.hdr["fb_golden"]                       # an inter frame written into its own reference
    assert submit(bad) == -1
    key = randrec.random_frame(rng, mb_cols, mb_rows, key=True)
    assert submit(key) == 0
    bpred = np.flatnonzero(key.mb["y_mode"] == 4)
    assert bpred.size
    bad = recfile.Frame(key.hdr.copy(), key.mb.copy(), key.aux.copy(), key.coef.copy(), 1, 0)
    bad.aux[int(np.array([bad.mb["mv_row"][bpred[0]], bad.mb["mv_col"][bpred[0]]], "<i2").view("<u4")[0])][7] = 10   # sub-block mode out of range
    assert submit(bad) == -1
    assert gpu_lib.vp8b200_frame_submit(ctx.h, 0, 0) == -1        # submit without begin
    assert submit(good) == 0                                      # the context stays usable
    ctx.sync()
    ctx.close()


def test_frame_abort_then_next_frame(gpu_lib):
    """The reference's longjmp error path abandons a frame between begin and submit."""
    rec, md5s = load_case("qcif_lq", max_frames=3)
    geo = frames.Geometry(rec.coded_width, rec.coded_height)
    ctx = abi.Context(rec.coded_width, rec.coded_height, rec.n_fb)
    ora = OracleDecoder(rec.coded_width, rec.coded_height, rec.n_fb)
    for i, fr in enumerate(rec.frames):
        bufs = abi.FrameBufs()
        hdr = np.asarray(fr.hdr, recfile.HDR_DTYPE).tobytes()
        assert gpu_lib.vp8b200_frame_begin(ctx.h, hdr, C.byref(bufs)) == 0
        assert gpu_lib.vp8b200_frame_abort(ctx.h) == 0            # abandoned ...
        ctx.submit(fr)                                            # ... then decoded properly
        ora.frame(fr)
        fb = int(fr.hdr["fb_new"])
        m = geo.defined_mask()                                    # row padding beyond the border is undefined
        assert np.array_equal(ctx.fetch(fb)[m], ora.fb(fb)[m])
    ctx.close()


def test_resolution_change_through_the_drop_in_decoder(gpu_lib, tmp_path):
    """Two golden clips of different size back to back in one IVF: the decoder re-creates its
    device context at the second key frame (vp8_alloc_frame_buffers path)."""
    def frames_of(name):
        d = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
        pos, out = 32, []
        while pos + 12 <= len(d):
            n = struct.unpack("<I", d[pos:pos + 4])[0]
            out.append(d[pos:pos + 12 + n])
            pos += 12 + n
        return d[:32], out
    h1, f1 = frames_of("qcif_lq")
    _, f2 = frames_of("odd_motion")
    _, f3 = frames_of("qcif_p1")
    ivf = tmp_path / "cat.ivf"
    ivf.write_bytes(h1 + b"".join(f1[:6] + f2[:6] + f3[:5]))
    want = (open(os.path.join(GOLD, "qcif_lq.md5")).read().split()[:6] +
            open(os.path.join(GOLD, "odd_motion.md5")).read().split()[:6] +
            open(os.path.join(GOLD, "qcif_p1.md5")).read().split()[:5])
    exe = os.path.join(ROOT, "hostdec", "_build", "vpxdec_b200")
    out = subprocess.run([exe, "--md5", "--i420", "-o", str(tmp_path / "f-%4.i420"), str(ivf)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    got = [l.split()[0] for l in out.stdout.splitlines() if l.strip()]
    assert got == want

"""CPU: function-level known-answer tests - every arithmetic routine of the oracle
restatement against the UNMODIFIED reference's own `_c` symbols (oracle/_ref/libvpxref.so,
compiled from /root/reference by oracle/refbuild) on seeded random inputs.  Skipped where the
reference build is absent; the golden-stream MD5 test pins the oracle independently."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
import oracle_lib

REFLIB = os.path.join(ROOT, "oracle", "_ref", "libvpxref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref not built here")


@pytest.fixture(scope="module")
def ref():
    return C.CDLL(REFLIB)


@pytest.fixture(scope="module")
def ora():
    return oracle_lib.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _coefs(rng, n, big):
    c = np.zeros((n, 16), np.int16)
    for i in range(n):
        k = rng.integers(1, 17)
        idx = rng.choice(16, k, replace=False)
        lim = 32767 if big and i % 3 == 0 else 2048
        c[i, idx] = rng.integers(-lim, lim + 1, k)
    return c


def test_idct_dc_walsh(ref, ora):
    rng = np.random.default_rng(1)
    for q in _coefs(rng, 400, True):
        pred = rng.integers(0, 256, (4, 8), dtype=np.uint8)
        a, b = pred.copy(), pred.copy()
        qa = q.copy()
        ref.vp8_short_idct4x4llm_c(_p(qa), _p(a), 8, _p(a), 8)
        ora.oracle_idct_add(_p(q.copy()), _p(b), 8)
        assert np.array_equal(a, b)
        # DC shortcut == full transform when only the DC is present
        dc = np.zeros(16, np.int16)
        dc[0] = q[0]
        a, b = pred.copy(), pred.copy()
        ref.vp8_dc_only_idct_add_c(C.c_short(int(q[0])), _p(a), 8, _p(a), 8)
        ref.vp8_short_idct4x4llm_c(_p(dc), _p(b), 8, _p(b), 8)
        assert np.array_equal(a, b)
        c = pred.copy()
        ora.oracle_dc_add(C.c_int16(int(q[0])), _p(c), 8)
        assert np.array_equal(a, c)
        # second-order transform
        out_r = np.zeros(16 * 16, np.int16)
        out_o = np.zeros(16, np.int16)
        ref.vp8_short_inv_walsh4x4_c(_p(q.copy()), _p(out_r))
        ora.oracle_iwalsh(_p(q.copy()), _p(out_o))
        assert np.array_equal(out_r[::16], out_o)
        one = np.zeros(16, np.int16)
        one[0] = q[0]
        out_1 = np.zeros(16 * 16, np.int16)
        ref.vp8_short_inv_walsh4x4_1_c(_p(one.copy()), _p(out_1))
        ora.oracle_iwalsh(_p(one.copy()), _p(out_o))
        assert np.array_equal(out_1[::16], out_o)       # _1 shortcut == full WHT on DC-only input


@pytest.mark.parametrize("w,h,name", [(4, 4, "4x4"), (8, 4, "8x4"), (8, 8, "8x8"), (16, 16, "16x16")])
def test_subpixel_filters(ref, ora, w, h, name):
    rng = np.random.default_rng(2)
    stride = 48
    for it in range(120):
        img = rng.integers(0, 256, (40, stride), dtype=np.uint8)
        if it % 4 == 0:
            img[:] = rng.choice([0, 255], img.shape)            # saturating content
        xo, yo = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        src = img.ctypes.data + 8 * stride + 8
        for fn, of in ((b"vp8_sixtap_predict", ora.oracle_sixtap), (b"vp8_bilinear_predict", ora.oracle_bilinear)):
            a = np.zeros((16, 32), np.uint8)
            b = np.zeros((16, 32), np.uint8)
            getattr(ref, (fn + name.encode() + b"_c").decode())(C.c_void_p(src), stride, xo, yo, _p(a), 32)
            of(C.c_void_p(src), stride, xo, yo, _p(b), 32, w, h)
            assert np.array_equal(a, b), (fn, name, xo, yo)


def test_intra4x4_all_modes(ref, ora):
    rng = np.random.default_rng(3)
    for it in range(300):
        for mode in range(10):
            img = rng.integers(0, 256, (8, 16), dtype=np.uint8)
            a, b = img.copy(), img.copy()
            off = 2 * 16 + 4
            ref.vp8_intra4x4_predict_c(C.c_void_p(a.ctypes.data + off), 16, mode, C.c_void_p(a.ctypes.data + off), 16)
            ora.oracle_intra4x4(C.c_void_p(b.ctypes.data + off), 16, mode)
            assert np.array_equal(a, b), mode


def test_loop_filter_edges(ref, ora):
    rng = np.random.default_rng(4)
    for it in range(400):
        base = rng.integers(0, 256, dtype=np.uint8)
        spread = int(rng.choice([2, 6, 20, 255]))
        img = np.clip(base + rng.integers(-spread, spread + 1, (24, 32)), 0, 255).astype(np.uint8)
        lvl = int(rng.integers(1, 64))
        ilim, thr = int(rng.integers(1, 10)), int(rng.integers(0, 4))
        blim, mblim = 2 * lvl + ilim, 2 * (lvl + 2) + ilim
        arr = lambda v: (C.c_ubyte * 16)(*([v] * 16))
        for mbedge in (0, 1):
            elim = mblim if mbedge else blim
            for horiz in (0, 1):
                a, b = img.copy(), img.copy()
                off = 8 * 32 + 8
                fn = {(0, 1): "vp8_loop_filter_horizontal_edge_c", (0, 0): "vp8_loop_filter_vertical_edge_c",
                      (1, 1): "vp8_mbloop_filter_horizontal_edge_c", (1, 0): "vp8_mbloop_filter_vertical_edge_c"}[(mbedge, horiz)]
                getattr(ref, fn)(C.c_void_p(a.ctypes.data + off), 32, arr(elim), arr(ilim), arr(thr), 2)
                along, across = (1, 32) if horiz else (32, 1)
                ora.oracle_edge_normal(C.c_void_p(b.ctypes.data + off), along, across, 16, mbedge, elim, ilim, thr)
                assert np.array_equal(a, b), fn
        for horiz in (0, 1):
            a, b = img.copy(), img.copy()
            off = 8 * 32 + 8
            fn = "vp8_loop_filter_simple_horizontal_edge_c" if horiz else "vp8_loop_filter_simple_vertical_edge_c"
            getattr(ref, fn)(C.c_void_p(a.ctypes.data + off), 32, arr(blim))
            along, across = (1, 32) if horiz else (32, 1)
            ora.oracle_edge_simple(C.c_void_p(b.ctypes.data + off), along, across, 16, blim)
            assert np.array_equal(a, b), fn

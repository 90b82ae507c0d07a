"""CPU: the packed two-lines-per-register loop-filter arithmetic of k_loopfilter
(libvpx.opencl_b200/csrc/lf_packed.cuh) against the UNMODIFIED reference's edge functions
(vp8/common/loopfilter_filters.c, from oracle/_ref/libvpxref.so) - or, where the reference build
is absent, against the oracle's restatement of them.  The header is compiled for the host: its
device primitives (VABSDIFF4, VIMNMX3.S16x2, VIADDMNMX.S16x2.RELU, sign-replicating PRMT) have
plain C twins, the filter code above them is the very source the kernel compiles."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
import oracle_lib

SRC = os.path.join(ROOT, "tests", "native", "lfp_host.cpp")
HDR = os.path.join(ROOT, "libvpx.opencl_b200", "csrc", "lf_packed.cuh")
LIB = os.path.join(ROOT, "oracle", "_build", "liblfp_host.so")
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libvpxref.so")


@pytest.fixture(scope="module")
def lfp():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-I" + os.path.dirname(HDR), "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.lfp_run.argtypes = [C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int]
    return L


def _lines(rng, n):
    """n lines of 8 pixels: smooth, stepped, noisy and saturated content around the edge."""
    base = rng.integers(0, 256, (n, 1))
    spread = rng.choice([0, 1, 2, 3, 5, 9, 20, 60, 255], (n, 1))
    px = base + rng.integers(-1, 2, (n, 8)) * rng.integers(0, 256, (n, 8)) % (spread + 1)
    step = rng.integers(-1, 2, (n, 1)) * rng.choice([0, 0, 2, 6, 14, 40, 130], (n, 1))
    px[:, 4:] += step
    px = np.clip(px, 0, 255)
    sat = rng.random(n) < 0.05
    px[sat] = rng.choice([0, 255, 128, 127], (int(sat.sum()), 8))
    return np.ascontiguousarray(px.astype(np.uint8))


def _reference(kind, px, ilim, elim, thr):
    """The same lines through the reference (or the oracle): lines are the rows of an image
    whose vertical edge sits at x = 4."""
    n = px.shape[0]
    img = np.zeros((n, 16), np.uint8)
    img[:, 4:12] = px
    arr = lambda v: (C.c_ubyte * 16)(*([v] * 16))
    p = C.c_void_p(img.ctypes.data + 8)
    if os.path.exists(REFLIB):
        R = C.CDLL(REFLIB)
        if kind == 0:
            R.vp8_mbloop_filter_vertical_edge_c(p, 16, arr(elim), arr(ilim), arr(thr), n // 8)
        elif kind == 1:
            R.vp8_loop_filter_vertical_edge_c(p, 16, arr(elim), arr(ilim), arr(thr), n // 8)
        else:
            assert n % 16 == 0
            for y in range(0, n, 16):
                R.vp8_loop_filter_simple_vertical_edge_c(C.c_void_p(img.ctypes.data + 16 * y + 8), 16, arr(elim))
    else:
        O = oracle_lib.lib()
        if kind == 2:
            O.oracle_edge_simple(p, 16, 1, n, elim)
        else:
            O.oracle_edge_normal(p, 16, 1, n, 1 if kind == 0 else 0, elim, ilim, thr)
    return np.ascontiguousarray(img[:, 4:12])


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_packed_edges_match_the_reference(lfp, kind):
    rng = np.random.default_rng(100 + kind)
    n = 4096
    levels = list(range(0, 64, 3)) + [1, 2, 62, 63]
    for lvl in levels:
        for sharp in (0, 1, 4, 5, 7):
            # loopfilter.c:66-96: interior limit from level and sharpness
            il = lvl >> (1 if sharp > 0 else 0)
            il >>= 1 if sharp > 4 else 0
            if sharp > 0:
                il = min(il, 9 - sharp)
            il = max(il, 1)
            for elim in (2 * lvl + il, 2 * (lvl + 2) + il):
                thr = int(rng.integers(0, 4))
                px = _lines(rng, n)
                want = _reference(kind, px.copy(), il, elim, thr)
                got = px.copy()
                lfp.lfp_run(kind, got.ctypes.data_as(C.c_void_p), n, il, elim, thr)
                bad = np.flatnonzero((got != want).any(axis=1))
                assert bad.size == 0, (kind, lvl, sharp, elim, thr, px[bad[0]], got[bad[0]], want[bad[0]])


def test_level_zero_and_skip_limits_are_the_identity(lfp):
    rng = np.random.default_rng(7)
    px = _lines(rng, 8192)
    for kind in (0, 1):
        got = px.copy()
        lfp.lfp_run(kind, got.ctypes.data_as(C.c_void_p), px.shape[0], 0, 0, 0)     # level 0: all-zero limits
        assert np.array_equal(got, px)
    got = px.copy()
    lfp.lfp_run(3, got.ctypes.data_as(C.c_void_p), px.shape[0], 9, 139, 2)          # macroblock without inner edges
    assert np.array_equal(got, px)


def test_pack_unpack_round_trip(lfp):
    rng = np.random.default_rng(8)
    px = rng.integers(0, 256, (4096, 8), dtype=np.uint8)
    got = px.copy()
    lfp.lfp_run(4, got.ctypes.data_as(C.c_void_p), px.shape[0], 0, 0, 0)
    assert np.array_equal(got, px)

"""GPU: the frame-level known-answer tests of tests/test_kat_reference_frames.py with the CUDA
path: seeded random records through the C ABI must reproduce what the UNMODIFIED reference's
own frame drivers produce (oracle/_ref/libkat.so: vp8_loop_filter_frame with random segment ids
and non-zero per-segment levels in both modes, vp8_build_inter_predictors_mb with clamped /
SPLITMV vectors at every frame edge, the intra predictors with the frame-edge rules)."""
import numpy as np
import pytest

import kat_lib
import oracle_lib
import randrec
from test_kat_reference_frames import LF_CASES, coded
from vp8b200 import abi, frames

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not kat_lib.available(), reason="oracle/_ref/libkat.so not built")]


@pytest.mark.parametrize("case", range(len(LF_CASES)))
def test_cuda_loop_filter_matches_vp8_loop_filter_frame(gpu_lib, case):
    """CUDA(recon + loop filter + borders) == borders(reference loop filter(recon)): the frame
    before the filter comes from the oracle's reconstruction stage (pinned separately), the
    filter under test is the reference's own vp8_loop_filter_frame."""
    kw = LF_CASES[case]
    rng = np.random.default_rng(5000 + case)
    mb_cols, mb_rows = 9, 6                                    # two warps of the filter kernel, a partial one
    geo = frames.Geometry(mb_cols * 16, mb_rows * 16)
    ctx = abi.Context(geo.w, geo.h, 4)
    ora = oracle_lib.OracleDecoder(geo.w, geo.h, 4)
    changed = 0
    for rep in range(3):
        fr = randrec.random_frame(rng, mb_cols, mb_rows, filter_level=int(rng.integers(1, 64)), p_skip=0.5,
                                  coef_density=0.15, **kw)
        bufs = [b for b in randrec.random_buffers(rng, geo.frame_size, 4)]
        for i in range(4):
            # smooth references, so that the reconstruction has edges the filter acts on
            bufs[i] = np.clip(128 + np.cumsum(rng.integers(-1, 2, geo.frame_size)) % 64, 0, 255).astype(np.uint8)
            ctx.upload(i, bufs[i])
            ora.fb(i)[:] = bufs[i]
        fb = int(fr.hdr["fb_new"])
        ora.frame(fr, stages=1)
        pre = ora.fb(fb).copy()
        ref = kat_lib.loop_filter_frame(geo.w, geo.h, fr, pre.copy())
        changed += int((ref != pre).sum())
        ora.fb(fb)[:] = ref
        ora.frame(fr, stages=4)                                 # border extension of the filtered frame
        ctx.submit(fr)
        got = ctx.fetch(fb)
        m = geo.defined_mask()
        assert np.array_equal(got[m], ora.fb(fb)[m]), (kw, rep, int((got[m] != ora.fb(fb)[m]).sum()))
    assert changed > 300, "the filter hardly did anything: weak test"
    ctx.close()
    ora.close()


@pytest.mark.parametrize("bilinear,full_pixel", [(False, False), (True, False), (True, True)])
def test_cuda_inter_prediction_matches_vp8_build_inter_predictors_mb(gpu_lib, bilinear, full_pixel):
    rng = np.random.default_rng(6000 + 2 * bilinear + full_pixel)
    for mb_cols, mb_rows in ((6, 4), (3, 7), (1, 1)):
        geo = frames.Geometry(mb_cols * 16, mb_rows * 16)
        ctx = abi.Context(geo.w, geo.h, 4)
        for rep in range(4):
            fr = randrec.random_frame(rng, mb_cols, mb_rows, bilinear=bilinear, full_pixel=full_pixel, filter_level=0,
                                      p_intra=0.0, p_split=0.4, p_skip=1.0)
            bufs = randrec.random_buffers(rng, geo.frame_size, 4)
            for i in range(4):
                ctx.upload(i, bufs[i])
            want = kat_lib.inter_frame(geo.w, geo.h, fr, [b.copy() for b in bufs])
            ctx.submit(fr)
            got = ctx.fetch(0)
            assert np.array_equal(coded(geo, got), coded(geo, want)), (mb_cols, mb_rows, rep)
        ctx.close()


def test_cuda_intra_prediction_matches_the_reference_predictors(gpu_lib):
    rng = np.random.default_rng(7000)
    for mb_cols, mb_rows in ((6, 5), (2, 9), (1, 1), (9, 1)):
        geo = frames.Geometry(mb_cols * 16, mb_rows * 16)
        ctx = abi.Context(geo.w, geo.h, 4)
        for rep in range(6):
            fr = randrec.random_frame(rng, mb_cols, mb_rows, key=True, filter_level=0, p_skip=1.0)
            start = rng.integers(0, 256, geo.frame_size, dtype=np.uint8)
            ctx.upload(0, start)
            want = kat_lib.intra_frame(geo.w, geo.h, fr, start.copy())
            ctx.submit(fr)
            got = ctx.fetch(0)
            assert np.array_equal(coded(geo, got), coded(geo, want)), (mb_cols, mb_rows, rep)
        ctx.close()

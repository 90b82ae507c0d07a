#!/usr/bin/env python3
"""Debug helper (GPU): the 1080p random P frame of tests/test_gpu_parity.py through the library
selected by VP8B200_LIB, every differing coded pixel listed by plane / macroblock / position."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import randrec
from oracle_lib import OracleDecoder
from vp8b200 import abi, frames

mb_cols, mb_rows = 120, 68
w, h = mb_cols * 16, mb_rows * 16
geo = frames.Geometry(w, h)
rng = np.random.default_rng(5)
ctx = abi.Context(w, h, 4)
ora = OracleDecoder(w, h, 4)
bufs = randrec.random_buffers(rng, geo.frame_size, 4)
for fb, buf in enumerate(bufs):
    ctx.upload(fb, buf); ora.fb(fb)[:] = buf
fr = randrec.random_frame(rng, mb_cols, mb_rows, fbs=(0, 1, 2, 3))
ctx.submit(fr)
got = ctx.fetch(0).copy()
ora.frame(fr, stages=1)
pre = ora.fb(0).copy()
ora.frame(fr, stages=2)
want = ora.fb(0)
print("hdr", {k: fr.hdr[k].tolist() for k in ("filter_type", "filter_level", "sharpness_level", "segmentation_enabled", "segment_abs_delta", "mode_ref_lf_delta_enabled", "segment_lf", "ref_lf_deltas", "mode_lf_deltas")})
for name, pg, pw, pp, bs in zip("YUV", geo.planes(got), geo.planes(want), geo.planes(pre), (16, 8, 8)):
    d = np.argwhere(pg != pw)
    print(name, "differing coded pixels:", len(d), " pixels the oracle's filter changed:", int((pw != pp).sum()))
    for (y, x) in d[:60]:
        i = (y // bs) * mb_cols + x // bs
        m = fr.mb[i]
        print("  %s (%d,%d) mb(r%d,c%d) in-mb(y%d,x%d) got %d want %d pre %d | mode %d ref %d flags %d" %
              (name, y, x, y // bs, x // bs, y % bs, x % bs, pg[y, x], pw[y, x], pp[y, x], m["y_mode"], m["ref_frame"], m["flags"]))

#!/usr/bin/env python3
"""Per-step, per-kernel device times of the batched replay (debug / profiling helper).
usage: tools/kernel_times.py [--streams 64] [--frames 8]"""
import argparse, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200")); sys.path.insert(0, ROOT)
import bench
from vp8b200 import abi, recfile

ap = argparse.ArgumentParser(); ap.add_argument("--streams", type=int, default=64); ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
clips = bench.find_clips()[:a.streams]
with tempfile.TemporaryDirectory() as tmp:
    recs = [recfile.read(p, a.frames) for p in bench.capture_records(clips, tmp)]
S = a.streams
r0 = recs[0]
ctxs = [abi.Context(r0.coded_width, r0.coded_height, r0.n_fb) for _ in range(S)]
staged = [[ctxs[0].stage(recs[s % len(recs)].frames[f]) for s in range(S)] for f in range(a.frames)]
ctxs[0].profile(True)
import ctypes
_dbg = getattr(abi.lib(), 'vp8b200_debug_intra_prof', None)
if _dbg is not None:
    _buf = (ctypes.c_ulonglong * 8)(); _dbg(_buf)
for rep in range(a.reps):
    for f in range(a.frames):
        abi.batch_run(ctxs, staged[f])
        p = ctxs[0].profile_read()
        if rep == a.reps - 1:
            fr = recs[0].frames[f]
            n_intra = [int((recs[s % len(recs)].frames[f].mb["ref_frame"] == 0).sum()) for s in range(S)]
            print("frame %2d type %d intraMBs(s0) %5d all %6d max %5d | " % (f, fr.hdr["frame_type"], int((fr.mb["ref_frame"] == 0).sum()), sum(n_intra), max(n_intra)) +
                  "  ".join("%s %.3f ms" % (k, v[0]) for k, v in p.items()))
            if _dbg is not None:
                _dbg(_buf)
                n = max(1, _buf[4])
                print("    intra B_PRED MBs %d: cycles per MB wait %.0f scatter %.0f predict %.0f export %.0f" % (_buf[4], _buf[0] / n, _buf[1] / n, _buf[2] / n, _buf[3] / n))
        elif _dbg is not None:
            _dbg(_buf)

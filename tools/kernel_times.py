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
for rep in range(a.reps):
    for f in range(a.frames):
        abi.batch_run(ctxs, staged[f])
        p = ctxs[0].profile_read()
        if rep == a.reps - 1:
            fr = recs[0].frames[f]
            n_intra = [int((recs[s % len(recs)].frames[f].mb["ref_frame"] == 0).sum()) for s in range(S)]
            print("frame %2d type %d intraMBs(s0) %5d all %6d max %5d | " % (f, fr.hdr["frame_type"], int((fr.mb["ref_frame"] == 0).sum()), sum(n_intra), max(n_intra)) +
                  "  ".join("%s %.3f ms" % (k, v[0]) for k, v in p.items()))

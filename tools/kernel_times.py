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
    _buf = (ctypes.c_ulonglong * 184)(); _dbg(_buf)
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
                t0 = min(x for x in _buf[24:184:2] if x) if any(_buf[24:184:2]) else 0
                print("    rows of job 0 (start us, duration us): " + " ".join("%d:%.0f+%.0f" % (r, (_buf[24 + 2 * r] - t0) / 1e3, (_buf[25 + 2 * r] - _buf[24 + 2 * r]) / 1e3) for r in (0, 1, 2, 3, 8, 16, 32, 48, 66, 67) if _buf[24 + 2 * r]))
                print("    above-right hand-off: %.0f ns from the producer's export to the consumer's fetch (%d samples)" % (_buf[11] / max(1, _buf[12]), _buf[12]))
                _ts = getattr(abi.lib(), 'vp8b200_debug_intra_ts', None)
                if _ts is not None and fr.hdr["frame_type"] == 0:
                    tb = (ctypes.c_ulonglong * 8192)(); _ts(tb)
                    cols = (r0.coded_width + 15) // 16
                    T = lambda r, c: tb[r * cols + c]
                    base_t = T(0, 0)
                    td = (ctypes.c_ulonglong * 8192)(); abi.lib().vp8b200_debug_intra_td(td)
                    t0b = (ctypes.c_ulonglong * 8192)(); abi.lib().vp8b200_debug_intra_t0(t0b)
                    ta = (ctypes.c_ulonglong * 8192)(); abi.lib().vp8b200_debug_intra_ta(ta)
                    tb2 = (ctypes.c_ulonglong * 8192)(); abi.lib().vp8b200_debug_intra_tb(tb2)
                    for r in (1, 2):
                        print("    row %d MB 0 and 1: start %.1f borders %.1f scattered %.1f predicted %.1f exported %.1f | %.1f %.1f %.1f %.1f %.1f" % ((r,) + tuple((x[r * cols + c] - base_t) / 1e3 for c in (0, 1) for x in (t0b, td, ta, tb2, tb))))
                    for r in (0, 1, 2, 3):
                        print("    row %d start / borders / export (us after (0,0) export): %s" % (r, " ".join("%.1f/%.1f/%.1f" % ((t0b[r * cols + c] - base_t) / 1e3, (td[r * cols + c] - base_t) / 1e3, (T(r, c) - base_t) / 1e3) for c in range(0, 7))))
                    for r in (1, 2, 3, 10, 30):
                        if (r + 1) * cols > 8192: break
                        pace = [(T(r, c) - T(r, c - 1)) for c in range(1, cols)]
                        dep = [(T(r, c) - T(r - 1, c + 1)) for c in range(0, cols - 1)]
                        print("    row %2d: export pace ns (first 12) %s ... median %d | behind (r-1,c+1) ns: first %s median %d" % (r, pace[:12], sorted(pace)[len(pace) // 2], dep[:6], sorted(dep)[len(dep) // 2]))
                h = max(1, _buf[9])
                print("    cycles per MB (%d MBs): luma warp slot wait %.0f fetch %.0f predict %.0f export %.0f | chroma warp slot wait %.0f mb %.0f | helper slot wait %.0f prepare %.0f" % (_buf[4], _buf[0] / n, _buf[1] / n, _buf[2] / n, _buf[3] / n, _buf[5] / n, _buf[6] / n, _buf[10] / h, _buf[8] / h))
        elif _dbg is not None:
            _dbg(_buf)

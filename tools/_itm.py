import ctypes, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200")); sys.path.insert(0, ROOT)
import bench
from vp8b200 import abi, recfile
S = int(sys.argv[1])
clips = bench.find_clips()[:S]
with tempfile.TemporaryDirectory() as tmp:
    recs = [recfile.read(p, 1) for p in bench.capture_records(clips, tmp)]
r0 = recs[0]
ctxs = [abi.Context(r0.coded_width, r0.coded_height, r0.n_fb) for _ in range(S)]
staged = [ctxs[0].stage(recs[s].frames[0]) for s in range(S)]
L = abi.lib()
out = (ctypes.c_ulonglong * 8)()
for rep in range(3):
    L.vp8b200_debug_itm(out, 1)
    abi.batch_run(ctxs, staged); ctxs[0].sync()
    L.vp8b200_debug_itm(out, 0)
n = out[5]
print("streams", S, "MBs", n, "bpred", out[6], "| cycles per MB: prologue+residual %.0f  wait %.0f  scatter+dc+16x16/chroma %.0f  bpred loop %.0f  rowout+export %.0f" % tuple(out[i] / n for i in range(5)))

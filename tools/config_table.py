#!/usr/bin/env python3
"""Per-configuration table of SURVEY.md section 8(d): CIF, 720p, 1080p (profiles 0/1/3) and 2160p.

For every stream it reports, on one GPU:
  single   resident replay of ONE stream (records already in HBM): frames/s and per-kernel ms
  batch    resident replay of B copies of the stream in one batched launch per kernel
           (B = 64 up to 1080p, 16 at 2160p): frames/s, per-kernel ms and the loop filter's
           achieved algorithmic GB/s
  e2e      one decoder instance through vpx_codec_decode / vpx_codec_get_frame (host IVF ->
           host frame), i.e. single-stream latency-bound fps; and as many instances as host
           cores
  cpu      the unmodified reference decoder, one process, same stream (oracle/_ref/refbench)
All device times are CUDA-event times from the library's profiling spans.
usage: tools/config_table.py [--out profiles/r01_configs.json]
"""
import argparse, json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200")); sys.path.insert(0, ROOT)
import bench
from vp8b200 import abi, recfile

CONFIGS = [("c1_cif", os.path.join(ROOT, "tests", "golden", "c1_cif.ivf"), 64),
           ("c2_720p", os.path.join(bench.STREAMS, "c2_720p.ivf"), 64),
           ("c3_1080p_p3", os.path.join(bench.STREAMS, "c3_1080p_p3.ivf"), 64),
           ("c3b_1080p_p1", os.path.join(bench.STREAMS, "c3b_1080p_p1.ivf"), 64),
           ("c5_1080p_p0", os.path.join(bench.STREAMS, "c5_1080p_s100.ivf"), 64),
           ("c4_2160p", os.path.join(bench.STREAMS, "c4_2160p.ivf"), 16)]


def replay(rec, B, reps=3):
    """-> (frames/s, {kernel: ms per clip pass}, loop-filter algorithmic bytes per pass)"""
    F = len(rec.frames)
    ctxs = [abi.Context(rec.coded_width, rec.coded_height, rec.n_fb) for _ in range(B)]
    staged = [[ctxs[0].stage(rec.frames[f]) for _ in range(B)] for f in range(F)]
    ctxs[0].profile(True)
    best = None
    for _ in range(reps):
        tot = {}
        for f in range(F):
            abi.batch_run(ctxs, staged[f])
            for k, v in ctxs[0].profile_read().items():
                tot[k] = tot.get(k, 0.0) + v[0]
        if best is None or sum(tot.values()) < sum(best.values()):
            best = tot
    na = rec.coded_width * rec.coded_height
    lf_bytes = sum(bench.frame_bytes_model(fr, na)["loopfilter"] for fr in rec.frames) * B
    for c in ctxs:                          # ctxs[0].close() frees the staged frames
        c.close()
    ms = sum(best.values())
    return B * F / (ms * 1e-3), {k: round(v, 3) for k, v in best.items()}, lf_bytes


def e2e(ivf, n):
    env = dict(os.environ, VP8B200_SYNC="block")
    out = subprocess.run([os.path.join(bench.HOSTDEC, "b200bench"), "--threads", str(n), "--streams", str(n),
                          "--repeat", "3", ivf], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if out.returncode:
        raise SystemExit(out.stderr[-400:])
    return json.loads(out.stdout.strip().splitlines()[-1])["fps"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r01_configs.json"))
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, ivf, B in CONFIGS:
            rec = recfile.read(bench.capture_records([ivf], tmp)[0])
            f1, k1, _ = replay(rec, 1)
            fb, kb, lfb = replay(rec, B)
            cpu = bench.run_refbench([ivf], 1, 2, procs=1)
            row = {"config": name, "coded": [rec.coded_width, rec.coded_height], "frames": len(rec.frames),
                   "single_stream_resident_fps": round(f1, 1), "single_stream_kernel_ms": k1,
                   "batch": B, "batch_resident_fps": round(fb, 1), "batch_kernel_ms": kb,
                   "batch_loopfilter_GBps": round(lfb / (kb["loopfilter"] * 1e-3) / 1e9, 1) if kb.get("loopfilter") else None,
                   "e2e_1_instance_fps": round(e2e(ivf, 1), 1), "e2e_%d_instances_fps" % cores: round(e2e(ivf, cores), 1),
                   "reference_cpu_1_process_fps": cpu["value"]}
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"host_cores": cores, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

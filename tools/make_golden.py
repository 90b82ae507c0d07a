#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ FROM THE REFERENCE.

For every case below:
  synthetic Y4M (tools/gen_y4m.py)  --reference vpxenc-->  <name>.ivf
  <name>.ivf  --reference vpxdec --md5 --i420 (generic C)-->  <name>.md5   (one line per shown frame)
  <name>.ivf  --hostdec in record-capture mode-->  <name>.rec.xz          (the C-ABI input)
  <name>.json header probe: profile, filter type, partitions, segmentation, modes seen

Runs in the build container only (needs /root/reference compiled into oracle/_ref by
oracle/refbuild and hostdec/_build).  The fixtures travel with the repository, so neither
tests nor bench need the reference at run time.

usage: tools/make_golden.py [case ...]
"""
import json
import lzma
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m  # noqa: E402
from vp8b200 import recfile  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
HOSTDEC = os.path.join(ROOT, "hostdec", "_build", "vpxdec_b200")
GOLD = os.path.join(ROOT, "tests", "golden")

# name: (kind, size, frames, seed, vpxenc args)
CASES = {
    # C1 of BASELINE.json / SURVEY.md 8d exactly: CIF, 60 frames, seed 1, moving gradient + noise
    "c1_cif": ("gradient", "352x288", 60, 1,
               ["--good", "--cpu-used=2", "--target-bitrate=800", "--kf-max-dist=30"]),
    # a shorter CIF clip with a second key frame inside 30 frames (smoke test, bench harness tests)
    "cif_p0": ("gradient", "352x288", 30, 1,
               ["--good", "--cpu-used=2", "--target-bitrate=800", "--kf-max-dist=20"]),
    # profile 1: bilinear MC + simple loop filter
    "qcif_p1": ("texture", "176x144", 20, 11,
                ["--good", "--cpu-used=2", "--profile=1", "--target-bitrate=300"]),
    # profile 3: bilinear, full-pixel chroma (encoder forces filter_level 0)
    "qcif_p3": ("texture", "176x144", 20, 12,
                ["--good", "--cpu-used=2", "--profile=3", "--target-bitrate=300"]),
    # 8 token partitions + error resilient (segmentation on), real-time mode
    "w320_er8": ("motion", "320x240", 20, 13,
                 ["--rt", "--cpu-used=4", "--token-parts=3", "--error-resilient=1",
                  "--target-bitrate=600"]),
    # dimensions that are not multiples of 16, high motion (split MVs, clamped MVs)
    "odd_motion": ("motion", "200x150", 20, 14,
                   ["--good", "--cpu-used=1", "--target-bitrate=500", "--kf-max-dist=9999"]),
    # two-pass with alt-ref frames (show_frame = 0, golden/altref references), sharpness
    "qcif_arf": ("texture", "176x144", 30, 15,
                 ["--good", "--cpu-used=2", "--passes=2", "--auto-alt-ref=1", "--lag-in-frames=16",
                  "--sharpness=3", "--target-bitrate=200", "--kf-max-dist=9999"]),
    # low quantizer / high bitrate: dense coefficients, strong loop filter off
    "qcif_hq": ("gradient", "176x144", 10, 16,
                ["--good", "--cpu-used=2", "--min-q=2", "--max-q=8", "--target-bitrate=4000"]),
    # very low bitrate: strong loop filter, many skipped MBs
    "qcif_lq": ("motion", "176x144", 20, 17,
                ["--good", "--cpu-used=2", "--min-q=50", "--max-q=63", "--target-bitrate=40"]),
}


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def make(name):
    kind, size, nframes, seed, encargs = CASES[name]
    w, h = (int(x) for x in size.split("x"))
    os.makedirs(GOLD, exist_ok=True)
    ivf = os.path.join(GOLD, name + ".ivf")
    with tempfile.TemporaryDirectory() as tmp:
        y4m = os.path.join(tmp, "in.y4m")
        gen_y4m.write_y4m(y4m, kind, w, h, nframes, seed)
        run([os.path.join(REF, "vpxenc"), "--ivf", "-o", ivf] + encargs + [y4m], cwd=tmp)
        out = run([os.path.join(REF, "vpxdec"), "--md5", "--i420", "-o",
                   os.path.join(tmp, "f-%4.i420"), ivf]).stdout.decode()
        md5s = [l.split()[0] for l in out.splitlines() if l.strip()]
        with open(os.path.join(GOLD, name + ".md5"), "w") as f:
            f.write("\n".join(md5s) + "\n")
        rec_path = os.path.join(tmp, "out.rec")
        env = dict(os.environ, VP8B200_NO_DEVICE="1", VP8B200_DUMP=rec_path)
        run([HOSTDEC, "--noblit", ivf], env=env)
        raw = open(rec_path, "rb").read()
        with open(os.path.join(GOLD, name + ".rec.xz"), "wb") as f:
            f.write(lzma.compress(raw, preset=9 | lzma.PRESET_EXTREME))
    rec = recfile.parse(raw)
    probe = {
        "display": [rec.display_width, rec.display_height],
        "coded": [rec.coded_width, rec.coded_height],
        "frames": len(rec.frames), "shown": len(md5s),
        "key_frames": int(sum(int(fr.hdr["frame_type"]) == 0 for fr in rec.frames)),
        "hidden_frames": int(sum(not fr.show_frame for fr in rec.frames)),
        "bilinear": sorted({int(fr.hdr["use_bilinear_mc"]) for fr in rec.frames}),
        "full_pixel": sorted({int(fr.hdr["full_pixel"]) for fr in rec.frames}),
        "filter_type": sorted({int(fr.hdr["filter_type"]) for fr in rec.frames}),
        "filter_level": sorted({int(fr.hdr["filter_level"]) for fr in rec.frames}),
        "sharpness": sorted({int(fr.hdr["sharpness_level"]) for fr in rec.frames}),
        "segmentation": sorted({int(fr.hdr["segmentation_enabled"]) for fr in rec.frames}),
        "lf_deltas": sorted({int(fr.hdr["mode_ref_lf_delta_enabled"]) for fr in rec.frames}),
        "y_modes": sorted({int(m) for fr in rec.frames for m in set(fr.mb["y_mode"].tolist())}),
        "ref_frames": sorted({int(m) for fr in rec.frames for m in set(fr.mb["ref_frame"].tolist())}),
        "clamped_mvs": int(sum(int(((fr.mb["flags"] & 8) != 0).sum()) for fr in rec.frames)),
        "encoder_args": encargs, "content": [kind, size, nframes, seed],
        "ivf_bytes": os.path.getsize(ivf), "rec_bytes": len(raw),
    }
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(probe, f, indent=1)
    print(name, json.dumps(probe))


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        make(n)

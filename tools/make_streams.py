#!/usr/bin/env python3
"""Generate the benchmark / full-size parity streams of BASELINE.json's configs into streams/
(git-ignored; it travels to the GPU box with the working tree).  Build-container only: needs
the reference encoder/decoder built by oracle/refbuild.

For each stream: synthetic Y4M -> reference vpxenc -> <name>.ivf, and the reference decoder's
per-frame MD5s (vpxdec --md5 --i420, generic C) -> <name>.md5.

usage: tools/make_streams.py [--jobs N] [--c5 N_STREAMS] [name ...]
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_y4m  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "streams")

C2 = ["--good", "--cpu-used=3", "--profile=0", "--kf-max-dist=9999", "--auto-alt-ref=0"]


def cases(n_c5):
    c = {
        # C2: 720p P-frame-heavy, profile 0 (six-tap, normal loop filter)
        "c2_720p": ("texture", "1280x720", 120, 2, C2 + ["--target-bitrate=3000"]),
        # C3: 1080p profile 3 (bilinear, full pixel; the encoder forces filter_level 0)
        "c3_1080p_p3": ("texture", "1920x1080", 60, 3, ["--good", "--cpu-used=3", "--profile=3",
                                                          "--target-bitrate=6000", "--kf-max-dist=9999"]),
        # C3b: profile 1 so that the SIMPLE loop filter really runs (SURVEY.md 8d caveat)
        "c3b_1080p_p1": ("texture", "1920x1080", 60, 3, ["--good", "--cpu-used=3", "--profile=1",
                                                           "--target-bitrate=6000", "--kf-max-dist=9999"]),
        # C4: 2160p high motion, 8 token partitions, error resilient (segmentation)
        "c4_2160p": ("motion", "3840x2160", 30, 4, ["--rt", "--cpu-used=4", "--token-parts=3",
                                                     "--error-resilient=1", "--target-bitrate=20000"]),
    }
    # C5: independent 1080p streams, C2-style settings, seeds 100..
    for s in range(n_c5):
        c["c5_1080p_s%03d" % (100 + s)] = ("texture", "1920x1080", 30, 100 + s, C2 + ["--target-bitrate=6000"])
    return c


def make(item):
    name, (kind, size, nframes, seed, encargs) = item
    ivf = os.path.join(OUT, name + ".ivf")
    md5 = os.path.join(OUT, name + ".md5")
    if os.path.exists(ivf) and os.path.exists(md5):
        return name + " (cached)"
    w, h = (int(x) for x in size.split("x"))
    with tempfile.TemporaryDirectory() as tmp:
        y4m = os.path.join(tmp, "in.y4m")
        gen_y4m.write_y4m(y4m, kind, w, h, nframes, seed)
        subprocess.run([os.path.join(REF, "vpxenc"), "--ivf", "-o", ivf + ".tmp"] + encargs + [y4m],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
        out = subprocess.run([os.path.join(REF, "vpxdec"), "--md5", "--i420", "-o",
                              os.path.join(tmp, "f-%4.i420"), ivf + ".tmp"], check=True,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    with open(md5, "w") as f:
        f.write("\n".join(l.split()[0] for l in out.splitlines() if l.strip()) + "\n")
    os.rename(ivf + ".tmp", ivf)
    return "%s %d bytes" % (name, os.path.getsize(ivf))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--c5", type=int, default=64)
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    todo = [(k, v) for k, v in cases(a.c5).items() if not a.names or k in a.names]
    with cf.ProcessPoolExecutor(a.jobs) as ex:
        for r in ex.map(make, todo):
            print(r, flush=True)

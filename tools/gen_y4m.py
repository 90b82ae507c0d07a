#!/usr/bin/env python3
"""Synthetic Y4M content for the BASELINE.json configs.

All clips are C420jpeg, 30 fps, numpy default_rng(seed).  Three content kinds:
  gradient : moving diagonal gradient + sigma~8 noise           (C1)
  texture  : translated random texture, 2-5 px/frame sub-pel drift,
             gradient chroma                                     (C2, C3, C5)
  motion   : high motion 8-24 px/frame, a different direction per quadrant,
             plus noise                                          (C4)

usage: gen_y4m.py --kind texture --size 1920x1080 --frames 60 --seed 3 -o out.y4m
"""
import argparse
import sys

import numpy as np


def _smooth_texture(rng, h, w, octaves=(64, 16, 4, 1)):
    """Band-limited random texture in [0,255] (sum of upsampled noise octaves)."""
    acc = np.zeros((h, w), np.float32)
    amp = 1.0
    for o in octaves:
        gh, gw = h // o + 2, w // o + 2
        g = rng.standard_normal((gh, gw)).astype(np.float32)
        up = np.kron(g, np.ones((o, o), np.float32))[:h, :w]
        acc += amp * up
        amp *= 0.6
    acc -= acc.min()
    acc *= 255.0 / max(acc.max(), 1e-6)
    return acc


def _shift_subpel(img, dy, dx):
    """Translate `img` (toroidal) by a fractional offset with bilinear weights."""
    iy, ix = int(np.floor(dy)), int(np.floor(dx))
    fy, fx = dy - iy, dx - ix
    a = np.roll(img, (iy, ix), (0, 1))
    b = np.roll(img, (iy, ix + 1), (0, 1))
    c = np.roll(img, (iy + 1, ix), (0, 1))
    d = np.roll(img, (iy + 1, ix + 1), (0, 1))
    return (1 - fy) * ((1 - fx) * a + fx * b) + fy * ((1 - fx) * c + fx * d)


def frames(kind, w, h, n, seed):
    rng = np.random.default_rng(seed)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    cyy, cxx = np.mgrid[0:ch, 0:cw].astype(np.float32)
    if kind == "gradient":
        for t in range(n):
            y = ((xx + yy) * 0.35 + 3.0 * t) % 256.0
            y = y + rng.normal(0, 8, (h, w))
            u = 128 + 40 * np.sin((cxx + 2 * t) / 37.0)
            v = 128 + 40 * np.cos((cyy - 3 * t) / 29.0)
            yield y, u, v
    elif kind == "texture":
        tex = _smooth_texture(rng, h, w)
        vy, vx = rng.uniform(2, 5, 2) * rng.choice([-1, 1], 2)
        for t in range(n):
            y = _shift_subpel(tex, vy * t, vx * t) + rng.normal(0, 1.5, (h, w))
            u = 128 + 50 * np.sin((cxx + cyy + 2 * t) / 53.0)
            v = 128 + 50 * np.cos((cxx - cyy - 2 * t) / 47.0)
            yield y, u, v
    elif kind == "motion":
        tex = _smooth_texture(rng, h, w, octaves=(32, 8, 2, 1))
        vel = rng.uniform(8, 24, (4, 2)) * rng.choice([-1, 1], (4, 2))
        hh, hw = h // 2, w // 2
        for t in range(n):
            y = np.empty((h, w), np.float32)
            q = 0
            for (r0, r1) in ((0, hh), (hh, h)):
                for (c0, c1) in ((0, hw), (hw, w)):
                    sh = _shift_subpel(tex, vel[q, 0] * t, vel[q, 1] * t)
                    y[r0:r1, c0:c1] = sh[r0:r1, c0:c1]
                    q += 1
            y = y + rng.normal(0, 4, (h, w))
            u = 128 + 60 * np.sin((cxx * 0.7 + 11 * t) / 31.0)
            v = 128 + 60 * np.cos((cyy * 0.9 - 13 * t) / 23.0)
            yield y, u, v
    else:
        raise SystemExit("unknown kind " + kind)


def write_y4m(path, kind, w, h, n, seed):
    out = sys.stdout.buffer if path == "-" else open(path, "wb")
    out.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h))
    for y, u, v in frames(kind, w, h, n, seed):
        out.write(b"FRAME\n")
        for p in (y, u, v):
            out.write(np.clip(np.rint(p), 0, 255).astype(np.uint8).tobytes())
    if out is not sys.stdout.buffer:
        out.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="texture")
    ap.add_argument("--size", default="352x288")
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("-o", "--out", default="-")
    a = ap.parse_args()
    W, H = (int(x) for x in a.size.split("x"))
    write_y4m(a.out, a.kind, W, H, a.frames, a.seed)

#!/usr/bin/env python
"""Stroke-to-fill: GPU (pfcu_stroke_to_fill: two kernel passes + the call's H2D / D2H) beside the reference's CPU stroker
(OutlineStrokeToFill::offset through oracle/_ref/libpfref.so, one thread: the reference strokes on the caller's thread).
Workloads: every stroked shape of tiger.svg and features.svg, and N random cubic blobs (glyph-density scene)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pathfinder-cpp_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import pfcu  # noqa: E402
import pfref  # noqa: E402
import scenes  # noqa: E402

lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
r = pfcu.Renderer(0, lut)


def run(name, pts, flags, first, closed, style):
    idx = np.zeros(len(closed), "<u4")
    for _ in range(3):
        out = r.stroke_to_fill(pts, flags, first, closed, idx, [style])
    wall = []
    for _ in range(10):
        t0 = time.perf_counter()
        out = r.stroke_to_fill(pts, flags, first, closed, idx, [style])
        wall.append((time.perf_counter() - t0) * 1e3)
    cpu = []
    for _ in range(3):
        cpu.append(pfref.stroke_outline(pts, flags, first, closed, *style)[3])
    print("%-28s %7d contours %8d -> %8d points | GPU passes %.3f ms, call (H2D + 2 passes + D2H) %.3f ms | reference CPU "
          "stroker %.3f ms | x%.1f" % (name, len(closed), len(pts), len(out[0]), out[3], float(np.median(wall)),
                                      float(np.median(cpu)), float(np.median(cpu)) / float(np.median(wall))))


for asset in ("tiger.svg", "features.svg"):
    d = pfref.svg_stroke_inputs(pfref.asset(asset))
    run(asset + " strokes", d["points"], d["flags"], d["contour_first"], d["closed"], (float(np.median(d["styles"][:, 0])), 0, 0, 4.0))
for n in (2000, 20000, 200000):
    paths, _ = scenes.synthetic_paths(n, 8192)
    pts = np.concatenate([p["contours"][0][0] for p in paths])
    flags = np.concatenate([p["contours"][0][1] for p in paths])
    first = np.concatenate([[0], np.cumsum([len(p["contours"][0][0]) for p in paths])]).astype("<u4")
    run("%d cubic blobs, round joins" % n, pts, flags, first, np.ones(n, "u1"), (1.5, 0, 2, 10.0))
r.close()

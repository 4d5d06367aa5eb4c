#!/usr/bin/env python
"""Incremental frames (pfcu_update_scene_range + PFCU_OPT_INCREMENTAL_DICE): what a frame costs when k of the N paths of a
scene moved, against re-uploading and re-dicing everything (what the reference does every frame,
core/d3d11/renderer.cpp:314, scene_builder.cpp:217-218). Scene: N random cubic blobs at SIZE^2 (tests/scenes.py); the moved
blobs are consecutive paths, so their points are one range. Wall clock per frame through the C-ABI (upload + metadata +
frame + counters back), dice stage time from CUDA events."""
import copy
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pathfinder-cpp_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import pfcu  # noqa: E402
import scenes  # noqa: E402

N, SIZE = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200000, 8192)
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
paths, colors = scenes.synthetic_paths(N, SIZE)
base = scenes.build_scene_from_outlines(SIZE, SIZE, paths, colors)


def moved_scene(first, k, dx, dy):
    q = list(paths)
    for i in range(first, first + k):
        p = copy.copy(paths[i])
        pts, fl = p["contours"][0]
        p["contours"] = [(pts + np.array([dx, dy], "<f4"), fl)]
        q[i] = p
    return scenes.build_scene_from_outlines(SIZE, SIZE, q, colors)


def timed(fn, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        st = fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), st


r = pfcu.Renderer(0, lut)
r.set_retain_frame_graph(False)
r.set_scene(base)
r.draw()
r.draw()
full_ms, st = timed(lambda: r.draw(clear=True, upload=True))
r.set_profiling(True)
r.draw(clear=True, upload=True)
dice_full = r.stage_times()["dice"]
r.set_profiling(False)
print("%d paths, %d segments at %d^2: full frame (upload %d KB + dice everything) %.3f ms wall, dice stage %.3f ms"
      % (N, st["segments"], SIZE, st["uploaded_bytes"] >> 10, full_ms, dice_full))
r.set_incremental_dice(True)
r.draw()
for k in (1, 16, 256, 4096):
    if k > N // 4:
        break
    first = N // 3
    a, b = moved_scene(first, k, 3.0, -2.0), moved_scene(first, k, -3.0, 2.0)
    r.update_scene(a)
    plans = [r.plan_update(a, b), r.plan_update(b, a)]  # (host-side diffing and descriptor building: not timed)
    state = [0]

    def frame():
        r.apply_update(plans[state[0]])
        state[0] ^= 1
        return r.draw(clear=True)

    frame()
    ms, st = timed(frame)
    r.set_profiling(True)
    frame()
    dice = r.stage_times()["dice"]
    r.set_profiling(False)
    print("  %5d paths moved: %7d segments diced, %7d KB uploaded, frame %.3f ms wall (x%.2f), dice stage %.3f ms"
          % (k, st["diced_segments"], st["uploaded_bytes"] >> 10, ms, full_ms / ms, dice))
r.close()

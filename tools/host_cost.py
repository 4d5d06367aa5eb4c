#!/usr/bin/env python
"""Host cost of one end-to-end frame through the Python harness + C-ABI: tiger.svg at 512^2 (the GPU needs ~15 us for it, the
host calls are the same as at 4096^2: the same 270 KB of segments and metadata), 8 frames in flight, wall clock per frame,
and where the host time goes (cProfile of the submit loop). Usage: tools/host_cost.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests"), ROOT]
import pfcu  # noqa: E402
import scenes  # noqa: E402

scene, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
n_ctx, steps = 8, 4000
rs = [pfcu.Renderer(0, lut) for _ in range(n_ctx)]
for q in rs:
    q.set_scene(scene)
    for _ in range(4):
        q.draw(clear=True, upload=True)


def loop(steps):
    pending = [False] * n_ctx
    for i in range(steps):
        k = i % n_ctx
        if pending[k]:
            rs[k].wait()
        rs[k].draw(clear=True, upload=True, wait=False)
        pending[k] = True
    for k in range(n_ctx):
        if pending[k]:
            rs[k].wait()


for rep in range(3):
    t0 = time.perf_counter()
    loop(steps)
    print("tiger 512^2 e2e, %d in flight: %.1f us/frame wall" % (n_ctx, (time.perf_counter() - t0) * 1e6 / steps), flush=True)
for rep in range(3):
    wall = pfcu.stream_frames(rs, steps)[0]
    print("the same through the C++ loop (pfhost_stream_frames): %.1f us/frame wall" % (wall * 1e6 / steps), flush=True)
pr = cProfile.Profile()
pr.enable()
loop(1000)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)

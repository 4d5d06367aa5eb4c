#!/usr/bin/env python
"""End-to-end frames/s through the C-ABI with host buffers as a function of the frames kept in flight
(pfcu_submit_frame / pfcu_wait_frame, one renderer context per frame in flight). Usage: tools/e2e_stream.py [workload]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests"), ROOT]
import bench  # noqa: E402
import pfcu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tiger4096"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
scene = bench.load_workload(name)[0]
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
for n_ctx in (1, 2, 3, 4, 6, 8):
    rs = [pfcu.Renderer(0, lut) for _ in range(n_ctx)]
    for q in rs:
        q.set_scene(scene)
        for _ in range(4):
            q.draw(clear=True, upload=True)
    best = None
    for rep in range(3):
        pending = [False] * n_ctx
        t0 = time.perf_counter()
        for i in range(steps):
            k = i % n_ctx
            if pending[k]:
                rs[k].wait()
            rs[k].draw(clear=True, upload=True, wait=False)
            pending[k] = True
        for k in range(n_ctx):
            if pending[k]:
                rs[k].wait()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        best = ms if best is None else min(best, ms)
    cpp = min(pfcu.stream_frames(rs, steps)[0] for _ in range(3)) * 1e6 / steps
    print("%s e2e, %d frame(s) in flight: %.1f us/frame (Python submit loop), %.1f us/frame (C++ loop, pfhost_stream_frames)" %
          (name, n_ctx, best * 1e3, cpp), flush=True)
    for q in rs:
        q.close()

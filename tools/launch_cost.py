#!/usr/bin/env python
"""Host cost of a frame-graph launch vs GPU time: is the multi-context throughput host- or GPU-bound?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests"), ROOT]
import bench, pfcu
scene = bench.load_workload(sys.argv[1] if len(sys.argv) > 1 else "tiger4096")[0]
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
n_ctx, frames = 4, 480
rs = []
for _ in range(n_ctx):
    q = pfcu.Renderer(0, lut); q.set_scene(scene); q.draw(); q.draw(); q.graph_capture(); q.graph_launch(); q.graph_finish(); rs.append(q)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(frames):
        rs[i % n_ctx].graph_launch()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("enqueue %.1f us/frame, total %.1f us/frame" % ((t1 - t0) / frames * 1e6, (t2 - t0) / frames * 1e6), flush=True)
# raw ctypes call cost with nothing to do
L, h = rs[0].L, rs[0].h
t0 = time.perf_counter()
for i in range(20000):
    L.pfcu_get_stream(h)
print("ctypes call: %.2f us" % ((time.perf_counter() - t0) / 20000 * 1e6))
for q in rs: q.graph_finish(); q.close()

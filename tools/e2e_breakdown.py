#!/usr/bin/env python
"""Where the end-to-end frame time (host buffers through the C-ABI) goes: wall-clock per API phase, retained frame graph
on and off. python tools/e2e_breakdown.py [fixture]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import pfcu  # noqa: E402
import scenes  # noqa: E402

fixture = sys.argv[1] if len(sys.argv) > 1 else "tiger_4096_scene"
scene, _ = scenes.load_scene(scenes.golden_path(fixture))
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
for retain in (False, True):
    r = pfcu.Renderer(0, lut)
    r.set_retain_frame_graph(retain)
    r.set_scene(scene)
    for _ in range(5):
        r.draw(clear=True, upload=True)
    L = r.L
    d = r._descs["draw"][0]
    cc = np.zeros(4, "<f4")
    acc = np.zeros(4)
    n = 200
    t_all = time.perf_counter()
    for _ in range(n):
        t0 = time.perf_counter()
        r.upload_segments(scene)
        t1 = time.perf_counter()
        L.pfcu_begin_frame(r.h)
        L.pfcu_prepare_batch(r.h, C.byref(d))
        t2 = time.perf_counter()
        L.pfcu_draw_batch(r.h, d.batch_id, -1, -1, 0, 1, cc.ctypes.data_as(C.c_void_p))
        t3 = time.perf_counter()
        st = pfcu.FrameStats()
        L.pfcu_end_frame(r.h, C.byref(st))
        t4 = time.perf_counter()
        acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
    tot = (time.perf_counter() - t_all) / n * 1e6
    print("retain=%d: %.1f us/frame | upload %.1f  begin+prepare %.1f  draw %.1f  end_frame %.1f | gpu_ms %.4f launches %d" % (
        (retain, tot) + tuple(acc / n * 1e6) + (st.gpu_ms, st.kernel_launches)))
    r.close()

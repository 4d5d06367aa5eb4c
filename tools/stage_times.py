#!/usr/bin/env python
"""Per-stage CUDA-event times (median of N frames, L2 flushed between frames) + whole-frame graph time, per library variant.
Quick A/B tool for kernel variants: PFCU_LIB=... python tools/stage_times.py [fixture]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import pfcu  # noqa: E402
import scenes  # noqa: E402

fixture = sys.argv[1] if len(sys.argv) > 1 else "tiger_4096_scene"
scene, _ = scenes.load_scene(scenes.golden_path(fixture))
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for ordered in [int(x) for x in os.environ.get("ORDER", "4,0").split(",")]:
    fused = False
    r = pfcu.Renderer(0, lut)
    try:
        r.set_order_tile_groups(ordered)
    except pfcu.PfcuError:
        pass  # (PFCU_LIB names a build from before the option)
    r.set_stream(stream.cuda_stream)
    r.set_scene(scene)
    r.draw(clear=True)
    r.draw(clear=True)
    r.set_profiling(True)
    samples = []
    for _ in range(15):
        with torch.cuda.stream(stream):
            flush.zero_()
        r.draw(clear=True)
        samples.append(r.stage_times())
    r.set_profiling(False)
    med = {k: float(np.median([s[k] for s in samples])) * 1e3 for k in samples[0]}
    r.draw(clear=True)
    r.graph_capture()
    ts = []
    for i in range(60):
        with torch.cuda.stream(stream):
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r.graph_launch()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    r.graph_finish()
    print("%s %s graph %.1f us (min %.1f) | " % (os.environ.get("PFCU_LIB", "default").split("/")[-1], "order>=%d" % ordered,
                                              float(np.median(ts[10:])), min(ts[10:])) +
          " ".join("%s %.1f" % (k, v) for k, v in med.items()))
    r.close()

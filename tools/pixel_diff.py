#!/usr/bin/env python
"""Histogram of |CUDA - oracle| per channel for every golden scene, (tolerance evidence)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pathfinder-cpp_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import pfcu  # noqa: E402
import pforacle  # noqa: E402
import scenes  # noqa: E402

lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
r = pfcu.Renderer(0, lut)
for name in sys.argv[1:] or ["tiger_512", "tiger_1024", "features_2048", "demo_clip_512"]:
    scene, _ = scenes.load_scene(scenes.golden_path(name))
    fr = pforacle.Frame(scene, lut)
    want = fr.render().astype(int)
    for fused in (False,):
        r.set_scene(scene)
        r.draw(clear=True)
        d = np.abs(r.pixels().astype(int) - want)
        print(name, "fused" if fused else "split", "hist", np.bincount(d.reshape(-1), minlength=4)[:6].tolist())
    fr.close()
r.close()

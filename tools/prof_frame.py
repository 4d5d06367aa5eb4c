#!/usr/bin/env python
"""Renders N frames of a workload through the C-ABI with no torch in the process: the command ncu wraps
(B200_PROFILING.md). Every kernel of a frame is a separate launch (no graph), so `-s <skip> -c <n>` selects frames."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import pfcu  # noqa: E402
import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fixture", default="tiger_4096_scene")
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--fused", action="store_true", help="PFCU_OPT_FUSED_FILL = 1: the tile kernel rasterizes the masks itself")
ap.add_argument("--grid-order", action="store_true", help="PFCU_OPT_ORDER_TILE_GROUPS = 0")
args = ap.parse_args()
scene, _ = scenes.load_scene(scenes.golden_path(args.fixture))
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
r = pfcu.Renderer(0, lut)
if args.fused:
    r.set_fused_fill(True)
if args.grid_order:
    r.set_order_tile_groups(0)
r.set_scene(scene)
for i in range(args.frames):
    st = r.draw(clear=True)
    print(i, st)
r.close()

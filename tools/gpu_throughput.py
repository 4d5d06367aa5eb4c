#!/usr/bin/env python
"""Frames/s with inputs resident: every context replays its captured frame graph, n contexts (streams) side by side.
Timed with CUDA events on a fork/join stream. Usage: tools/gpu_throughput.py [workload] [frames]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests"), ROOT]
import bench  # noqa: E402
import pfcu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tiger4096"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 480
scene = bench.load_workload(name)[0]
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
main = torch.cuda.Stream()
for n_ctx in [int(x) for x in os.environ.get("CONTEXTS", "1,2,3,4,6,8").split(",")]:
    rs, streams = [], []
    for _ in range(n_ctx):
        q = pfcu.Renderer(0, lut)
        st = torch.cuda.Stream()
        q.set_stream(st.cuda_stream)
        if "CONCURRENT" in os.environ:
            q.set_concurrent_batches(int(os.environ["CONCURRENT"]))
        if "ORDER" in os.environ:
            try:
                q.set_order_tile_groups(int(os.environ["ORDER"]))
            except pfcu.PfcuError:
                pass  # (a library built before the option existed)
        q.set_scene(scene)
        q.draw(clear=True)
        q.draw(clear=True)
        q.graph_capture()
        q.graph_launch()
        q.graph_finish()
        rs.append(q)
        streams.append(st)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fork = torch.cuda.Event()
        e0.record(main)
        fork.record(main)
        for st in streams:
            st.wait_event(fork)
        for i in range(frames):
            rs[i % n_ctx].graph_launch()
        for st in streams:
            j = torch.cuda.Event()
            j.record(st)
            main.wait_event(j)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / frames
        best = ms if best is None else min(best, ms)
    for q in rs:
        assert q.graph_finish()["retries"] == 0
    print("%s resident, %d context(s): %.1f us/frame" % (name, n_ctx, best * 1e3), flush=True)
    for q in rs:
        q.close()

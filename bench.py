#!/usr/bin/env python
"""bench.py -- the headline measurement of BASELINE.json: dice -> composite of tiger.svg at 4096 x 4096.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one JSON line on rank 0)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU path (hybrid SceneBuilderD3D9; one
                                                           unmodified 4-thread builder per 4 host cores, side by side)

A step is one pass of the hot path over one BATCH of --frames-per-step frames (default 256; a frame = every batch of the
scene: bound, dice, bin, propagate, sort, fill, tile). A single 4096^2 frame takes 0.06 ms, so K single-frame steps would
make a timed region of a millisecond; with batches the driver's 20 steps are 5120 frames, a third of a second.
  value   : segments/s with all inputs resident in HBM: K x frames-per-step frames, each one CUDA graph launch,
            --frames-in-flight of them in flight on as many contexts (streams); one CUDA-event pair around the K steps,
            max over ranks. config.latency_ms_per_frame: one frame at a time, L2 flushed before each (CUDA events per
            frame) -- the single-frame latency.
  e2e     : the same metric through the public C-ABI with HOST buffers: every frame uploads the scene's segments and
            batch metadata (pinned staging -> H2D), runs, and reads its counters back (D2H); wall clock over all steps
            with --e2e-contexts frames in flight (pfcu_submit_frame / pfcu_wait_frame, one context per frame in flight),
            the submit loop in C++ (host/frame_streamer.cpp; e2e.python_loop_ms_per_frame: the same calls from Python);
            e2e.serial_ms_per_frame is the blocking one-context figure (pfcu_end_frame every frame);
            e2e.with_pixels: the same with the 64 MiB frame read back into page-locked host memory as well
            (pfcu_read_target_async on a copy stream, the next contexts render meanwhile): PCIe-bound.
  roofline: the composite ("tile") kernel, algorithmic bytes of SURVEY.md section 8d / measured HBM copy bandwidth.
  sharded : the two configurations that actually shard (BASELINE.json configs 4 and 5), run in the same process and
            embedded under this key: `batch` (4096 independent tiger 512^2 frames, scene-sharded, strong scaling) and
            `strips` (200 k synthetic paths at 8192^2, one horizontal strip per rank, assembled on rank 0; strong
            scaling). Both carry "verified": the N-GPU result was compared byte for byte with the 1-GPU render.
N > 1 (torchrun): the top-level line is the scene-per-rank configuration -- every rank renders its own frames, no
data-path collective, weak scaling; the sharded configurations are under `sharded`.

Other workloads (not what the driver runs; evidence for BASELINE.json configs 4 and 5, see profiles/):
  --workload synthetic --paths 200000 --size 8192   config 4: one large canvas, horizontal strips, one per rank, assembled
                                                    by one NCCL all-gather (--gather nccl) or by the tile kernel storing
                                                    straight into rank 0's framebuffer over NVLink (--gather p2p); strong
  --workload tiger512 --frames 4096                 config 5: a batch of independent frames, scene-sharded; frames/s
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import scenes  # noqa: E402

WORKLOADS = {
    # name: (fixture, asset, size, native size)
    "tiger4096": ("tiger_4096_scene", "tiger.svg", 4096, 900.0),
    "tiger1024": ("tiger_1024", "tiger.svg", 1024, 900.0),
    "tiger512": ("tiger_512", "tiger.svg", 512, 900.0),
    "features2048": ("features_2048", "features.svg", 2048, 720.0),
    # the primitives scene of the reference demo (clip, blurred shadow, image, gradient, render-target pattern): the
    # other half of BASELINE.json configs[1]; 9 prepared batches, 11 tile passes, 4 of them into render targets
    "demo2048": ("demo_full_2048", "demo:sea.png", 2048, 720.0),
}
METRIC = "segments/s (dice->composite), %s at %dx%d"  # default workload: tiger.svg at 4096x4096
UNIT = "segments/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_workload(name):
    fixture, asset, size, native = WORKLOADS[name]
    scene, _ = scenes.load_scene(scenes.golden_path(fixture))
    return scene, asset, size, native


def n_segments(scene):
    return int(sum(int(b["info"][3]) for b in scene["draw_batches"] + scene["clip_batches"]))


# ---------------------------------------------------------------------------------------------- reference arm

def reference_cpu(asset, size, native, steps, warmup, budget_s=20.0):
    """The reference's own CPU implementation of the path: SceneBuilderD3D9::build (hybrid tiler), as shipped
    (4 worker threads, core/d3d9/scene_builder.cpp:13). Falls back to the C restatement when libpfref.so is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pfref

    if pfref.available():
        def make():
            if asset.startswith("demo:"):
                return pfref.RefScene.demo(size, size, size / native, pfref.asset(asset[5:]), 0x3f)
            return pfref.RefScene.from_svg(pfref.asset(asset), size, size, size / native)

        # The reference's builder runs 4 worker threads, hard-coded (core/d3d9/scene_builder.cpp:13). To give it every host
        # core, independent frames are built side by side, one unmodified 4-thread builder per 4 cores -- the CPU
        # counterpart of our frames in flight.
        try:
            host_cores = len(os.sched_getaffinity(0))  # the cores this process may run on
        except AttributeError:
            host_cores = os.cpu_count() or 4
        builders = max(1, min(host_cores // 4, 64))
        handles = [make() for _ in range(builders)]
        handles[0].time_d3d9_build(max(warmup, 1))
        probe = float(np.median(handles[0].time_d3d9_build(3)))
        n = int(max(1, min(steps, budget_s * 1000.0 / max(probe, 1e-3))))
        single = handles[0].time_d3d9_build(n)
        results = [None] * builders

        def work(k):
            results[k] = handles[k].time_d3d9_build(n)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(builders)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        wall_ms = (time.perf_counter() - t0) * 1e3
        for hnd in handles:
            hnd.close()
        per_frame = wall_ms / (builders * n)
        return dict(ms=np.full(n, per_frame), kind="reference", cores=4 * builders, steps=n,
                    single_builder_ms=float(np.median(single)), builders=builders,
                    sample="%d x %d frames of SceneBuilderD3D9::build, %d builders of 4 threads side by side (tiling only; the "
                           "reference has no CPU rasteriser), %s @ %d^2" % (builders, n, builders, asset, size))
    import pforacle

    scene, _ = scenes.load_scene(scenes.golden_path(WORKLOADS["tiger4096"][0]))
    fr = pforacle.Frame(scene, None)
    ts = []
    for i in range(warmup + steps):
        lib = pforacle.lib()
        lib.pfo_frame_reset(fr.h)
        t0 = time.perf_counter()
        fr.prepare_all(geometry_only=True)
        ts.append((time.perf_counter() - t0) * 1e3)
    fr.close()
    return dict(ms=np.array(ts[warmup:]), kind="port", cores=1, steps=steps,
                sample="%d x oracle/pf_oracle.c geometry (dice+bin+propagate), %s @ %d^2" % (steps, asset, size))


def run_reference(args, rank):
    if rank != 0:
        return
    scene, asset, size, native = load_workload(args.workload)
    segs = n_segments(scene)
    # a step of our arm is a batch of --frames-per-step frames; the CPU times a bounded sample of those frames (20 s of
    # host work at most) and the step time is that rate times the batch size
    r = reference_cpu(asset, size, native, args.steps * args.frames_per_step, args.warmup)
    ms = float(np.mean(r["ms"]))
    value = segs / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC % (asset, size, size), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * args.frames_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "reference asset %s parsed by the reference front end" % asset,
        "config": {"workload": "%s@%dx%d" % (asset, size, size), "frames_per_step": args.frames_per_step},
        "details": {"threads": r["cores"], "frames_timed": r["steps"], "ms_per_frame": ms,
                    "note": "CPU tiling only (no pixels): the reference has no CPU rasteriser"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "single_builder_ms_per_frame": r.get("single_builder_ms"), "builders": r.get("builders", 1)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- our arm

def measured_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from THIS build's ncu --set full capture
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep of tools/gpu_final.sh; tiger 4096 only:
    that is the workload the capture is taken on). DRAM counters cannot be read from inside the process, so a build without
    a capture reports None rather than an older build's number."""
    if workload != "tiger4096":
        return None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fp:
            doc = json.load(fp)
        t = doc["kernels"][kernel]
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), doc.get("source")
    except Exception:
        return None, None


def algorithmic_bytes(r, scene):
    """B_fill and B_tile from the frame's own unit counts (read through the parity taps)."""
    F = A = Ac = L = La = 0
    T = ((scene["width"] + 15) // 16) * ((scene["height"] + 15) // 16)
    paints = set()
    for b in scene["draw_batches"]:
        bid = int(b["info"][0])
        tiles = r.tiles(bid)
        off, lst = r.tile_lists(bid)
        # masks the fill stage rasterizes: tiles with fills whose tile survives the z-cull (PFCU_OPT_FILL_CULLED_TILES = 0)
        kept = np.zeros(len(tiles), bool)
        kept[lst] = True
        own = (tiles["alpha_tile_id"] >= 0) & (tiles["fill_count"] > 0) & kept
        F += int(tiles["fill_count"][own].sum())
        A += int(own.sum())
        Ac += int((tiles["clip_alpha_tile_id"] >= 0).sum())
        L += len(lst)
        La += int((tiles["alpha_tile_id"][lst] >= 0).sum())
        paints.update(int(c) for c in np.unique(b["tile_path_info"]["color"]))
    G = sum(int(px.size) for px in scene.get("pages", {}).values())
    P = len(paints)
    b_fill = 12 * F + 24 * A + 256 * A + 256 * Ac
    b_tile = 4 * T + 16 * L + 256 * La + 80 * P + G + 1024 * T
    return dict(F=F, A=A, A_c=Ac, L=L, L_a=La, T=T, P=P, G=G, B_fill=b_fill, B_tile=b_tile)


def pin_rank_to_cores(local, world):
    """One disjoint slice of the host's cores per rank: eight Python submit loops (each with CUDA's own helper threads) on
    shared cores is what made the e2e leg scale 0.70 at N = 4 in round 1."""
    if world <= 1:
        return None
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per >= 2:
            mine = cores[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            return len(mine)
    except (AttributeError, OSError):
        pass
    return None


def run_ours(args, rank, world, local):
    import torch

    import pfcu

    if world > 1:
        import torch.distributed as dist
    scene, asset, size, native = load_workload(args.workload)
    segs = n_segments(scene)
    lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
    stream = torch.cuda.Stream()
    r = pfcu.Renderer(local, lut)
    r.set_stream(stream.cuda_stream)
    r.set_scene(scene)
    first = r.draw(clear=True)      # sizes every buffer (may retry once)
    steady = r.draw(clear=True)     # steady state: no allocation, no retry
    assert steady["retries"] == 0, steady
    bytes_ = algorithmic_bytes(r, scene)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    frames = args.steps * args.frames_per_step  # frames of the timed region

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.zero_()

    # ---- per-kernel times (events around every kernel), for the roofline of the dominant kernel
    r.set_profiling(True)
    stage_samples = []
    for _ in range(max(33, args.warmup)):
        flush_l2()
        r.draw(clear=True)
        stage_samples.append(r.stage_times())
    r.set_profiling(False)
    # (CUDA event times come in steps of about 1 us: the mean of 30 frames, not the median, which flips between two steps)
    stage_ms = {k: float(np.mean(sorted(s[k] for s in stage_samples[3:])[2:-2])) for k in stage_samples[0]}

    # ---- single-frame latency: whole frame as one CUDA graph, one at a time, L2 flushed before each
    r.draw(clear=True)
    r.graph_capture()
    for _ in range(max(args.warmup, 3)):
        flush_l2()
        r.graph_launch()
    gstats = r.graph_finish()
    evs = []
    for _ in range(min(frames, 200)):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r.graph_launch()
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    latency_ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    gstats = r.graph_finish()

    # ---- value: K steps of frames_per_step frames as a job, args.frames_in_flight of them in flight (one context + stream
    # + captured frame graph each): a single frame is a chain of ten short kernels that leaves most of the GPU idle,
    # independent frames fill it. One CUDA-event pair on a fork/join stream around all K steps.
    n_fly = max(1, args.frames_in_flight)
    fly = [(r, stream)]
    for _ in range(n_fly - 1):
        q, qs = pfcu.Renderer(local, lut), torch.cuda.Stream()
        q.set_stream(qs.cuda_stream)
        q.set_scene(scene)
        q.draw(clear=True)
        q.draw(clear=True)
        q.graph_capture()
        fly.append((q, qs))
    main = torch.cuda.Stream()

    def run_job(k):
        fork = torch.cuda.Event()
        fork.record(main)
        for _, qs in fly:
            qs.wait_event(fork)
        for i in range(k):
            fly[i % n_fly][0].graph_launch()
        for _, qs in fly:
            j = torch.cuda.Event()
            j.record(qs)
            main.wait_event(j)

    run_job(max(args.warmup, 3) * min(args.frames_per_step, 32))
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    j0.record(main)
    run_job(frames)
    j1.record(main)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = j0.elapsed_time(j1)
    for q, _ in fly:
        assert q.graph_finish()["retries"] == 0
    for q, _ in fly[1:]:
        q.close()

    # ---- e2e: public C-ABI with host buffers, H2D + frame + D2H counters every frame (wall clock, synchronised)
    for _ in range(3):
        r.draw(clear=True, upload=True)
    torch.cuda.synchronize()
    n_serial = min(frames, 400)
    t0 = time.perf_counter()
    for _ in range(n_serial):
        st = r.draw(clear=True, upload=True)  # pfcu_end_frame synchronises and reads the counters back
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3 / n_serial
    # ... and streamed: the same per-frame work (H2D of the frame's inputs, the frame, D2H of its counters) with
    # args.e2e_contexts frames in flight, one renderer context (own stream, own framebuffer) each -- double buffering
    # as an application that renders a sequence of frames does it (pfcu_submit_frame / pfcu_wait_frame)
    n_ctx = max(1, args.e2e_contexts)
    rs = [pfcu.Renderer(local, lut) for _ in range(n_ctx)]
    for q in rs:
        q.set_scene(scene)
        for _ in range(4):
            q.draw(clear=True, upload=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # the submit loop is the application's: C++ (pathfinder-cpp_b200/host/frame_streamer.cpp, pfhost_stream_frames), like the
    # reference's own applications; the same loop through the Python harness is timed beside it (e2e.python_loop_ms_per_frame)
    pfcu.stream_frames(rs, 4 * n_ctx)
    wall, st, retries = pfcu.stream_frames(rs, frames)
    e2e_ms = wall * 1e3
    assert retries == 0 and st["fills"] == steady["fills"], (retries, st)
    n_py = min(frames, 2000)
    pending = [False] * n_ctx
    t0 = time.perf_counter()
    for i in range(n_py):
        k = i % n_ctx
        if pending[k]:
            rs[k].wait()
        rs[k].draw(clear=True, upload=True, wait=False)
        pending[k] = True
    for k in range(n_ctx):
        if pending[k]:
            st = rs[k].wait()
    e2e_py_ms = (time.perf_counter() - t0) * 1e3 / n_py
    assert st["retries"] == 0 and st["fills"] == steady["fills"], st
    clocks = sampler.stop()
    h2d = int(sum(scene[k].nbytes for k in ("draw_points", "draw_indices", "clip_points", "clip_indices")))
    for b in scene["draw_batches"] + scene["clip_batches"]:
        h2d += int(sum(b[k].nbytes for k in ("backdrops", "propagate_metadata", "dice_metadata", "tile_path_info")))
    d2h = 64 * (len(scene["draw_batches"]) + len(scene["clip_batches"]) + 1)
    # ... and with the pixels: every frame's target is also read back into page-locked host memory
    # (pfcu_read_target_async: copy stream, behind the submitted frame); a bounded number of frames, PCIe-bound
    frame_bytes = int(scene["width"]) * int(scene["height"]) * 4
    n_px = min(frames, 96)
    bufs = [q.pinned_frame() for q in rs]
    for k in range(n_ctx):
        rs[k].draw(clear=True, upload=True, wait=False)
        rs[k].read_async(bufs[k])
        rs[k].wait()
        rs[k].wait_read()
    wall, st, retries = pfcu.stream_frames(rs, n_px, pixels=bufs)
    px_ms = wall * 1e3 / n_px
    assert retries == 0, retries
    px_ok = bool(bufs[0][..., 3].any())
    # the round-1 way, for comparison: blocking cudaMemcpy2D into pageable memory after a blocking frame
    t0 = time.perf_counter()
    r.draw(clear=True, upload=True)
    r.pixels()
    px_blocking_ms = (time.perf_counter() - t0) * 1e3
    for q in rs:
        q.close()

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, e2e_serial_ms, latency_ms, px_ms, e2e_py_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, e2e_serial_ms, latency_ms, px_ms, e2e_py_ms = (float(x) for x in t)
    r.close()
    if rank != 0:
        return None
    value = world * frames * segs / (total_ms / 1e3)
    e2e_value = world * frames * segs / (e2e_ms / 1e3)
    peak, peak_src = measured_peak()
    comp_ms = stage_ms["composite"]
    achieved = bytes_["B_tile"] / (comp_ms * 1e-3) / 1e9 if comp_ms > 0 else 0.0
    fill_gbs = bytes_["B_fill"] / (stage_ms["fill"] * 1e-3) / 1e9 if stage_ms["fill"] > 0 else 0.0
    both = (bytes_["B_fill"] + bytes_["B_tile"]) / ((stage_ms["fill"] + comp_ms) * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic("k_composite<1>", args.workload)
    line = {
        "metric": METRIC % (asset, size, size), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic-free: reference asset %s, scene fixture built by the reference front end"
                                % asset,
        # (config is the same dict in both arms; everything else about this run is under details)
        "config": {"workload": "%s@%dx%d" % (asset, size, size), "frames_per_step": args.frames_per_step},
        "details": {"sharding": "scene-per-rank, no collective",
                   "timing": "one CUDA-event pair around all K steps (K x %d frames) on a fork/join stream; a frame = 1 "
                             "CUDA graph launch; %d frames in flight on %d contexts (streams)" % (args.frames_per_step, n_fly, n_fly),
                   "l2": "no flush: the contexts in flight rotate over %d x (%d MiB framebuffer + intermediates) > L2; "
                         "latency_ms_per_frame is measured with a 256 MiB memset between frames (not timed)"
                         % (n_fly, frame_bytes >> 20),
                   "frames_in_flight": n_fly,
                   "timed_region_s": total_ms / 1e3,
                   "ms_per_frame": total_ms / frames,  # reciprocal throughput with frames in flight
                   # one frame at a time, CUDA events per frame on its stream, L2 flushed before every frame
                   "latency_ms_per_frame": latency_ms,
                   "frames_per_s": world * frames / (total_ms / 1e3),
                   "units": {k: gstats[k] for k in ("segments", "lines", "fills", "alpha_tiles", "dense_tiles",
                                                    "listed_tiles", "listed_after_cull", "max_list_len", "fb_tiles")}},
        "roofline": {"bound": "hbm", "kernel": "k_composite (tile)", "achieved": achieved, "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": bytes_["B_tile"], "kernel_ms": comp_ms,
                     "fill": {"achieved": fill_gbs, "frac": fill_gbs / peak, "algorithmic_bytes": bytes_["B_fill"],
                              "kernel_ms": stage_ms["fill"]},
                     "fill_plus_tile": {"achieved": both, "frac": both / peak},
                     "stage_ms": stage_ms, "counts": bytes_},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_frame": e2e_ms / frames, "h2d_bytes_per_step": h2d * args.frames_per_step,
                "d2h_bytes_per_step": d2h * args.frames_per_step, "h2d_bytes_per_frame": h2d, "d2h_bytes_per_frame": d2h,
                "frames_in_flight": n_ctx, "result": "RGBA8 frame stays on the device (the reference renders into a device "
                                                    "texture too); counters read back",
                "serial_ms_per_frame": e2e_serial_ms,  # one context, pfcu_end_frame blocks every frame
                "submit_loop": "C++ (pfhost_stream_frames, pathfinder-cpp_b200/host/frame_streamer.cpp) over the C-ABI",
                "python_loop_ms_per_frame": e2e_py_ms,  # the same calls issued by the Python harness
                "with_pixels": {"value": world * segs / (px_ms / 1e3), "unit": UNIT, "ms_per_frame": px_ms, "frames": n_px,
                                "d2h_bytes_per_frame": d2h + frame_bytes, "gb_per_s": frame_bytes / (px_ms * 1e-3) / 1e9,
                                "how": "pfcu_read_target_async into page-locked memory behind every submitted frame, "
                                       "%d contexts in flight" % n_ctx, "nonzero": px_ok,
                                "blocking_pageable_ms_per_frame": px_blocking_ms}},
        "gpu_launches": int(gstats["kernel_launches"]) * frames,
        "clocks": clocks,
        "first_frame": {"retries": first["retries"], "gpu_ms": first["gpu_ms"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        c = reference_cpu(asset, size, native, 400, 3, budget_s=12.0)
        ms = float(np.median(c["ms"]))
        line["cpu_baseline"] = {"value": segs / (ms / 1e3), "unit": UNIT, "cores": c["cores"], "kind": c["kind"],
                                "sample": c["sample"], "ms_per_frame": ms, "host_cpus": os.cpu_count(),
                                "single_builder_ms_per_frame": c.get("single_builder_ms"), "builders": c.get("builders", 1),
                                "scope": "CPU tiling only (SceneBuilderD3D9::build: no pixels); the GPU arm renders the pixels "
                                         "too -- a reported baseline, not like for like"}
    return line


def run_strips(args, rank, world, local, steps, warmup, gather):
    """BASELINE.json config 4: synthetic cubic blobs on one size x size canvas, strip k on rank k (sharding.py).
    Returns the JSON line (rank 0) or None. The assembled frame is compared with a 1-GPU render of the whole canvas."""
    import torch
    import torch.distributed as dist

    import pfcu
    import sharding

    size = args.size
    paths, colors = scenes.synthetic_paths(args.paths, size)
    y0, y1 = sharding.strip_bounds(size, world, rank)
    rows = sharding.strip_rows(size, world)
    scene = scenes.build_scene_from_outlines(size, size, paths, colors, strip=(y0, y1))
    # every on-curve point starts one segment (SegmentsD3D11::add_path closes each contour, gpu_data.cpp:109)
    segs_total = int(sum(int((c[1] == 0).sum()) for p_ in paths for c in p_["contours"]))
    lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
    dev = torch.device("cuda", local)
    # frames in flight (p2p gather only): frame k + 1 is computed while frame k's strips are still arriving at rank 0
    # (7/8 of the canvas through one GPU's NVLink ingress); every frame in flight has its own context, stream and
    # presenting framebuffer
    peer_modes = ("p2p", "copy")
    sweep = [int(x) for x in str(args.strip_sweep).split(",") if x] if (args.strip_sweep and gather in peer_modes and world > 1) else []
    n_fly = max([max(1, args.strip_frames_in_flight)] + sweep) if (gather in peer_modes and world > 1) else 1
    lanes = []
    full = None
    for k in range(n_fly):
        stream = torch.cuda.Stream()
        peer = None
        local_strip = None
        if gather == "p2p" and world > 1:
            peer = sharding.PeerFramebuffer(size, size, world, rank, dev)
            target_ptr = peer.strip_ptr()
        elif gather == "copy" and world > 1:
            # render into a local strip; a copy engine pushes it into rank 0's framebuffer (rank 0 renders in place)
            peer = sharding.PeerFramebuffer(size, size, world, rank, dev)
            if rank == 0:
                target_ptr = peer.strip_ptr()
            else:
                local_strip = torch.zeros((rows, size, 4), dtype=torch.uint8, device=dev)
                target_ptr = local_strip.data_ptr()
        else:
            full = torch.zeros((rows * world, size, 4), dtype=torch.uint8, device=dev)
            target_ptr = full[rank * rows:].data_ptr()
        r = pfcu.Renderer(local, lut)
        r.set_stream(stream.cuda_stream)
        r.set_scene(scene, target_ptr, size * 4)
        first = r.draw(clear=True)
        steady = r.draw(clear=True)
        assert steady["retries"] == 0, steady
        if k == 0:
            r.set_profiling(True)
            r.draw(clear=True)
            stage_ms = r.stage_times()
            r.set_profiling(False)
            r.draw(clear=True)
        r.graph_capture()
        lanes.append((r, stream, peer, local_strip))
    r, stream, peer, _ = lanes[0]
    main = torch.cuda.Stream()
    counter = [0]
    all_lanes = lanes
    n_max = n_fly

    def timed(k_fly):
        """ms per frame with k_fly frames in flight (the first k_fly lanes)."""
        nonlocal lanes, n_fly
        lanes, n_fly = all_lanes[:k_fly], k_fly
        counter[0] = 0
        for _ in range(max(warmup, 3)):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        fork()
        for _ in range(steps):
            step()
        join()
        b.record(main)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tt = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]) / steps

    def step():
        q, qs, qp, qlocal = lanes[counter[0] % n_fly]
        counter[0] += 1
        q.graph_launch()
        if world > 1:
            if qp is not None:
                if qlocal is not None:
                    qp.push_strip(qlocal, qs)
                with torch.cuda.stream(qs):
                    qp.barrier()
            else:
                with torch.cuda.stream(qs):
                    dist.all_gather_into_tensor(full.view(-1), full[rank * rows:(rank + 1) * rows].view(-1))

    def fork():
        ev = torch.cuda.Event()
        ev.record(main)
        for _, qs, _, _ in lanes:
            qs.wait_event(ev)

    def join():
        for _, qs, _, _ in lanes:
            ev = torch.cuda.Event()
            ev.record(qs)
            main.wait_event(ev)

    sweep_ms = {}
    for k_fly in sweep:  # (diagnostic: the same frames with other numbers of frames in flight)
        sweep_ms[str(k_fly)] = timed(k_fly)
    lanes, n_fly = all_lanes[:max(1, args.strip_frames_in_flight) if (gather in peer_modes and world > 1) else 1], \
        (max(1, args.strip_frames_in_flight) if (gather in peer_modes and world > 1) else 1)
    counter[0] = 0
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    fork()
    for _ in range(steps):
        step()
    join()
    e1.record(main)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    lanes = all_lanes
    for q, _, _, _ in lanes:
        gstats = q.graph_finish()
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    units = torch.tensor([gstats[k] for k in ("segments", "lines", "fills", "alpha_tiles", "dense_tiles")],
                         device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)

    # ---- verification (not timed): the frame assembled from the N strips against ONE GPU rendering the whole canvas
    verified = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            assembled = (peer.frame(size) if peer is not None else full[:size])
            whole = scenes.build_scene_from_outlines(size, size, paths, colors)
            ref = torch.zeros((size, size, 4), dtype=torch.uint8, device=dev)
            q = pfcu.Renderer(local, lut)
            q.set_scene(whole, ref.data_ptr(), size * 4)
            q.draw(clear=True)
            q.close()
            torch.cuda.synchronize()
            verified = bool(torch.equal(assembled, ref))
            del ref
        dist.barrier()
    line = None
    if rank == 0:
        ms = float(t[0]) / steps
        line = {
            "metric": "segments/s (dice->composite), %d synthetic cubic paths at %dx%d, strip-sharded" % (args.paths, size, size),
            "value": segs_total / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (PCG32 seed 0x5EED5EED, SURVEY.md section 8d config 4)",
            "verified": verified,
            "verified_how": ("assembled %d-strip frame == 1-GPU render of the whole canvas (torch.equal, %d x %d x 4 bytes)"
                             % (world, size, size)) if world > 1 else "single GPU: nothing to assemble",
            "config": {"workload": "synthetic%d@%dx%d" % (args.paths, size, size), "sharding": "horizontal strips, one per rank",
                       "gather": ({"p2p": "tile-kernel stores into rank 0's framebuffer over NVLink (peer mapping) + barrier",
                                   "copy": "strip rendered locally, pushed into rank 0's framebuffer by a copy engine "
                                           "(cudaMemcpyAsync over the NVLink peer mapping) + barrier",
                                   "nccl": "one NCCL all-gather of %d-row blocks" % rows}[gather]) if world > 1 else "none",
                       "l2": "framebuffer (%d MiB) larger than L2" % (size * size * 4 >> 20),
                       "frames_in_flight": n_fly, "frames_in_flight_sweep_ms": sweep_ms,
                       "units_all_ranks": dict(zip(("segments", "lines", "fills", "alpha_tiles", "dense_tiles"),
                                                   (int(x) for x in units.tolist()))),
                       "rank0_stage_ms": stage_ms, "frame_ms_per_rank0": first["gpu_ms"]},
            "gpu_launches": int(gstats["kernel_launches"]) * steps, "clocks": clocks,
        }
    for q, _, _, _ in lanes:
        q.close()
    del lanes, full
    torch.cuda.empty_cache()
    return line


def run_batch(args, rank, world, local, frames, steps, warmup, workload):
    """BASELINE.json config 5: a batch of independent frames (the same small scene), scene-sharded, frames/s.
    Returns the JSON line (rank 0) or None. Every rank's last frame is compared with rank 0's (sha256)."""
    import hashlib

    import torch
    import torch.distributed as dist

    import pfcu
    import sharding

    scene, asset, size, native = load_workload(workload)
    lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
    mine = sharding.scene_share(frames, world, rank)
    # Independent frames do not wait for each other: K renderer contexts (own stream, own framebuffer, own captured
    # frame graph) take the frames round-robin, so the small kernels of several 512 x 512 frames share the GPU.
    n_ctx = max(1, min(args.contexts, len(mine) or 1))
    master = torch.cuda.Stream()
    streams = [torch.cuda.Stream() for _ in range(n_ctx)]
    rs = []
    for st in streams:
        r = pfcu.Renderer(local, lut)
        r.set_stream(st.cuda_stream)
        r.set_scene(scene)
        r.draw(clear=True)
        r.draw(clear=True)
        r.graph_capture()
        for _ in range(max(warmup, 3)):
            r.graph_launch()
        rs.append(r)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(master)
    for st in streams:
        st.wait_event(e0)
    for _ in range(steps):
        for i, _f in enumerate(mine):
            rs[i % n_ctx].graph_launch()
    for st in streams:
        done = torch.cuda.Event()
        done.record(st)
        master.wait_event(done)
    e1.record(master)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    gstats = rs[0].graph_finish()
    for r in rs[1:]:
        r.graph_finish()
    # ---- verification (not timed): the frames of every context of every rank are the same bytes as rank 0's first
    digests = set(hashlib.sha256(r.pixels().tobytes()).hexdigest() for r in rs)
    mine_ok = len(digests) == 1
    d = torch.tensor(list(bytes.fromhex(sorted(digests)[0])), device="cuda", dtype=torch.int32)
    ok = torch.tensor([1 if mine_ok else 0], device="cuda", dtype=torch.int32)
    if world > 1:
        d0 = d.clone()
        dist.broadcast(d0, 0)
        ok &= (d0 == d).all().to(torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    line = None
    if rank == 0:
        ms = float(t[0]) / steps
        line = {
            "metric": "frames/s, batch of %d independent %s renders at %dx%d, scene-sharded" % (frames, asset, size, size),
            "value": frames / (ms / 1e3), "unit": "frames/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "reference asset %s, scene fixture built by the reference front end" % asset,
            "verified": bool(int(ok[0])),
            "verified_how": "sha256 of the frame of every context on every rank == rank 0's",
            "config": {"workload": "%d x %s@%dx%d" % (frames, asset, size, size),
                       "sharding": "contiguous share of the batch per rank, no collective",
                       "contexts_per_gpu": n_ctx,
                       "l2": "working set of one frame < L2 (frames are back to back, as in a batch)",
                       "segments_per_s": frames * n_segments(scene) / (ms / 1e3)},
            "gpu_launches": int(gstats["kernel_launches"]) * steps * len(mine), "clocks": clocks}
    for r in rs:
        r.close()
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tiger4096", choices=sorted(WORKLOADS) + ["synthetic"])
    ap.add_argument("--frames-per-step", type=int, default=256, help="frames of one step (a step is a batch of frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the embedded batch / strip configurations")
    ap.add_argument("--paths", type=int, default=200000, help="synthetic workload: number of paths")
    ap.add_argument("--size", type=int, default=8192, help="synthetic workload: canvas size")
    ap.add_argument("--gather", default="copy", choices=["nccl", "p2p", "copy"], help="synthetic workload: strip assembly")
    ap.add_argument("--frames", type=int, default=0, help="batch mode: render this many independent frames per step")
    ap.add_argument("--contexts", type=int, default=16, help="batch mode: renderer contexts (streams) per GPU")
    ap.add_argument("--strip-frames-in-flight", type=int, default=4, help="synthetic workload, --gather p2p: frames in flight")
    ap.add_argument("--strip-sweep", default="", help="synthetic workload, --gather p2p: also time these numbers of frames in flight (e.g. 1,3,4)")
    ap.add_argument("--frames-in-flight", type=int, default=4, help="value leg: contexts replaying their frame graph side by side")
    ap.add_argument("--e2e-contexts", type=int, default=6, help="e2e leg: frames in flight (1 = blocking pfcu_end_frame)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        if args.workload == "synthetic":
            args.workload = "tiger4096"
        run_reference(args, rank)
        return
    import torch

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    cores = pin_rank_to_cores(local, world)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.workload == "synthetic":
        line = run_strips(args, rank, world, local, args.steps, args.warmup, args.gather)
    elif args.frames > 0:
        line = run_batch(args, rank, world, local, args.frames, args.steps, args.warmup, args.workload)
    else:
        line = run_ours(args, rank, world, local)
        if not args.no_sharded:
            # the configurations that shard (BASELINE.json configs 4 and 5), under the driver's own command and clock
            sharded = {"batch": run_batch(args, rank, world, local, 4096, 3, 1, "tiger512"),
                       "strips": run_strips(args, rank, world, local, 20, 3, "copy")}
            if world > 1:  # the other two ways of assembling the frame, for comparison
                sharded["strips_kernel_stores"] = run_strips(args, rank, world, local, 20, 3, "p2p")
                sharded["strips_nccl"] = run_strips(args, rank, world, local, 20, 3, "nccl")
            if line is not None:
                line["sharded"] = sharded
        if line is not None and cores:
            line["details"]["host_cores_per_rank"] = cores
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

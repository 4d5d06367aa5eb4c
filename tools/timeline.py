#!/usr/bin/env python
"""Device-side timeline (pfcu_set_timeline: %globaltimer stamps of every CTA) of (1) one frame alone and (2) a stream of
frames on n contexts: which kernels run when, how much of the GPU each holds, what the frames in flight do to each other.
Usage: tools/timeline.py [workload] [contexts] [frames per context]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the stamps are only compiled into the trace build of the library
os.environ.setdefault("PFCU_LIB", os.path.join(ROOT, "pathfinder-cpp_b200", "lib", "libpfcu_trace.so"))
sys.path[:0] = [os.path.join(ROOT, "pathfinder-cpp_b200"), os.path.join(ROOT, "tests"), ROOT]
import bench  # noqa: E402
import pfcu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tiger4096"
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
per_ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 24
scene = bench.load_workload(name)[0]
lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
N_SM = torch.cuda.get_device_properties(0).multi_processor_count
STAGE = {i: s for i, s in enumerate(pfcu.STAGES)}
STAGE[2 | 0x100] = "bin_long"
ORDER = ["init", "dice", "bin", "bin_long", "scan_tiles", "fill_scatter", "propagate", "scan_fb", "list_scatter", "fill", "composite"]


def make(stream):
    q = pfcu.Renderer(0, lut)
    q.set_stream(stream.cuda_stream)
    q.set_timeline(per_ctx * 16000 + 16000)
    q.set_scene(scene)
    q.draw(clear=True)
    q.draw(clear=True)
    q.graph_capture()
    q.graph_launch()
    q.graph_finish()
    q.read_timeline()
    return q


def split_frames(rec, n_frames):
    """Records of one context -> per-frame record arrays. The frames of a context are whole graph launches on one stream:
    they never overlap and each leaves the same number of records (one per CTA), so sorting by placement time and cutting
    by count separates them."""
    assert len(rec) % n_frames == 0, (len(rec), n_frames)
    rec = rec[np.argsort(rec["t_placed_ns"], kind="stable")]
    per = len(rec) // n_frames
    out = [rec[k * per : (k + 1) * per] for k in range(n_frames)]
    for a, b in zip(out, out[1:]):
        assert a["t_end_ns"].max() <= b["t_placed_ns"].min() + 2000, "frames of one context overlap"
    return out


def describe(frame, t0, label):
    print(label)
    print("  %-13s %6s %9s %9s %9s %9s %7s %8s" % ("kernel", "CTAs", "placed", "started", "ended", "busy us", "SMs", "SM-us"))
    for s in ORDER:
        r = frame[np.array([STAGE.get(int(x), "?") == s for x in frame["stage"]])]
        if not len(r):
            continue
        placed, start, end = r["t_placed_ns"].min(), r["t_start_ns"].min(), r["t_end_ns"].max()
        sm_us = float((r["t_end_ns"] - r["t_start_ns"]).sum()) * 1e-3
        print("  %-13s %6d %9.1f %9.1f %9.1f %9.1f %7d %8.0f" % (s, len(r), (placed - t0) * 1e-3, (start - t0) * 1e-3,
                                                               (end - t0) * 1e-3, (end - start) * 1e-3, len(np.unique(r["sm"])), sm_us))
    print("  frame: %.1f us from the first CTA placed to the last CTA ended" %
          ((frame["t_end_ns"].max() - frame["t_placed_ns"].min()) * 1e-3))


# ---- (1) one frame alone, L2 flushed
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
q = make(stream)
solo = []
for _ in range(8):
    with torch.cuda.stream(stream):
        flush.zero_()
    torch.cuda.synchronize()
    q.graph_launch()
    q.graph_finish()
    solo.append(q.read_timeline())
q.close()
res = np.diff(np.unique(np.concatenate([s["t_end_ns"] for s in solo])))
print("%%globaltimer resolution: smallest step %d ns, median step %d ns" % (res.min(), np.median(res)))
f = solo[-1]
if os.environ.get("TIMELINE_DUMP"):
    np.save(os.environ["TIMELINE_DUMP"], f)
describe(f, f["t_placed_ns"].min(), "== one frame of %s alone (graph launch, L2 flushed), times in us from the first CTA placed" % name)
solo_busy = {}
for s in ORDER:
    r = f[np.array([STAGE.get(int(x), "?") == s for x in f["stage"]])]
    if len(r):
        solo_busy[s] = (r["t_end_ns"].max() - r["t_start_ns"].min()) * 1e-3

# ---- (2) n contexts, frames streamed
streams = [torch.cuda.Stream() for _ in range(n_ctx)]
rs = [make(st) for st in streams]
main = torch.cuda.Stream()
fork = torch.cuda.Event()
fork.record(main)
for st in streams:
    st.wait_event(fork)
for i in range(per_ctx * n_ctx):
    rs[i % n_ctx].graph_launch()
torch.cuda.synchronize()
recs = []
for q in rs:
    q.graph_finish()
    recs.append(q.read_timeline())
    q.close()
frames = []
for c, rec in enumerate(recs):
    for k, fr in enumerate(split_frames(rec, per_ctx)):
        frames.append((c, k, fr))
# steady state: drop the first and last quarter of every context's frames
steady = [(c, k, fr) for c, k, fr in frames if per_ctx // 4 <= k < per_ctx - per_ctx // 4]
t_lo = min(fr["t_placed_ns"].min() for _, _, fr in steady)
t_hi = max(fr["t_end_ns"].max() for _, _, fr in steady)
print()
print("== %d contexts x %d frames streamed; steady-state window: %d frames in %.1f us = %.1f us per frame" %
      (n_ctx, per_ctx, len(steady), (t_hi - t_lo) * 1e-3, (t_hi - t_lo) * 1e-3 / len(steady)))
lat = [(fr["t_end_ns"].max() - fr["t_placed_ns"].min()) * 1e-3 for _, _, fr in steady]
print("  latency of a frame in the stream: median %.1f us (min %.1f, max %.1f)" % (np.median(lat), min(lat), max(lat)))
print("  %-13s %12s %12s %12s" % ("kernel", "alone us", "streamed us", "CTA-time x"))
allrec = np.concatenate([fr for _, _, fr in steady])
for s in ORDER:
    durs, cta = [], []
    for _, _, fr in steady:
        r = fr[np.array([STAGE.get(int(x), "?") == s for x in fr["stage"]])]
        if len(r):
            durs.append((r["t_end_ns"].max() - r["t_start_ns"].min()) * 1e-3)
            cta.append(float((r["t_end_ns"] - r["t_start_ns"]).mean()))
    if durs:
        r0 = f[np.array([STAGE.get(int(x), "?") == s for x in f["stage"]])]
        print("  %-13s %12.1f %12.1f %12.2f" % (s, solo_busy.get(s, 0.0), float(np.median(durs)),
                                               float(np.mean(cta)) / max(float((r0["t_end_ns"] - r0["t_start_ns"]).mean()), 1.0)))
# what the GPU is doing over the steady window, in 1 us buckets
nb = int((t_hi - t_lo) // 1000) + 1
kernels_active = np.zeros(nb, int)      # distinct (context, frame, stage) with a CTA running
sm_busy = np.zeros((N_SM, nb), bool)    # SM has a running CTA
comp_active = np.zeros(nb, int)
fill_active = np.zeros(nb, int)
for c, k, fr in steady:
    for st in np.unique(fr["stage"]):
        r = fr[fr["stage"] == st]
        a = max(int((r["t_start_ns"].min() - t_lo) // 1000), 0)
        b = min(int((r["t_end_ns"].max() - t_lo) // 1000), nb - 1)
        kernels_active[a : b + 1] += 1
        if int(st) == 9:
            comp_active[a : b + 1] += 1
        if int(st) == 8:
            fill_active[a : b + 1] += 1
    a = np.clip((fr["t_start_ns"].astype(np.int64) - int(t_lo)) // 1000, 0, nb - 1)
    b = np.clip((fr["t_end_ns"].astype(np.int64) - int(t_lo)) // 1000, 0, nb - 1)
    for sm, x, y in zip(fr["sm"], a, b):
        sm_busy[int(sm) % N_SM, x : y + 1] = True
core = slice(nb // 10, nb - nb // 10)
print("  kernels running at once (any frame): mean %.2f; tile kernels at once: mean %.2f (two or more %.0f %% of the time); fill kernels: mean %.2f" %
      (kernels_active[core].mean(), comp_active[core].mean(), 100.0 * (comp_active[core] >= 2).mean(), fill_active[core].mean()))
print("  SMs with a running CTA: mean %.1f of %d (%.0f %%); time with fewer than half of the SMs busy: %.0f %%" %
      (sm_busy[:, core].sum(0).mean(), N_SM, 100.0 * sm_busy[:, core].mean(), 100.0 * (sm_busy[:, core].sum(0) < N_SM / 2).mean()))
tot_cta_us = float((allrec["t_end_ns"] - allrec["t_start_ns"]).sum()) * 1e-3 / len(steady)
print("  CTA residency per frame (sum over CTAs of start..end): %.0f us = %.1f us per SM" % (tot_cta_us, tot_cta_us / N_SM))

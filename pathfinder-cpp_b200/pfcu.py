"""ctypes binding of the C-ABI in include/pfcu.h (lib/libpfcu.so) + a thin `Renderer` that sequences a frame
the way the reference's RendererD3D11::draw does (pathfinder/core/d3d11/renderer.cpp:302-336).

This is harness glue for tests/, bench.py and __graft_entry__.py; the product is the shared library. There is
no fallback of any kind: if the library is missing or no CUDA device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PFCU_LIB selects an experimental build of the same sources (tools/variants.sh); the default is the product library
LIB_PATH = os.environ.get("PFCU_LIB") or os.path.join(_HERE, "lib", "libpfcu.so")
NONE = 0xFFFFFFFF

EXPORTS = [
    "pfcu_abi_version", "pfcu_last_error", "pfcu_create", "pfcu_destroy", "pfcu_set_stream", "pfcu_get_stream",
    "pfcu_set_area_lut", "pfcu_set_target", "pfcu_set_target_origin", "pfcu_upload_scene", "pfcu_update_scene_range",
    "pfcu_upload_paint_metadata", "pfcu_alloc_page",
    "pfcu_upload_page_region", "pfcu_begin_frame", "pfcu_prepare_batch", "pfcu_draw_batch", "pfcu_end_frame",
    "pfcu_submit_frame", "pfcu_wait_frame", "pfcu_read_target_region",
    "pfcu_read_target_async", "pfcu_wait_read", "pfcu_host_alloc", "pfcu_host_free",
    "pfcu_stroke_to_fill", "pfcu_stroke_result", "pfcu_stroke_gpu_ms", "pfcu_dash_outlines", "pfcu_dash_result",
    "pfcu_read_target", "pfcu_read_page", "pfcu_target_device_ptr", "pfcu_read_lines", "pfcu_read_fills",
    "pfcu_read_tiles", "pfcu_read_z", "pfcu_read_tile_lists", "pfcu_read_mask", "pfcu_set_profiling",
    "pfcu_get_stage_times", "pfcu_set_option", "pfcu_graph_capture", "pfcu_graph_launch", "pfcu_graph_finish",
    "pfcu_set_timeline", "pfcu_read_timeline",
]
STAGES = ["init", "dice", "bin", "scan_tiles", "fill_scatter", "propagate", "scan_fb", "list_scatter", "fill",
          "composite"]

LINE_DT = np.dtype([("from_x", "<f4"), ("from_y", "<f4"), ("to_x", "<f4"), ("to_y", "<f4"), ("path_index", "<u4")])
FILL_DT = np.dtype([("tile_index", "<u4"), ("from_x", "<u2"), ("from_y", "<u2"), ("to_x", "<u2"), ("to_y", "<u2")])
TIMELINE_DT = np.dtype([("stage", "<u4"), ("sm", "<u4"), ("cta", "<u4"), ("n_ctas", "<u4"), ("t_placed_ns", "<u8"),
                        ("t_start_ns", "<u8"), ("t_end_ns", "<u8")])
TILE_DT = np.dtype([("alpha_tile_id", "<i4"), ("clip_alpha_tile_id", "<i4"), ("fill_count", "<i4"),
                    ("backdrop", "i1"), ("backdrop_delta", "i1"), ("backdrop_d3d9", "i1"), ("listed", "u1")])


class BatchDesc(C.Structure):
    _fields_ = [("batch_id", C.c_uint32), ("path_count", C.c_uint32), ("tile_count", C.c_uint32),
                ("segment_count", C.c_uint32), ("column_count", C.c_uint32), ("path_source", C.c_int32),
                ("clip_batch_id", C.c_int32), ("backdrops", C.c_void_p), ("propagate_metadata", C.c_void_p),
                ("dice_metadata", C.c_void_p), ("tile_path_info", C.c_void_p), ("transform", C.c_float * 6)]


class FrameStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("batches", "segments", "lines", "fills", "alpha_tiles", "dense_tiles",
                                          "listed_tiles", "listed_after_cull", "fb_tiles", "max_list_len",
                                          "overflow_flags", "retries", "kernel_launches", "diced_segments",
                                          "uploaded_bytes")] + \
               [("reserved", C.c_uint32 * 1), ("gpu_ms", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class HostDraw(C.Structure):
    _fields_ = [("batch_id", C.c_uint32), ("target_page", C.c_int32), ("color_page", C.c_int32), ("sampling_flags", C.c_uint32)]


class HostScene(C.Structure):  # pfhost_scene, host/frame_streamer.h
    _fields_ = [("points", C.c_void_p * 2), ("indices", C.c_void_p * 2), ("n_points", C.c_uint32 * 2),
                ("n_segments", C.c_uint32 * 2), ("clip_batches", C.c_void_p), ("n_clip_batches", C.c_uint32),
                ("draw_batches", C.c_void_p), ("draws", C.c_void_p), ("n_draw_batches", C.c_uint32),
                ("clear_color", C.c_float * 4)]


class PfcuError(RuntimeError):
    pass


_host = None


def host_lib():
    """lib/libpfhost.so: the application-side C++ loop over the C-ABI (host/frame_streamer.h)."""
    global _host
    if _host is None:
        lib()  # first: its symbols must be global
        H = C.CDLL(os.path.join(_HERE, "lib", "libpfhost.so"))
        H.pfhost_stream_frames.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_double),
                                           C.POINTER(FrameStats), C.POINTER(C.c_uint32)]
        _host = H
    return _host


def stream_frames(renderers, n_frames, clear_color=(0.0, 0.0, 0.0, 0.0), pixels=None):
    """pfhost_stream_frames: n_frames frames of the renderers' (common) scene, frame i on renderers[i % n], every frame with
    its uploads, its counters read back and -- pixels: one page-locked array per renderer -- its target read back.
    Returns (wall seconds, statistics of the last frame, retries over all frames)."""
    r0 = renderers[0]
    scene = r0.scene
    keep = []
    hs = HostScene()
    for which, name in ((0, "draw"), (1, "clip")):
        pts = np.ascontiguousarray(scene[name + "_points"], "<f4")
        idx = np.ascontiguousarray(scene[name + "_indices"], "<u4")
        keep += [pts, idx]
        hs.points[which], hs.indices[which] = pts.ctypes.data, idx.ctypes.data
        hs.n_points[which], hs.n_segments[which] = len(pts), len(idx)
    clips = (BatchDesc * max(1, len(r0._descs["clip"])))(*r0._descs["clip"])
    draws_d = (BatchDesc * max(1, len(r0._descs["draw"])))(*r0._descs["draw"])
    draws = (HostDraw * max(1, len(r0._descs["draw"])))()
    for i, (d, b) in enumerate(zip(r0._descs["draw"], scene["draw_batches"])):
        info = b["info"]
        draws[i].batch_id = d.batch_id
        draws[i].target_page = -1 if int(info[10]) == NONE else int(info[11])
        draws[i].color_page = -1 if int(info[7]) == NONE else int(info[7])
        draws[i].sampling_flags = 0 if int(info[8]) == NONE else int(info[8])
    hs.clip_batches, hs.n_clip_batches = C.addressof(clips), len(r0._descs["clip"])
    hs.draw_batches, hs.draws, hs.n_draw_batches = C.addressof(draws_d), C.addressof(draws), len(r0._descs["draw"])
    for i in range(4):
        hs.clear_color[i] = float(clear_color[i])
    ctxs = (C.c_void_p * len(renderers))(*[q.h for q in renderers])
    px = None
    if pixels is not None:
        px = (C.c_void_p * len(renderers))(*[p.ctypes.data for p in pixels])
    wall, st, retries = C.c_double(), FrameStats(), C.c_uint32()
    _check(host_lib().pfhost_stream_frames(ctxs, len(renderers), C.byref(hs), int(n_frames), px, C.byref(wall), C.byref(st),
                                           C.byref(retries)))
    return wall.value, st.as_dict(), retries.value


_lib = None


def lib():
    """Loads lib/libpfcu.so. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PfcuError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(make -C pathfinder-cpp_b200)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)  # (lib/libpfhost.so resolves its pfcu_* references against it)
        vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
        L.pfcu_last_error.restype = C.c_char_p
        L.pfcu_create.argtypes = [i32, C.POINTER(vp)]
        L.pfcu_destroy.argtypes = [vp]
        L.pfcu_destroy.restype = None
        L.pfcu_set_stream.argtypes = [vp, vp]
        L.pfcu_get_stream.argtypes = [vp]
        L.pfcu_get_stream.restype = vp
        L.pfcu_set_area_lut.argtypes = [vp, vp, i32, i32]
        L.pfcu_set_target.argtypes = [vp, i32, i32, vp, sz, vp]
        L.pfcu_set_target_origin.argtypes = [vp, i32, i32]
        L.pfcu_upload_scene.argtypes = [vp, i32, vp, u32, vp, u32]
        L.pfcu_update_scene_range.argtypes = [vp, i32, u32, vp, u32, u32, u32]
        L.pfcu_upload_paint_metadata.argtypes = [vp, vp, u32]
        L.pfcu_alloc_page.argtypes = [vp, u32, i32, i32]
        L.pfcu_upload_page_region.argtypes = [vp, u32, i32, i32, i32, i32, vp]
        L.pfcu_begin_frame.argtypes = [vp]
        L.pfcu_prepare_batch.argtypes = [vp, C.POINTER(BatchDesc)]
        L.pfcu_draw_batch.argtypes = [vp, u32, i32, i32, u32, i32, vp]
        L.pfcu_end_frame.argtypes = [vp, C.POINTER(FrameStats)]
        L.pfcu_submit_frame.argtypes = [vp]
        L.pfcu_wait_frame.argtypes = [vp, C.POINTER(FrameStats)]
        L.pfcu_read_target.argtypes = [vp, vp]
        L.pfcu_read_target_region.argtypes = [vp, i32, i32, i32, i32, vp]
        L.pfcu_read_page.argtypes = [vp, u32, vp]
        L.pfcu_stroke_to_fill.argtypes = [vp, vp, vp, u32, vp, vp, vp, u32, vp, u32, C.POINTER(u32), C.POINTER(u32)]
        L.pfcu_stroke_result.argtypes = [vp, vp, vp, vp]
        L.pfcu_dash_outlines.argtypes = [vp, vp, vp, u32, vp, vp, u32, vp, u32, vp, vp, vp, C.POINTER(u32), C.POINTER(u32)]
        L.pfcu_dash_result.argtypes = [vp, vp, vp, vp, vp]
        L.pfcu_stroke_gpu_ms.argtypes = [vp]
        L.pfcu_stroke_gpu_ms.restype = C.c_float
        L.pfcu_read_target_async.argtypes = [vp, vp, sz]
        L.pfcu_wait_read.argtypes = [vp]
        L.pfcu_host_alloc.argtypes = [sz]
        L.pfcu_host_alloc.restype = vp
        L.pfcu_host_free.argtypes = [vp]
        L.pfcu_host_free.restype = None
        L.pfcu_target_device_ptr.argtypes = [vp, C.POINTER(sz)]
        L.pfcu_target_device_ptr.restype = vp
        for n in ("pfcu_read_lines", "pfcu_read_fills", "pfcu_read_tiles", "pfcu_read_z"):
            getattr(L, n).argtypes = [vp, u32, vp]
            getattr(L, n).restype = C.c_int64
        L.pfcu_read_tile_lists.argtypes = [vp, u32, vp, vp]
        L.pfcu_read_tile_lists.restype = C.c_int64
        L.pfcu_read_mask.argtypes = [vp, u32, vp]
        L.pfcu_set_option.argtypes = [vp, i32, i32]
        L.pfcu_set_profiling.argtypes = [vp, i32]
        L.pfcu_get_stage_times.argtypes = [vp, vp, i32]
        if hasattr(L, "pfcu_set_timeline"):  # (PFCU_LIB may name an older build in an A/B run)
            L.pfcu_set_timeline.argtypes = [vp, C.c_uint32]
            L.pfcu_read_timeline.argtypes = [vp, vp, C.c_int64]
            L.pfcu_read_timeline.restype = C.c_int64
        L.pfcu_graph_capture.argtypes = [vp]
        L.pfcu_graph_launch.argtypes = [vp]
        L.pfcu_graph_finish.argtypes = [vp, C.POINTER(FrameStats)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PfcuError("pfcu error %d: %s" % (rc, lib().pfcu_last_error().decode()))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_desc(batch, keep):
    info = batch["info"]
    d = BatchDesc()
    d.batch_id, d.path_count, d.tile_count = int(info[0]), int(info[1]), int(info[2])
    d.segment_count, d.column_count, d.path_source = int(info[3]), int(info[4]), int(info[5])
    d.clip_batch_id = -1 if int(info[6]) == NONE else int(info[6])
    arrays = [np.ascontiguousarray(batch[k]) for k in ("backdrops", "propagate_metadata", "dice_metadata",
                                                        "tile_path_info")]
    keep.extend(arrays)
    d.backdrops, d.propagate_metadata, d.dice_metadata, d.tile_path_info = [a.ctypes.data for a in arrays]
    for i in range(6):
        d.transform[i] = float(batch["transform"][i])
    return d


class Renderer:
    """One pfcu context bound to one GPU. Usage mirrors Canvas::draw (pathfinder/core/canvas.cpp:557-567):
    `set_scene(scene)` once per scene (upload), `draw(clear)` per frame."""

    def __init__(self, device=0, area_lut=None):
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.pfcu_create(device, C.byref(h)))
        self.h = h
        self.scene = None
        self._descs = None
        if area_lut is not None:
            self.set_area_lut(area_lut)

    def close(self):
        if getattr(self, "h", None):
            self.L.pfcu_destroy(self.h)
            self.h = None
            for ptr in getattr(self, "_pinned", []):
                self.L.pfcu_host_free(ptr)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _check(self.L.pfcu_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def set_area_lut(self, lut):
        lut = np.ascontiguousarray(lut, "u1")
        _check(self.L.pfcu_set_area_lut(self.h, _p(lut), lut.shape[1], lut.shape[0]))

    def set_target(self, width, height, view_box, device_ptr=None, pitch=0):
        vb = np.ascontiguousarray(view_box, "<f4")
        _check(self.L.pfcu_set_target(self.h, int(width), int(height), device_ptr, pitch, _p(vb)))
        self.width, self.height = int(width), int(height)

    # -- scene upload: RendererD3D11::upload_scene + Palette::build_paint_info's uploads
    def upload_segments(self, scene):
        # (the arrays and their addresses are looked up once per scene object: this runs every frame in the e2e legs)
        cached = getattr(self, "_seg_cache", None)
        if cached is None or cached[0] is not scene:
            args = []
            for which, name in ((0, "draw"), (1, "clip")):
                pts = np.ascontiguousarray(scene[name + "_points"], "<f4")
                idx = np.ascontiguousarray(scene[name + "_indices"], "<u4")
                args.append((which, _p(pts), len(pts), _p(idx), len(idx), pts, idx))
            cached = self._seg_cache = (scene, args)
        for which, pp, n_pts, ip, n_idx, _, _ in cached[1]:
            _check(self.L.pfcu_upload_scene(self.h, which, pp, n_pts, ip, n_idx))

    def set_incremental_dice(self, enabled):
        """PFCU_OPT_INCREMENTAL_DICE (default off): later frames dice only the paths whose segments were updated."""
        _check(self.L.pfcu_set_option(self.h, 2, int(bool(enabled))))

    def plan_update(self, old, new):
        """What update_scene needs to go from scene `old` to `new` (same topology, the points of some draw segments
        moved): the changed point range, the segments it belongs to, the new batch descriptors. Host-side preparation, kept
        apart so that a benchmark can time the C-ABI calls alone."""
        assert np.array_equal(old["draw_indices"], new["draw_indices"]) and old["draw_points"].shape == new["draw_points"].shape
        changed = np.nonzero((old["draw_points"] != new["draw_points"]).any(axis=1))[0]
        first_point, n_points = int(changed[0]), int(changed[-1] - changed[0] + 1)
        fp = new["draw_indices"][:, 0]
        first_seg = int(np.searchsorted(fp, first_point, side="right") - 1)
        end_seg = int(np.searchsorted(fp, first_point + n_points - 1, side="right"))
        keep = []
        descs = {"clip": [make_desc(b, keep) for b in new["clip_batches"]],
                 "draw": [make_desc(b, keep) for b in new["draw_batches"]]}
        pts = np.ascontiguousarray(new["draw_points"][first_point:first_point + n_points], "<f4")
        return dict(scene=new, first_point=first_point, points=pts, first_seg=first_seg, n_seg=end_seg - first_seg,
                    descs=descs, keep=keep)

    def apply_update(self, plan):
        """pfcu_update_scene_range with a prepared plan; the next draw() renders plan['scene']."""
        _check(self.L.pfcu_update_scene_range(self.h, 0, plan["first_point"], _p(plan["points"]), len(plan["points"]),
                                              plan["first_seg"], plan["n_seg"]))
        self.scene = plan["scene"]
        self._keep = plan["keep"]
        self._descs = plan["descs"]

    def update_scene(self, scene):
        """Switch to `scene`, which differs from the current one only in the POINTS of some draw segments (same indices,
        same batch structure: a path moved): uploads the changed point range (pfcu_update_scene_range) and takes the new
        batch metadata. Returns (first_segment, n_segments) of the update."""
        plan = self.plan_update(self.scene, scene)
        self.apply_update(plan)
        return plan["first_seg"], plan["n_seg"]

    def upload_paints(self, scene):
        md = np.ascontiguousarray(scene["metadata"], "<u2")
        _check(self.L.pfcu_upload_paint_metadata(self.h, _p(md), md.shape[0]))
        for page, px in scene.get("pages", {}).items():
            px = np.ascontiguousarray(px, "u1")
            _check(self.L.pfcu_alloc_page(self.h, int(page), px.shape[1], px.shape[0]))
            _check(self.L.pfcu_upload_page_region(self.h, int(page), 0, 0, px.shape[1], px.shape[0], _p(px)))

    def set_scene(self, scene, target_ptr=None, pitch=0):
        self.scene = scene
        self.set_target(scene["width"], scene["height"], scene["view_box"], target_ptr, pitch)
        org = scene.get("origin_tiles", (0, 0))
        _check(self.L.pfcu_set_target_origin(self.h, int(org[0]) * 16, int(org[1]) * 16))
        self.upload_segments(scene)
        self.upload_paints(scene)
        self._keep = []
        self._descs = {"clip": [make_desc(b, self._keep) for b in scene["clip_batches"]],
                       "draw": [make_desc(b, self._keep) for b in scene["draw_batches"]]}

    # -- frame
    def draw(self, clear=True, clear_color=(0.0, 0.0, 0.0, 0.0), upload=False, wait=True):
        """RendererD3D11::draw: clip batches in reverse, then prepare + composite every draw batch.
        upload=True re-uploads the segments first (what the reference does every frame, renderer.cpp:314).
        wait=False submits the frame (pfcu_submit_frame) and returns; wait() collects it."""
        scene = self.scene
        if upload:
            self.upload_segments(scene)
        L, h = self.L, self.h
        _check(L.pfcu_begin_frame(h))
        plan = getattr(self, "_draw_plan", None)
        if plan is None or plan[0] is not self._descs:  # per scene: the calls of a frame, arguments resolved
            clips = [C.byref(d) for d in reversed(self._descs["clip"]) if d.path_count > 0]
            draws = []
            for d, b in zip(self._descs["draw"], scene["draw_batches"]):
                info = b["info"]
                color_page = -1 if int(info[7]) == NONE else int(info[7])
                flags = 0 if int(info[8]) == NONE else int(info[8])
                target = -1 if int(info[10]) == NONE else int(info[11])
                draws.append((C.byref(d), d.batch_id, target, color_page, flags))
            plan = self._draw_plan = (self._descs, clips, draws, np.zeros(4, "<f4"))
        _, clips, draws, zero = plan
        for ref in clips:
            _check(L.pfcu_prepare_batch(h, ref))
        cc = np.array(clear_color, "<f4")
        cc_p, zero_p = _p(cc), _p(zero)
        first = bool(clear)
        for ref, batch_id, target, color_page, flags in draws:
            _check(L.pfcu_prepare_batch(h, ref))
            if target < 0:
                _check(L.pfcu_draw_batch(h, batch_id, -1, color_page, flags, int(first), cc_p))
                first = False
            else:
                _check(L.pfcu_draw_batch(h, batch_id, target, color_page, flags, 1, zero_p))
        if not wait:
            _check(L.pfcu_submit_frame(self.h))
            return None
        st = FrameStats()
        _check(L.pfcu_end_frame(self.h, C.byref(st)))
        return st.as_dict()

    def wait(self):
        """Second half of draw(wait=False): pfcu_wait_frame."""
        st = FrameStats()
        _check(self.L.pfcu_wait_frame(self.h, C.byref(st)))
        return st.as_dict()

    # -- options / measurement
    def set_retain_frame_graph(self, enabled):
        """PFCU_OPT_RETAIN_FRAME_GRAPH (default on): identical consecutive frames become one graph launch."""
        _check(self.L.pfcu_set_option(self.h, 0, int(bool(enabled))))

    def set_fused_fill(self, enabled):
        """PFCU_OPT_FUSED_FILL (default off): the tile kernel rasterizes a draw batch's masks itself (no separate fill)."""
        _check(self.L.pfcu_set_option(self.h, 3, int(bool(enabled))))

    def set_concurrent_batches(self, enabled):
        """PFCU_OPT_CONCURRENT_BATCHES (default on): the batches of a frame prepare side by side on four stream pairs."""
        _check(self.L.pfcu_set_option(self.h, 5, int(bool(enabled))))

    def set_order_tile_groups(self, enabled):
        """PFCU_OPT_ORDER_TILE_GROUPS (default 4): the tile kernel starts with the groups of 16 tiles that have at least
        this many masked tiles; 0 / False = grid order."""
        _check(self.L.pfcu_set_option(self.h, 4, 4 if enabled is True else int(enabled)))

    def set_fill_culled_tiles(self, enabled):
        """PFCU_OPT_FILL_CULLED_TILES (default off): rasterize the masks of z-culled tiles too, like fill.comp."""
        _check(self.L.pfcu_set_option(self.h, 1, int(bool(enabled))))

    def set_profiling(self, enabled):
        _check(self.L.pfcu_set_profiling(self.h, int(bool(enabled))))

    def stage_times(self):
        ms = np.zeros(len(STAGES), "<f4")
        _check(self.L.pfcu_get_stage_times(self.h, _p(ms), len(STAGES)))
        return dict(zip(STAGES, (float(x) for x in ms)))

    def set_timeline(self, capacity):
        """pfcu_set_timeline: record (stage, SM, placed / started / ended in %globaltimer ns) for every CTA; 0 = off."""
        _check(self.L.pfcu_set_timeline(self.h, int(capacity)))
        self._timeline_cap = int(capacity)

    def read_timeline(self):
        """The records since the last call (TIMELINE_DT); waits for the context first."""
        out = np.zeros(self._timeline_cap, TIMELINE_DT)
        n = self.L.pfcu_read_timeline(self.h, _p(out), len(out))
        if n < 0:
            raise PfcuError(self.L.pfcu_last_error().decode())
        return out[: min(n, len(out))]

    def graph_capture(self):
        _check(self.L.pfcu_graph_capture(self.h))

    def graph_launch(self):
        _check(self.L.pfcu_graph_launch(self.h))

    def graph_finish(self):
        st = FrameStats()
        _check(self.L.pfcu_graph_finish(self.h, C.byref(st)))
        return st.as_dict()

    def pixels(self):
        px = np.zeros((self.height, self.width, 4), "u1")
        _check(self.L.pfcu_read_target(self.h, _p(px)))
        return px

    def stroke_to_fill(self, points, flags, contour_first, closed, style_index, styles):
        """pfcu_stroke_to_fill + pfcu_stroke_result: OutlineStrokeToFill::offset for a batch of contours.
        styles: (k, 4) rows of (line_width, line_cap, line_join, miter_limit). Returns (points (n, 2) f32, flags (n,) u8,
        contour_first (m + 1,) u32, gpu_ms)."""
        pts = np.ascontiguousarray(points, "<f4").reshape(-1, 2)
        fl = np.ascontiguousarray(flags, "u1")
        cf = np.ascontiguousarray(contour_first, "<u4")
        cl = np.ascontiguousarray(closed, "u1")
        si = np.ascontiguousarray(style_index, "<u4")
        st = np.zeros(len(styles), np.dtype([("w", "<f4"), ("cap", "<i4"), ("join", "<i4"), ("miter", "<f4"),
                                             ("transform", "<f4", (6,))]))
        for i, row in enumerate(styles):  # (line_width, cap, join, miter_limit[, (m11, m21, m12, m22, m13, m23)])
            st[i] = (row[0], int(row[1]), int(row[2]), row[3], row[4] if len(row) > 4 else (1, 0, 0, 1, 0, 0))
        nc, npts = C.c_uint32(), C.c_uint32()
        _check(self.L.pfcu_stroke_to_fill(self.h, _p(pts), _p(fl), len(pts), _p(cf), _p(cl), _p(si), len(cl), _p(st), len(st),
                                          C.byref(nc), C.byref(npts)))
        op = np.zeros((npts.value, 2), "<f4")
        of = np.zeros(npts.value, "u1")
        oc = np.zeros(nc.value + 1, "<u4")
        _check(self.L.pfcu_stroke_result(self.h, _p(op), _p(of), _p(oc)))
        return op, of, oc, float(self.L.pfcu_stroke_gpu_ms(self.h))

    def dash_outlines(self, points, flags, contour_first, closed, outline_first, dashes, dash_first, dash_offset):
        """pfcu_dash_outlines + pfcu_dash_result: OutlineDash for a batch of outlines. Returns (points, flags,
        contour_first, outline_first, gpu_ms)."""
        pts = np.ascontiguousarray(points, "<f4").reshape(-1, 2)
        fl = np.ascontiguousarray(flags, "u1")
        cf = np.ascontiguousarray(contour_first, "<u4")
        cl = np.ascontiguousarray(closed, "u1")
        of_ = np.ascontiguousarray(outline_first, "<u4")
        da = np.ascontiguousarray(dashes, "<f4")
        df = np.ascontiguousarray(dash_first, "<u4")
        do = np.ascontiguousarray(dash_offset, "<f4")
        nc, npts = C.c_uint32(), C.c_uint32()
        _check(self.L.pfcu_dash_outlines(self.h, _p(pts), _p(fl), len(pts), _p(cf), _p(cl), len(cl), _p(of_), len(do), _p(da),
                                         _p(df), _p(do), C.byref(nc), C.byref(npts)))
        op = np.zeros((npts.value, 2), "<f4")
        ofl = np.zeros(npts.value, "u1")
        oc = np.zeros(nc.value + 1, "<u4")
        oo = np.zeros(len(do) + 1, "<u4")
        _check(self.L.pfcu_dash_result(self.h, _p(op), _p(ofl), _p(oc), _p(oo)))
        return op, ofl, oc, oo, float(self.L.pfcu_stroke_gpu_ms(self.h))

    def pinned_frame(self):
        """A page-locked (height, width, 4) u8 array for read_async (pfcu_host_alloc); freed with the renderer."""
        n = self.height * self.width * 4
        ptr = self.L.pfcu_host_alloc(n)
        if not ptr:
            raise PfcuError(self.L.pfcu_last_error().decode())
        self._pinned = getattr(self, "_pinned", []) + [ptr]
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr)).reshape(self.height, self.width, 4)

    def read_async(self, out):
        """pfcu_read_target_async: enqueue the read-back of the frame in flight into `out` (use pinned_frame())."""
        _check(self.L.pfcu_read_target_async(self.h, _p(out), out.strides[0]))

    def wait_read(self):
        _check(self.L.pfcu_wait_read(self.h))

    def pixels_region(self, x, y, width, height):
        px = np.zeros((height, width, 4), "u1")
        _check(self.L.pfcu_read_target_region(self.h, int(x), int(y), int(width), int(height), _p(px)))
        return px

    # -- parity taps
    def _tap(self, fn, batch_id, dt):
        n = fn(self.h, batch_id, None)
        if n < 0:
            _check(int(n))
        a = np.zeros(n, dt)
        if n:
            m = fn(self.h, batch_id, _p(a))
            a = a[:m]
        return a

    def lines(self, batch_id):
        return self._tap(self.L.pfcu_read_lines, batch_id, LINE_DT)

    def fills(self, batch_id):
        return self._tap(self.L.pfcu_read_fills, batch_id, FILL_DT)

    def tiles(self, batch_id):
        return self._tap(self.L.pfcu_read_tiles, batch_id, TILE_DT)

    def z(self, batch_id):
        return self._tap(self.L.pfcu_read_z, batch_id, "<i4")

    def tile_lists(self, batch_id):
        fbt = ((self.width + 15) // 16) * ((self.height + 15) // 16)
        n = self.L.pfcu_read_tile_lists(self.h, batch_id, None, None)
        off = np.zeros(fbt + 1, "<u4")
        t = np.zeros(max(int(n), 1), "<u4")
        self.L.pfcu_read_tile_lists(self.h, batch_id, _p(off), _p(t))
        return off, t[:n]

    def mask(self, alpha_id):
        m = np.zeros(256, "u1")
        _check(self.L.pfcu_read_mask(self.h, int(alpha_id), _p(m)))
        return m.reshape(16, 16)

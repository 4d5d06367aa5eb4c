"""TEST INFRASTRUCTURE (oracle). ctypes view of oracle/_ref/libpfref.so.

libpfref.so is the UNMODIFIED reference (floppyhammer/pathfinder-cpp) CPU code compiled by
oracle/Makefile from /root/reference, driven through oracle/ref_harness/ref_harness.cpp.
Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpfref.so")

# dtypes of the reference's POD records (sizes probe-checked by static_asserts in the harness).
FILL_DT = np.dtype([("from_x", "<u2"), ("from_y", "<u2"), ("to_x", "<u2"), ("to_y", "<u2"), ("link", "<u4")])
TILE9_DT = np.dtype([("tile_x", "<i2"), ("tile_y", "<i2"), ("alpha_tile_id", "<u4"), ("path_id", "<u4"),
                     ("ctrl", "u1"), ("backdrop", "i1"), ("metadata_id", "<u2")])
CLIP9_DT = np.dtype([("dest_tile_id", "<u4"), ("dest_backdrop", "<i4"), ("src_tile_id", "<u4"),
                     ("src_backdrop", "<i4")])
# core/d3d11/gpu_data.h:54-94
BACKDROP_DT = np.dtype([("initial_backdrop", "<i4"), ("tile_x_offset", "<i4"), ("path_index", "<u4")])
PROPAGATE_DT = np.dtype([("rect", "<i4", (4,)), ("tile_offset", "<u4"), ("path_index", "<u4"), ("z_write", "<u4"),
                         ("clip_path_index", "<u4"), ("backdrop_offset", "<u4"), ("pad", "<u4", (3,))])
DICE_DT = np.dtype([("global_path_id", "<u4"), ("first_global_segment_index", "<u4"),
                    ("first_batch_segment_index", "<u4"), ("pad", "<u4")])
TILE_PATH_INFO_DT = np.dtype([("tile_min_x", "<i2"), ("tile_min_y", "<i2"), ("tile_max_x", "<i2"),
                              ("tile_max_y", "<i2"), ("first_tile_index", "<u4"), ("color", "<u2"), ("ctrl", "u1"),
                              ("backdrop", "i1")])
assert FILL_DT.itemsize == 12 and TILE9_DT.itemsize == 16 and CLIP9_DT.itemsize == 16
assert BACKDROP_DT.itemsize == 12 and PROPAGATE_DT.itemsize == 48 and DICE_DT.itemsize == 16
assert TILE_PATH_INFO_DT.itemsize == 16


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, sz, u32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)
        L.pfref_scene_from_svg.restype = vp
        L.pfref_scene_from_svg.argtypes = [C.c_char_p, sz, C.c_int, C.c_int, C.c_float]
        L.pfref_scene_demo.restype = vp
        L.pfref_scene_demo.argtypes = [C.c_int, C.c_int, C.c_float, C.c_char_p, sz, C.c_int]
        L.pfref_scene_paints.restype = vp
        L.pfref_scene_paints.argtypes = [C.c_int, C.c_int, C.c_float, C.c_char_p, sz]
        L.pfref_scene_free.argtypes = [vp]
        L.pfref_scene_counts.argtypes = [vp, u32p]
        L.pfref_view_box.argtypes = [vp, C.POINTER(C.c_float)]
        L.pfref_build_d3d11.argtypes = [vp]
        L.pfref_d3d11_points.restype = sz
        L.pfref_d3d11_points.argtypes = [vp, C.c_int, vp]
        L.pfref_d3d11_indices.restype = sz
        L.pfref_d3d11_indices.argtypes = [vp, C.c_int, vp]
        L.pfref_d3d11_num_batches.argtypes = [vp, C.c_int]
        L.pfref_d3d11_batch_info.argtypes = [vp, C.c_int, C.c_int, u32p]
        L.pfref_d3d11_batch_array.restype = sz
        L.pfref_d3d11_batch_array.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.pfref_metadata_texels.restype = sz
        L.pfref_metadata_texels.argtypes = [vp, vp, u32p]
        L.pfref_area_lut.argtypes = [vp, vp, C.POINTER(C.c_int)]
        L.pfref_num_pages.argtypes = [vp]
        L.pfref_page.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int)]
        L.pfref_build_d3d9.argtypes = [vp]
        L.pfref_d3d9_fills.restype = sz
        L.pfref_d3d9_fills.argtypes = [vp, vp]
        L.pfref_d3d9_num_batches.argtypes = [vp]
        L.pfref_d3d9_batch_tiles.restype = sz
        L.pfref_d3d9_batch_tiles.argtypes = [vp, C.c_int, vp]
        L.pfref_d3d9_batch_clips.restype = sz
        L.pfref_d3d9_batch_clips.argtypes = [vp, C.c_int, vp]
        L.pfref_d3d9_batch_z.restype = sz
        L.pfref_d3d9_batch_z.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int)]
        L.pfref_d3d9_batch_info.argtypes = [vp, C.c_int, u32p]
        L.pfref_time_d3d9_build.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.pfref_time_d3d11_build.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.pfref_asset.restype = C.c_void_p
        L.pfref_asset.argtypes = [C.c_char_p, C.POINTER(sz)]
        _lib = L
    return _lib


def asset(name):
    """Bytes of a reference asset linked into libpfref.so (tiger.svg, features.svg, sea.png)."""
    n = C.c_size_t(0)
    p = lib().pfref_asset(name.encode(), C.byref(n))
    if not p:
        raise KeyError(name)
    return C.string_at(p, n.value)


def _arr(dtype, count):
    return np.zeros(count, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """One reference `Scene` built by the reference front end."""

    def __init__(self, handle, width, height):
        if not handle:
            raise RuntimeError("reference failed to build the scene")
        self.h, self.width, self.height = handle, width, height

    @classmethod
    def from_svg(cls, svg_bytes, width, height, scale):
        return cls(lib().pfref_scene_from_svg(svg_bytes, len(svg_bytes), width, height, scale), width, height)

    @classmethod
    def demo(cls, width, height, scale, image_bytes=b"", features=0x3f):
        return cls(lib().pfref_scene_demo(width, height, scale, image_bytes, len(image_bytes), features), width,
                   height)

    @classmethod
    def paints(cls, width, height, scale, image_bytes=b""):
        """Every blend mode, radial gradients, a repeating unsmoothed image pattern (ref_harness.cpp)."""
        return cls(lib().pfref_scene_paints(width, height, scale, image_bytes, len(image_bytes)), width, height)

    def close(self):
        if self.h:
            lib().pfref_scene_free(self.h)
            self.h = None

    def translate_draw_path(self, index, dx, dy):
        """Outline::transform(Transform2::from_translation) on one draw path (an animated path)."""
        L = lib()
        L.pfref_scene_translate_draw_path.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float]
        if L.pfref_scene_translate_draw_path(self.h, index, dx, dy) != 0:
            raise IndexError(index)

    def counts(self):
        out = (C.c_uint32 * 4)()
        lib().pfref_scene_counts(self.h, out)
        return dict(draw_paths=out[0], clip_paths=out[1], paints=out[2], display_items=out[3])

    def view_box(self):
        out = (C.c_float * 4)()
        lib().pfref_view_box(self.h, out)
        return tuple(out)

    # ---- SceneBuilderD3D11: what the reference hands to its GPU-driven renderer
    def build_d3d11(self):
        """Returns the scene in the form the drop-in boundary receives (see pfcu.Scene)."""
        L = lib()
        L.pfref_build_d3d11(self.h)
        scene = {"width": self.width, "height": self.height, "view_box": np.array(self.view_box(), "<f4")}
        for which, name in ((0, "draw"), (1, "clip")):
            n = L.pfref_d3d11_points(self.h, which, None)
            pts = _arr("<f4", n * 2)
            L.pfref_d3d11_points(self.h, which, _ptr(pts))
            n = L.pfref_d3d11_indices(self.h, which, None)
            idx = _arr("<u4", n * 2)
            L.pfref_d3d11_indices(self.h, which, _ptr(idx))
            scene[name + "_points"] = pts.reshape(-1, 2)
            scene[name + "_indices"] = idx.reshape(-1, 2)
        for kind, name in ((0, "draw_batches"), (1, "clip_batches")):
            batches = []
            for i in range(L.pfref_d3d11_num_batches(self.h, kind)):
                info = (C.c_uint32 * 16)()
                L.pfref_d3d11_batch_info(self.h, kind, i, info)
                b = {"info": np.array(list(info), "<u4")}
                for which, key, dt in ((0, "backdrops", BACKDROP_DT), (1, "propagate_metadata", PROPAGATE_DT),
                                       (2, "dice_metadata", DICE_DT), (3, "tile_path_info", TILE_PATH_INFO_DT)):
                    n = L.pfref_d3d11_batch_array(self.h, kind, i, which, None)
                    a = _arr(dt, n)
                    if n:
                        L.pfref_d3d11_batch_array(self.h, kind, i, which, _ptr(a))
                    b[key] = a
                t = _arr("<f4", 6)
                L.pfref_d3d11_batch_array(self.h, kind, i, 4, _ptr(t))
                b["transform"] = t
                batches.append(b)
            scene[name] = batches
        rows = C.c_uint32(0)
        n = L.pfref_metadata_texels(self.h, None, C.byref(rows))
        md = _arr("<u2", n)
        L.pfref_metadata_texels(self.h, _ptr(md), C.byref(rows))
        scene["metadata"] = md.reshape(-1, 1280 * 4)[: max(rows.value, 1)].copy()
        pages = {}
        for p in range(L.pfref_num_pages(self.h)):
            wh = (C.c_int * 2)()
            if L.pfref_page(self.h, p, None, wh) == 0:
                px = _arr("u1", wh[0] * wh[1] * 4)
                L.pfref_page(self.h, p, _ptr(px), wh)
                pages[p] = px.reshape(wh[1], wh[0], 4)
        scene["pages"] = pages
        return scene

    def area_lut(self):
        wh = (C.c_int * 2)()
        lib().pfref_area_lut(self.h, None, wh)
        a = _arr("u1", wh[0] * wh[1] * 4)
        lib().pfref_area_lut(self.h, _ptr(a), wh)
        return a.reshape(wh[1], wh[0], 4)

    # ---- SceneBuilderD3D9: the hybrid CPU tiler (parity truth)
    def build_d3d9(self):
        L = lib()
        nb = L.pfref_build_d3d9(self.h)
        n = L.pfref_d3d9_fills(self.h, None)
        fills = _arr(FILL_DT, n)
        if n:
            L.pfref_d3d9_fills(self.h, _ptr(fills))
        batches = []
        for i in range(nb):
            n = L.pfref_d3d9_batch_tiles(self.h, i, None)
            tiles = _arr(TILE9_DT, n)
            if n:
                L.pfref_d3d9_batch_tiles(self.h, i, _ptr(tiles))
            n = L.pfref_d3d9_batch_clips(self.h, i, None)
            clips = _arr(CLIP9_DT, n)
            if n:
                L.pfref_d3d9_batch_clips(self.h, i, _ptr(clips))
            rect = (C.c_int * 4)()
            n = L.pfref_d3d9_batch_z(self.h, i, None, rect)
            z = _arr("<u4", n)
            if n:
                L.pfref_d3d9_batch_z(self.h, i, _ptr(z), rect)
            info = (C.c_uint32 * 4)()
            L.pfref_d3d9_batch_info(self.h, i, info)
            batches.append(dict(tiles=tiles, clips=clips, z=z.reshape(rect[3] - rect[1], rect[2] - rect[0]),
                                z_rect=tuple(rect), info=np.array(list(info), "<u4")))
        return dict(fills=fills, batches=batches)

    def time_d3d9_build(self, iters):
        out = (C.c_double * iters)()
        lib().pfref_time_d3d9_build(self.h, iters, out)
        return np.array(list(out))

    def time_d3d11_build(self, iters):
        out = (C.c_double * iters)()
        lib().pfref_time_d3d11_build(self.h, iters, out)
        return np.array(list(out))


def stroke_outline(points, flags, contour_first, closed, line_width, line_cap=0, line_join=0, miter_limit=10.0):
    """The reference's OutlineStrokeToFill::offset (core/stroke.cpp:124-167) on an outline given as arrays (the oracle of
    pfcu_stroke_to_fill). Returns (points (n, 2) f32, flags (n,) u8, contour_first (m + 1,) u32, cpu_ms)."""
    L = lib()
    fn = L.pfref_stroke_outline
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.c_int, C.c_float,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    pts = np.ascontiguousarray(points, "<f4").reshape(-1, 2)
    fl = np.ascontiguousarray(flags, "u1")
    cf = np.ascontiguousarray(contour_first, "<u4")
    cl = np.ascontiguousarray(closed, "u1")
    nc = C.c_uint32()
    args = (pts.ctypes.data, fl.ctypes.data, cf.ctypes.data, cl.ctypes.data, len(cl), line_width, line_cap, line_join,
            miter_limit)
    n = fn(*args, None, None, None, C.byref(nc), None)
    op = np.zeros((n, 2), "<f4")
    of = np.zeros(n, "u1")
    oc = np.zeros(nc.value + 1, "<u4")
    ms = C.c_double()
    fn(*args, op.ctypes.data, of.ctypes.data, oc.ctypes.data, C.byref(nc), C.byref(ms))
    return op, of, oc, ms.value


def dash_outline(points, flags, contour_first, closed, dashes, offset=0.0):
    """The reference's OutlineDash (core/dash.cpp) on one outline given as arrays (the oracle of pfcu_dash_outlines).
    Returns (points, flags, contour_first)."""
    L = lib()
    fn = L.pfref_dash_outline
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_float,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
    pts = np.ascontiguousarray(points, "<f4").reshape(-1, 2)
    fl = np.ascontiguousarray(flags, "u1")
    cf = np.ascontiguousarray(contour_first, "<u4")
    cl = np.ascontiguousarray(closed, "u1")
    da = np.ascontiguousarray(dashes, "<f4")
    nc = C.c_uint32()
    args = (pts.ctypes.data, fl.ctypes.data, cf.ctypes.data, cl.ctypes.data, len(cl), da.ctypes.data, len(da), offset)
    n = fn(*args, None, None, None, C.byref(nc))
    op = np.zeros((n, 2), "<f4")
    of = np.zeros(n, "u1")
    oc = np.zeros(nc.value + 1, "<u4")
    fn(*args, op.ctypes.data, of.ctypes.data, oc.ctypes.data, C.byref(nc))
    return op, of, oc


def svg_stroke_inputs(svg_bytes):
    """Every stroked shape of an SVG as SvgScene hands it to Canvas::stroke_path (core/svg.cpp:155-196): outline before
    dash / stroke / transform + style. Returns dict(points, flags, contour_first, closed, shape_first, styles (n, 5:
    width, cap, join, miter, dash offset), dashes, dash_first)."""
    L = lib()
    fn = L.pfref_svg_stroke_inputs
    fn.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p] + [C.c_void_p] * 8
    counts = np.zeros(4, "<u4")
    if fn(svg_bytes, len(svg_bytes), counts.ctypes.data, *([None] * 8)) != 0:
        raise RuntimeError("nanosvg could not parse the document")
    ns, nc, npt, nd = (int(x) for x in counts)
    out = dict(points=np.zeros((npt, 2), "<f4"), flags=np.zeros(npt, "u1"), contour_first=np.zeros(nc + 1, "<u4"),
               closed=np.zeros(nc, "u1"), shape_first=np.zeros(ns + 1, "<u4"), styles=np.zeros((ns, 5), "<f4"),
               dashes=np.zeros(max(nd, 1), "<f4"), dash_first=np.zeros(ns + 1, "<u4"))
    fn(svg_bytes, len(svg_bytes), counts.ctypes.data, *(out[k].ctypes.data for k in
       ("points", "flags", "contour_first", "closed", "shape_first", "styles", "dashes", "dash_first")))
    out["dashes"] = out["dashes"][:nd]
    return out

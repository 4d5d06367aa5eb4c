"""TEST INFRASTRUCTURE (oracle). ctypes view of oracle/libpforacle.so (the plain-C restatement, pf_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module: it is the
checker for the CUDA path, never the thing shipped or measured as the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpforacle.so")

LINE_DT = np.dtype([("from_x", "<f4"), ("from_y", "<f4"), ("to_x", "<f4"), ("to_y", "<f4"), ("path_index", "<u4")])
FILL_DT = np.dtype([("tile_index", "<u4"), ("from_x", "<u2"), ("from_y", "<u2"), ("to_x", "<u2"), ("to_y", "<u2")])
TILE_DT = np.dtype([("alpha_tile_id", "<i4"), ("clip_alpha_tile_id", "<i4"), ("fill_count", "<i4"),
                    ("backdrop", "i1"), ("backdrop_delta", "i1"), ("backdrop_d3d9", "i1"), ("listed", "u1")])
assert LINE_DT.itemsize == 20 and FILL_DT.itemsize == 12 and TILE_DT.itemsize == 16


class BatchDesc(C.Structure):
    _fields_ = [("batch_id", C.c_uint32), ("path_count", C.c_uint32), ("tile_count", C.c_uint32),
                ("segment_count", C.c_uint32), ("column_count", C.c_uint32), ("path_source", C.c_int32),
                ("clip_batch_id", C.c_int32), ("backdrops", C.c_void_p), ("propagate_metadata", C.c_void_p),
                ("dice_metadata", C.c_void_p), ("tile_path_info", C.c_void_p), ("transform", C.c_float * 6)]


def build():
    """Compile the restatement with gcc (building the checker is not using it)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.pfo_frame_create.restype = vp
        L.pfo_frame_create.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, C.c_int]
        L.pfo_frame_set_origin.argtypes = [vp, C.c_int, C.c_int]
        L.pfo_frame_destroy.argtypes = [vp]
        L.pfo_frame_reset.argtypes = [vp]
        L.pfo_frame_set_segments.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, C.c_uint32]
        L.pfo_frame_set_metadata.argtypes = [vp, vp, C.c_uint32]
        L.pfo_frame_set_page.argtypes = [vp, C.c_uint32, C.c_int, C.c_int, vp]
        L.pfo_frame_prepare_batch.argtypes = [vp, C.POINTER(BatchDesc)]
        L.pfo_frame_prepare_batch_geometry_only.argtypes = [vp, C.POINTER(BatchDesc)]
        L.pfo_batch_counts.argtypes = [vp, C.c_int, vp]
        for name in ("pfo_batch_lines", "pfo_batch_clipped_lines", "pfo_batch_fills", "pfo_batch_tiles"):
            getattr(L, name).restype = sz
            getattr(L, name).argtypes = [vp, C.c_int, vp]
        L.pfo_batch_column_backdrops.restype = sz
        L.pfo_batch_column_backdrops.argtypes = [vp, C.c_int, vp]
        L.pfo_batch_z.restype = sz
        L.pfo_batch_z.argtypes = [vp, C.c_int, vp, vp]
        L.pfo_batch_tile_lists.restype = sz
        L.pfo_batch_tile_lists.argtypes = [vp, C.c_int, vp, vp]
        L.pfo_frame_mask.argtypes = [vp, C.c_uint32, vp]
        L.pfo_frame_draw_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, vp]
        L.pfo_frame_pixels.argtypes = [vp, vp]
        L.pfo_frame_page_pixels.argtypes = [vp, C.c_uint32, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_desc(batch, keep):
    """BatchDesc from a scene-dict batch (see oracle/pfref.py RefScene.build_d3d11 / tests/scenes.py)."""
    info = batch["info"]
    d = BatchDesc()
    d.batch_id = int(info[0])
    d.path_count = int(info[1])
    d.tile_count = int(info[2])
    d.segment_count = int(info[3])
    d.column_count = int(info[4])
    d.path_source = int(info[5])
    d.clip_batch_id = -1 if int(info[6]) == 0xFFFFFFFF else int(info[6])
    arrays = [np.ascontiguousarray(batch[k]) for k in ("backdrops", "propagate_metadata", "dice_metadata",
                                                        "tile_path_info")]
    keep.extend(arrays)
    d.backdrops, d.propagate_metadata, d.dice_metadata, d.tile_path_info = [a.ctypes.data for a in arrays]
    for i in range(6):
        d.transform[i] = float(batch["transform"][i])
    return d


class Frame:
    """Oracle state for one frame of one scene."""

    def __init__(self, scene, area_lut=None, pixel_model=0):
        self.scene = scene
        vb = np.ascontiguousarray(scene["view_box"], "<f4")
        lut = np.ascontiguousarray(area_lut, "u1") if area_lut is not None else None
        lw, lh = (lut.shape[1], lut.shape[0]) if lut is not None else (0, 0)
        self.h = lib().pfo_frame_create(int(scene["width"]), int(scene["height"]), _p(vb), _p(lut), lw, lh)
        org = scene.get("origin_tiles", (0, 0))
        lib().pfo_frame_set_origin(self.h, int(org[0]), int(org[1]))
        if pixel_model:  # 1: the hybrid raster shaders' pixel semantics on the same geometry (pf_oracle.c, pixel_model)
            lib().pfo_frame_set_pixel_model(self.h, int(pixel_model))
        self.fb_tiles = ((int(scene["width"]) + 15) // 16) * ((int(scene["height"]) + 15) // 16)
        for which, name in ((0, "draw"), (1, "clip")):
            pts = np.ascontiguousarray(scene[name + "_points"], "<f4")
            idx = np.ascontiguousarray(scene[name + "_indices"], "<u4")
            lib().pfo_frame_set_segments(self.h, which, _p(pts), len(pts), _p(idx), len(idx))
        md = np.ascontiguousarray(scene["metadata"], "<u2")
        lib().pfo_frame_set_metadata(self.h, _p(md), md.shape[0])
        for page, px in scene.get("pages", {}).items():
            px = np.ascontiguousarray(px, "u1")
            lib().pfo_frame_set_page(self.h, int(page), px.shape[1], px.shape[0], _p(px))
        self.slots = {}

    def close(self):
        if self.h:
            lib().pfo_frame_destroy(self.h)
            self.h = None

    def prepare(self, batch, geometry_only=False):
        keep = []
        d = make_desc(batch, keep)
        fn = lib().pfo_frame_prepare_batch_geometry_only if geometry_only else lib().pfo_frame_prepare_batch
        slot = fn(self.h, C.byref(d))
        if slot < 0:
            raise RuntimeError("oracle prepare failed")
        self.slots[int(batch["info"][0])] = slot
        return slot

    def prepare_all(self, geometry_only=False):
        """Clip batches in reverse storage order, then draw batches (core/d3d11/renderer.cpp:318-332)."""
        out = {"clip": [], "draw": []}
        for b in reversed(self.scene["clip_batches"]):
            if int(b["info"][1]) > 0:
                out["clip"].append(self.prepare(b, geometry_only))
        for b in self.scene["draw_batches"]:
            out["draw"].append(self.prepare(b, geometry_only))
        return out

    def counts(self, slot):
        c = np.zeros(8, "<u4")
        lib().pfo_batch_counts(self.h, slot, _p(c))
        return dict(lines=int(c[0]), fills=int(c[1]), alpha_tiles=int(c[2]), first_alpha=int(c[3]),
                    listed=int(c[4]), listed_culled=int(c[5]), max_list=int(c[6]))

    def _get(self, fn, slot, dt):
        n = fn(self.h, slot, None)
        a = np.zeros(n, dt)
        if n:
            fn(self.h, slot, _p(a))
        return a

    def lines(self, slot):
        return self._get(lib().pfo_batch_lines, slot, LINE_DT)

    def clipped_lines(self, slot):
        return self._get(lib().pfo_batch_clipped_lines, slot, LINE_DT)

    def fills(self, slot):
        return self._get(lib().pfo_batch_fills, slot, FILL_DT)

    def tiles(self, slot):
        return self._get(lib().pfo_batch_tiles, slot, TILE_DT)

    def column_backdrops(self, slot):
        return self._get(lib().pfo_batch_column_backdrops, slot, "<i4")

    def z(self, slot):
        z11 = np.zeros(self.fb_tiles, "<i4")
        z9 = np.zeros(self.fb_tiles, "<u4")
        lib().pfo_batch_z(self.h, slot, _p(z11), _p(z9))
        return z11, z9

    def tile_lists(self, slot):
        n = lib().pfo_batch_tile_lists(self.h, slot, None, None)
        off = np.zeros(self.fb_tiles + 1, "<u4")
        t = np.zeros(max(n, 1), "<u4")
        lib().pfo_batch_tile_lists(self.h, slot, _p(off), _p(t))
        return off, t[:n]

    def mask(self, alpha_id):
        m = np.zeros(256, "u1")
        if lib().pfo_frame_mask(self.h, alpha_id, _p(m)) != 0:
            raise IndexError(alpha_id)
        return m.reshape(16, 16)

    def draw(self, slot, target_page=-1, color_page=-1, sampling_flags=0, clear=True,
             clear_color=(0.0, 0.0, 0.0, 0.0)):
        cc = np.array(clear_color, "<f4")
        return lib().pfo_frame_draw_batch(self.h, slot, target_page, color_page, sampling_flags, int(clear), _p(cc))

    def pixels(self):
        px = np.zeros((int(self.scene["height"]), int(self.scene["width"]), 4), "u1")
        lib().pfo_frame_pixels(self.h, _p(px))
        return px

    def render(self, clear=True, clear_color=(0.0, 0.0, 0.0, 0.0)):
        """Whole frame the way RendererD3D11::draw sequences it (core/d3d11/renderer.cpp:302-336)."""
        slots = self.prepare_all()
        first = clear
        for b, slot in zip(self.scene["draw_batches"], slots["draw"]):
            info = b["info"]
            color_page = -1 if int(info[7]) == 0xFFFFFFFF else int(info[7])
            flags = 0 if int(info[8]) == 0xFFFFFFFF else int(info[8])
            target = -1 if int(info[10]) == 0xFFFFFFFF else int(info[11])
            if target < 0:
                self.draw(slot, -1, color_page, flags, first, clear_color)
                first = False
            else:
                self.draw(slot, target, color_page, flags, True, (0.0, 0.0, 0.0, 0.0))
        return self.pixels()

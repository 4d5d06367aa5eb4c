"""TEST INFRASTRUCTURE (oracle). The reference's own compute shaders run on the CPU (oracle/_ref/libpfshader.so: fill.comp
and tile.comp compiled by g++ through oracle/ref_harness/glsl_shim.h from the text under /root/reference).

This module turns the taps of oracle/pf_oracle.c's geometry stages -- which are pinned bit-exact to the reference's CPU
tiler -- into the buffers RendererD3D11 hands those shaders (core/d3d11/gpu_data.h:26-52, bin.comp's fill lists,
propagate.comp's alpha tile records) and runs them, so that the PIXEL half of the restatement (masks, composite) can be
pinned against the reference's shader text. Only tests/ may import it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpfshader.so")
MASK_W = 4096  # MASK_FRAMEBUFFER_WIDTH: 256 alpha tiles of 16 texels per row (core/d3d11/renderer.cpp:1053-1056)


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, i = C.c_void_p, C.c_int
        L.pfshader_fill.argtypes = [vp, vp, vp, i, i, vp, i, i, vp, i, i]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def mask_image(n_alpha_total):
    """The mask texture: alpha tile i lives at texel origin ((i & 0xff) * 16, (i >> 8) * 4), 4 pixel rows per texel."""
    pages = max(1, (int(n_alpha_total) + 65535) // 65536)  # MASK_FRAMEBUFFER_HEIGHT = 1024 texel rows per page
    return np.zeros((1024 * pages, MASK_W, 4), "u1")


def fill_buffers(batch, tiles, fills, dense_path):
    """(iFills, iTiles, iAlphaTiles, first_alpha) in the reference's layouts from the oracle's taps of one batch.
    tiles / fills: pforacle.Frame.tiles / fills; dense_path: batch-local path index of every dense tile."""
    n_tiles = len(tiles)
    # fill linked lists (bin.comp:94-110): the order inside a tile's list is irrelevant to the sum's terms, not to its
    # rounding -- the lists are chained in the oracle's (canonical) order
    i_fills = np.zeros((len(fills), 3), "<u4")
    i_fills[:, 0] = fills["from_x"].astype("<u4") | (fills["from_y"].astype("<u4") << 16)
    i_fills[:, 1] = fills["to_x"].astype("<u4") | (fills["to_y"].astype("<u4") << 16)
    first_fill = np.full(n_tiles, -1, "<i4")
    ti = fills["tile_index"]
    nxt = np.full(len(fills), -1, "<i4")
    same = ti[1:] == ti[:-1]
    nxt[:-1][same] = np.arange(1, len(fills))[same]
    starts = np.ones(len(fills), bool)
    starts[1:] = ~same
    first_fill[ti[starts]] = np.nonzero(starts)[0]
    i_fills[:, 2] = nxt.view("<u4")
    tpi = batch["tile_path_info"]
    i_tiles = np.zeros((n_tiles, 4), "<u4")
    i_tiles[:, 0] = 0xFFFFFFFF
    i_tiles[:, 1] = first_fill.view("<u4")
    i_tiles[:, 2] = (tiles["alpha_tile_id"].astype("<i4").view("<u4") & 0x00FFFFFF) | \
                    (tiles["backdrop_delta"].astype("<i4").view("<u4") << 24)
    i_tiles[:, 3] = tpi["color"][dense_path].astype("<u4") | (tpi["ctrl"][dense_path].astype("<u4") << 16) | \
                    (tiles["backdrop"].astype("<i4").view("<u4") << 24)
    own = np.nonzero((tiles["alpha_tile_id"] >= 0) & (tiles["fill_count"] > 0))[0]
    if len(own) == 0:
        return i_fills, i_tiles, np.zeros((0, 2), "<u4"), 0
    ids = tiles["alpha_tile_id"][own]
    first = int(ids.min())
    assert np.array_equal(np.sort(ids), np.arange(first, first + len(own))), "alpha tile ids of a batch are contiguous"
    i_alpha = np.zeros((len(own), 2), "<u4")
    i_alpha[ids - first, 0] = own
    i_alpha[ids - first, 1] = tiles["clip_alpha_tile_id"][own].astype("<i4").view("<u4")
    return i_fills, i_tiles, i_alpha, first


def run_fill(i_fills, i_tiles, i_alpha, first_alpha, area_lut, mask):
    """fill.comp over one batch's alpha tiles; `mask` (mask_image) is read for clips and written."""
    i_fills = np.ascontiguousarray(i_fills, "<u4")
    i_tiles = np.ascontiguousarray(i_tiles, "<u4")
    i_alpha = np.ascontiguousarray(i_alpha, "<u4")
    lut = np.ascontiguousarray(area_lut, "u1")
    if len(i_alpha) == 0:
        return
    if len(i_fills) == 0:
        i_fills = np.zeros((1, 3), "<u4")
    lib().pfshader_fill(_p(i_fills), _p(i_tiles), _p(i_alpha), int(first_alpha), len(i_alpha), _p(lut), lut.shape[1],
                        lut.shape[0], _p(mask), mask.shape[1], mask.shape[0])


def mask_of(mask, alpha_id):
    """16 x 16 coverage bytes of one alpha tile out of the mask texture (channel c of texel row r = pixel row 4 r + c)."""
    x0, y0 = (alpha_id & 0xFF) * 16, (alpha_id >> 8) * 4
    t = mask[y0:y0 + 4, x0:x0 + 16, :]           # [texel row, column, channel]
    return np.transpose(t, (0, 2, 1)).reshape(16, 16)


# ---------------------------------------------------------------------------------------------- tile.comp

MD_W, MD_H = 1280, 512  # TEXTURE_METADATA_TEXTURE_WIDTH x HEIGHT (core/paint/palette.h:11-12)
NONE = 0xFFFFFFFF


def metadata_texels(scene):
    """The RGBA16F paint metadata texture as fp32 texels (MD_H x MD_W x 4), rows beyond the uploaded ones zero."""
    md = np.ascontiguousarray(scene["metadata"], "<u2")
    out = np.zeros((MD_H, MD_W, 4), "<f4")
    rows = md.shape[0]
    out[:rows] = md.view("<f2").astype("<f4").reshape(rows, MD_W, 4)
    return out


def tile_buffers(i_tiles, offsets, sorted_tiles):
    """iTiles with word 0 = next tile of the framebuffer tile's sorted, z-culled list (what sort.comp leaves), and the
    first-tile map. offsets / sorted_tiles: pforacle.Frame.tile_lists (CSR per framebuffer tile, paint order)."""
    t = np.array(i_tiles, "<u4", copy=True)
    t[:, 0] = NONE
    n_fb = len(offsets) - 1
    first = np.full(n_fb, -1, "<i4")
    lens = np.diff(offsets.astype(np.int64))
    has = lens > 0
    first[has] = sorted_tiles[offsets[:-1][has]].astype("<i4")
    if len(sorted_tiles) > 1:
        nxt = np.full(len(sorted_tiles), NONE, "<u4")
        nxt[:-1] = sorted_tiles[1:]
        last = (offsets[1:][has] - 1).astype(np.int64)  # last entry of every non-empty list ends the chain
        nxt[last] = NONE
        t[sorted_tiles, 0] = nxt
    return t, first


def run_tile(i_tiles, first_map, fb_tw, fb_th, md, color, sampling_flags, mask, dest, clear, clear_color):
    L = lib()
    if not getattr(L, "_tile_ready", False):
        vp, i = C.c_void_p, C.c_int
        L.pfshader_tile.argtypes = [vp, vp, i, i, vp, i, i, vp, i, i, C.c_uint32, vp, i, i, vp, i, i, i, vp]
        L._tile_ready = True
    i_tiles = np.ascontiguousarray(i_tiles, "<u4")
    first_map = np.ascontiguousarray(first_map, "<i4")
    color = np.ascontiguousarray(color, "u1")
    cc = np.array(clear_color, "<f4")
    L.pfshader_tile(_p(i_tiles), _p(first_map), fb_tw, fb_th, _p(md), md.shape[1], md.shape[0], _p(color), color.shape[1],
                    color.shape[0], int(sampling_flags), _p(mask), mask.shape[1], mask.shape[0], _p(dest), dest.shape[1],
                    dest.shape[0], 0 if clear else 1, _p(cc))


def render_frame(scene, fr, area_lut, clear_color=(0.0, 0.0, 0.0, 0.0)):
    """The whole frame through the reference's fill.comp + tile.comp, sequenced like RendererD3D11::draw
    (core/d3d11/renderer.cpp:302-336), on the geometry of the oracle frame `fr` (already rendered: its taps are read).
    Returns (destination pixels, {page: pixels}, mask texture)."""
    import scenes  # tests/scenes.py (dense tile -> path)

    total = max([fr.counts(s)["first_alpha"] + fr.counts(s)["alpha_tiles"] for s in fr.slots.values()] + [1])
    mask = mask_image(total)
    md = metadata_texels(scene)
    w, h = int(scene["width"]), int(scene["height"])
    fb_tw, fb_th = (w + 15) // 16, (h + 15) // 16
    dest = np.zeros((h, w, 4), "u1")
    pages = {int(k): np.array(v, "u1", copy=True) for k, v in scene.get("pages", {}).items()}
    dummy = np.zeros((1, 1, 4), "u1")
    order = [("clip", i) for i in reversed(range(len(scene["clip_batches"])))
             if int(scene["clip_batches"][i]["info"][1]) > 0] + [("draw", i) for i in range(len(scene["draw_batches"]))]
    first = True
    for kind, i in order:
        b = scene[kind + "_batches"][i]
        slot = fr.slots[int(b["info"][0])]
        tiles, fills = fr.tiles(slot), fr.fills(slot)
        path = scenes.dense_tile_coords(b)[0]
        i_fills, i_tiles, i_alpha, first_alpha = fill_buffers(b, tiles, fills, path)
        run_fill(i_fills, i_tiles, i_alpha, first_alpha, area_lut, mask)
        if kind == "clip":
            continue
        off, lst = fr.tile_lists(slot)
        t, first_map = tile_buffers(i_tiles, off, lst)
        info = b["info"]
        color = pages[int(info[7])] if int(info[7]) != NONE else dummy
        flags = 0 if int(info[8]) == NONE else int(info[8])
        if int(info[10]) == NONE:
            run_tile(t, first_map, fb_tw, fb_th, md, color, flags, mask, dest, first, clear_color)
            first = False
        else:
            run_tile(t, first_map, fb_tw, fb_th, md, color, flags, mask, pages[int(info[11])], True, (0.0, 0.0, 0.0, 0.0))
    return dest, pages, mask


# ---------------------------------------------------------------------------------------------- propagate.comp + sort.comp

def run_propagate_sort(scene, fr, kind, index, clip_state=None):
    """propagate.comp then sort.comp on one batch, fed with what bin left behind according to the oracle frame `fr`
    (first fill ids, backdrop deltas, column backdrops). Returns a dict: tiles (n x 4 u32 after both shaders), z (per
    framebuffer tile), first_map (sorted list heads), alpha_tiles (k x 2), lists (CSR offsets, dense tile indices walked
    from the sorted linked lists). clip_state: the same dict of the batch's clip batch (its tiles are read)."""
    import scenes

    L = lib()
    if not getattr(L, "_prop_ready", False):
        vp, i = C.c_void_p, C.c_int
        L.pfshader_propagate.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i]
        L.pfshader_sort.argtypes = [vp, vp, vp, i]
        L._prop_ready = True
    b = scene[kind + "_batches"][index]
    slot = fr.slots[int(b["info"][0])]
    tiles, fills = fr.tiles(slot), fr.fills(slot)
    path = scenes.dense_tile_coords(b)[0]
    i_fills, i_tiles, _, _ = fill_buffers(b, tiles, fills, path)
    # as bound.comp + bin.comp leave the tiles: no alpha tile yet (-1 in the low 24 bits), backdrop 0 in the control word
    i_tiles[:, 2] = 0x00FFFFFF | (tiles["backdrop_delta"].astype("<i4").view("<u4") << 24)
    i_tiles[:, 3] &= 0x00FFFFFF
    w, h = int(scene["width"]), int(scene["height"])
    fb_tw, fb_th = (w + 15) // 16, (h + 15) // 16
    z = np.zeros(8 + fb_tw * fb_th, "<i4")
    first_map = np.full(fb_tw * fb_th, -1, "<i4")
    alpha = np.zeros((max(len(tiles), 1), 2), "<u4")
    md = np.ascontiguousarray(b["propagate_metadata"]).view("<u4").reshape(-1)
    bd = np.ascontiguousarray(b["backdrops"]).view("<i4").reshape(-1, 3).copy()
    bd[:, 0] = fr.column_backdrops(slot)
    first_alpha = fr.counts(slot)["first_alpha"]
    clip_md = clip_tiles = None
    if clip_state is not None:
        clip_md, clip_tiles = clip_state["metadata_words"], clip_state["tiles"]
    L.pfshader_propagate(_p(md), _p(clip_md) if clip_md is not None else None, _p(bd), _p(i_tiles),
                         _p(clip_tiles) if clip_tiles is not None else None, _p(z), _p(first_map), _p(alpha), fb_tw, fb_th,
                         len(bd), first_alpha)
    L.pfshader_sort(_p(i_tiles), _p(first_map), _p(z), fb_tw * fb_th)
    offsets, lst = [0], []
    for m in range(fb_tw * fb_th):
        t = int(first_map[m])
        while t >= 0:
            lst.append(t)
            t = int(np.int32(i_tiles[t, 0]))
        offsets.append(len(lst))
    return dict(tiles=i_tiles, z=z[8:], n_alpha=int(z[4]), first_map=first_map, alpha_tiles=alpha[:int(z[4])],
                lists=(np.array(offsets, "<u4"), np.array(lst, "<u4")), metadata_words=md, first_alpha=first_alpha)


# ---------------------------------------------------------------------------------------------- the whole GPU-driven mode

def render_frame_gpu_driven(scene, area_lut, clear_color=(0.0, 0.0, 0.0, 0.0)):
    """The reference's GPU-driven mode END TO END on the CPU, from the scene builder's vectors alone: dice.comp, bound.comp,
    bin.comp, propagate.comp, fill.comp, sort.comp and tile.comp in RendererD3D11::draw's order
    (core/d3d11/renderer.cpp:302-336, :510-616). No part of oracle/pf_oracle.c is involved. This is the render the
    reference's own shaders produce -- with THEIR flattening (uniform-t microlines, dice.comp:176-222) and THEIR fixed
    point (truncation, bin.comp:94), i.e. not the hybrid tiler's geometry that the CUDA path is bit-exact with.
    Returns (destination pixels, stats)."""
    L = lib()
    if not getattr(L, "_geo_ready", False):
        vp, i = C.c_void_p, C.c_int
        L.pfshader_dice.argtypes = [vp, vp, vp, i, vp, vp, vp, i, i, i]
        L.pfshader_bound.argtypes = [vp, vp, i, i]
        L.pfshader_bin.argtypes = [vp, vp, vp, vp, vp, vp, i, i]
        L.pfshader_propagate.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i]
        L.pfshader_sort.argtypes = [vp, vp, vp, i]
        L._geo_ready = L._prop_ready = True
    w, h = int(scene["width"]), int(scene["height"])
    fb_tw, fb_th = (w + 15) // 16, (h + 15) // 16
    n_fb = fb_tw * fb_th
    md = metadata_texels(scene)
    dest = np.zeros((h, w, 4), "u1")
    pages = {int(k): np.array(v, "u1", copy=True) for k, v in scene.get("pages", {}).items()}
    dummy = np.zeros((1, 1, 4), "u1")
    mask = mask_image(1 << 16)
    state, stats = {}, dict(microlines=0, fills=0, alpha_tiles=0)
    next_alpha = 0
    order = [("clip", i) for i in reversed(range(len(scene["clip_batches"])))
             if int(scene["clip_batches"][i]["info"][1]) > 0] + [("draw", i) for i in range(len(scene["draw_batches"]))]
    first = True
    for kind, bi in order:
        b = scene[kind + "_batches"][bi]
        info = b["info"]
        n_paths, n_tiles, n_segs = int(info[1]), int(info[2]), int(info[3])
        src = "clip" if int(info[5]) else "draw"
        pts = np.ascontiguousarray(scene[src + "_points"], "<f4")
        idx = np.ascontiguousarray(scene[src + "_indices"], "<u4")
        dice_md = np.ascontiguousarray(b["dice_metadata"]).view("<u4").reshape(-1)
        prop_md = np.ascontiguousarray(b["propagate_metadata"]).view("<u4").reshape(-1)
        tpi = np.ascontiguousarray(b["tile_path_info"]).view("<u4").reshape(-1)
        transform = np.ascontiguousarray(b["transform"], "<f4")
        cap = max(16384, n_segs * 8)
        while True:  # dice_segments retries with a larger buffer (renderer.cpp:537-556)
            indirect = np.zeros(8, "<u4")
            microlines = np.zeros((cap, 4), "<u4")
            L.pfshader_dice(_p(indirect), _p(dice_md), _p(pts), len(pts), _p(idx), _p(microlines), _p(transform), n_paths,
                            n_segs, cap)
            if int(indirect[3]) <= cap:
                break
            cap = int(indirect[3])
        n_micro = int(indirect[3])
        tiles = np.zeros((max(n_tiles, 1), 4), "<u4")
        L.pfshader_bound(_p(tpi), _p(tiles), n_paths, n_tiles)
        backdrops = np.ascontiguousarray(b["backdrops"]).view("<u4").reshape(-1, 3).copy()
        fcap = max(65536, n_micro * 4)
        while True:
            z = np.zeros(8 + n_fb, "<i4")
            t2, bd2 = tiles.copy(), backdrops.copy()
            fills = np.zeros((fcap, 3), "<u4")
            L.pfshader_bin(_p(microlines), _p(prop_md.view("<i4")), _p(z.view("<u4")), _p(fills), _p(t2), _p(bd2), n_micro, fcap)
            if int(z[1]) <= fcap:
                break
            fcap = int(z[1])
        tiles, backdrops = t2, bd2
        first_map = np.full(n_fb, -1, "<i4")
        alpha = np.zeros((max(n_tiles, 1), 2), "<u4")
        clip_md = clip_tiles = None
        cb = int(info[6])
        if cb != NONE and cb in state:
            clip_md, clip_tiles = state[cb]["metadata"], state[cb]["tiles"]
        L.pfshader_propagate(_p(prop_md), _p(clip_md) if clip_md is not None else None, _p(backdrops.view("<i4")),
                             _p(tiles), _p(clip_tiles) if clip_tiles is not None else None, _p(z), _p(first_map), _p(alpha),
                             fb_tw, fb_th, len(backdrops), next_alpha)
        n_alpha = int(z[4])
        run_fill(fills, tiles, alpha[:n_alpha], next_alpha, area_lut, mask)
        L.pfshader_sort(_p(tiles), _p(first_map), _p(z), n_fb)
        state[int(info[0])] = dict(metadata=prop_md, tiles=tiles)
        stats["microlines"] += n_micro
        stats["fills"] += int(z[1])
        stats["alpha_tiles"] += n_alpha
        next_alpha += n_alpha
        if kind == "clip":
            continue
        color = pages[int(info[7])] if int(info[7]) != NONE else dummy
        flags = 0 if int(info[8]) == NONE else int(info[8])
        if int(info[10]) == NONE:
            run_tile(tiles, first_map, fb_tw, fb_th, md, color, flags, mask, dest, first, clear_color)
            first = False
        else:
            run_tile(tiles, first_map, fb_tw, fb_th, md, color, flags, mask, pages[int(info[11])], True, (0.0, 0.0, 0.0, 0.0))
    return dest, stats

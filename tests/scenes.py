"""Scene containers, golden-fixture IO and canonical forms shared by the tests and bench.py.

A *scene dict* holds exactly what the reference's SceneBuilderD3D11 hands to its renderer
(core/d3d11/scene_builder.h:50-55, core/d3d11/gpu_data.h:97-189):

    width, height, view_box[4]
    draw_points (n,2) f32, draw_indices (m,2) u32, clip_points, clip_indices
    draw_batches / clip_batches: list of dicts
        info[16] u32  (batch_id, path_count, tile_count, segment_count, n_backdrops, path_source, clip_batch_id|~0,
                       color page|~0, sampling flags, composite op, render target|~0, rt page|~0, rt rect[4])
        backdrops, propagate_metadata, dice_metadata, tile_path_info (structured, see DTYPES), transform[6]
    metadata (rows, 5120) u16   RGBA16F paint metadata rows (core/renderer.cpp:167-251)
    pages {page: (h, w, 4) u8}  gradient / image pages

Nothing in this module touches oracle/ or the CUDA library; it is plain numpy.
"""
import hashlib
import io
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

BACKDROP_DT = np.dtype([("initial_backdrop", "<i4"), ("tile_x_offset", "<i4"), ("path_index", "<u4")])
PROPAGATE_DT = np.dtype([("rect", "<i4", (4,)), ("tile_offset", "<u4"), ("path_index", "<u4"), ("z_write", "<u4"),
                         ("clip_path_index", "<u4"), ("backdrop_offset", "<u4"), ("pad", "<u4", (3,))])
DICE_DT = np.dtype([("global_path_id", "<u4"), ("first_global_segment_index", "<u4"),
                    ("first_batch_segment_index", "<u4"), ("pad", "<u4")])
TILE_PATH_INFO_DT = np.dtype([("tile_min_x", "<i2"), ("tile_min_y", "<i2"), ("tile_max_x", "<i2"),
                              ("tile_max_y", "<i2"), ("first_tile_index", "<u4"), ("color", "<u2"), ("ctrl", "u1"),
                              ("backdrop", "i1")])
DTYPES = dict(backdrops=BACKDROP_DT, propagate_metadata=PROPAGATE_DT, dice_metadata=DICE_DT,
              tile_path_info=TILE_PATH_INFO_DT)
NONE = 0xFFFFFFFF

CURVE_IS_QUADRATIC = 0x80000000
CURVE_IS_CUBIC = 0x40000000

# canonical records
CANON_TILE_DT = np.dtype([("path", "<u4"), ("tile_x", "<i2"), ("tile_y", "<i2"), ("ctrl", "u1"), ("backdrop", "i1"),
                          ("paint", "<u2"), ("has_alpha", "u1")])
CANON_FILL_DT = np.dtype([("path", "<u4"), ("tile_x", "<i2"), ("tile_y", "<i2"), ("from_x", "<u2"), ("from_y", "<u2"),
                          ("to_x", "<u2"), ("to_y", "<u2")])
CANON_CLIP_DT = np.dtype([("path", "<u4"), ("tile_x", "<i2"), ("tile_y", "<i2"), ("dest_backdrop", "<i4"),
                          ("src_backdrop", "<i4"), ("src_group", "<u8")])


# ------------------------------------------------------------------------------------------------ fixture IO

def save_scene(path, scene, extra=None):
    flat = {"width": np.int32(scene["width"]), "height": np.int32(scene["height"]),
            "view_box": np.asarray(scene["view_box"], "<f4")}
    for k in ("draw_points", "draw_indices", "clip_points", "clip_indices", "metadata"):
        flat[k] = scene[k]
    for page, px in scene.get("pages", {}).items():
        flat["page_%d" % page] = px
    for kind in ("draw", "clip"):
        batches = scene[kind + "_batches"]
        flat["n_%s_batches" % kind] = np.int32(len(batches))
        for i, b in enumerate(batches):
            for k, v in b.items():
                flat["%s%d_%s" % (kind, i, k)] = v
    for k, v in (extra or {}).items():
        flat["x_" + k] = v
    np.savez_compressed(path, **flat)


def load_scene(path):
    z = np.load(path)
    scene = {"width": int(z["width"]), "height": int(z["height"]), "view_box": z["view_box"], "pages": {}}
    for k in ("draw_points", "draw_indices", "clip_points", "clip_indices", "metadata"):
        scene[k] = z[k]
    extra = {}
    for k in z.files:
        if k.startswith("page_"):
            scene["pages"][int(k[5:])] = z[k]
        elif k.startswith("x_"):
            extra[k[2:]] = z[k]
    for kind in ("draw", "clip"):
        batches = []
        for i in range(int(z["n_%s_batches" % kind])):
            prefix = "%s%d_" % (kind, i)
            batches.append({k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)})
        scene[kind + "_batches"] = batches
    return scene, extra


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


# ------------------------------------------------------------------------------------------------ canonical forms

def _group_hash(rows):
    """Order-independent identity of one tile's fill list: hash of its sorted (fx, fy, tx, ty) rows."""
    a = np.ascontiguousarray(rows, "<u2").reshape(-1, 4)
    a = a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest()[:8], "<u8")[0]


def _sorted(a):
    return np.sort(a, order=list(a.dtype.names)) if len(a) else a


def canonical_from_reference(d9):
    """Order/ID-independent form of SceneBuilderD3D9's output (oracle/pfref.py RefScene.build_d3d9).

    Alpha tile ids and fill order race between the reference's 4 worker threads
    (core/d3d9/data/alpha_tile_id.cpp:5-10, scene_builder.cpp:293-298), so fills are regrouped by the tile that
    owns them and every list is sorted.
    """
    fills = d9["fills"]
    order = np.argsort(fills["link"], kind="stable")
    fs = fills[order]
    links, starts = np.unique(fs["link"], return_index=True)
    ends = np.append(starts[1:], len(fs))
    quad = np.stack([fs["from_x"], fs["from_y"], fs["to_x"], fs["to_y"]], axis=1)
    group_of = {}
    hashes = np.zeros(len(links), "<u8")
    for i, (l, s, e) in enumerate(zip(links, starts, ends)):
        hashes[i] = _group_hash(quad[s:e])
        group_of[int(l)] = (s, e, hashes[i])
    out = {"group_hashes": np.sort(hashes), "batches": []}
    for b in d9["batches"]:
        t = b["tiles"]
        ct = np.zeros(len(t), CANON_TILE_DT)
        ct["path"], ct["tile_x"], ct["tile_y"] = t["path_id"], t["tile_x"], t["tile_y"]
        ct["ctrl"], ct["backdrop"], ct["paint"] = t["ctrl"], t["backdrop"], t["metadata_id"]
        ct["has_alpha"] = t["alpha_tile_id"] != NONE
        # keyed fills: every alpha tile referenced by exactly one listed tile that owns its fills
        keyed = []
        alpha_ids, counts = np.unique(t["alpha_tile_id"][t["alpha_tile_id"] != NONE], return_counts=True)
        multi = set(int(a) for a, c in zip(alpha_ids, counts) if c > 1)
        owner = {}
        for rec in t[t["alpha_tile_id"] != NONE]:
            owner[int(rec["alpha_tile_id"])] = rec
        clips = np.zeros(len(b["clips"]), CANON_CLIP_DT)
        clip_src = set(int(c["src_tile_id"]) for c in b["clips"])
        for i, c in enumerate(b["clips"]):
            rec = owner.get(int(c["dest_tile_id"]))
            if rec is not None:
                clips[i]["path"], clips[i]["tile_x"], clips[i]["tile_y"] = rec["path_id"], rec["tile_x"], rec["tile_y"]
            clips[i]["dest_backdrop"], clips[i]["src_backdrop"] = c["dest_backdrop"], c["src_backdrop"]
            g = group_of.get(int(c["src_tile_id"]))
            clips[i]["src_group"] = g[2] if g else 0
        for a, rec in owner.items():
            if a in multi or a in clip_src or a not in group_of:
                continue
            s, e, _ = group_of[a]
            k = np.zeros(e - s, CANON_FILL_DT)
            k["path"], k["tile_x"], k["tile_y"] = rec["path_id"], rec["tile_x"], rec["tile_y"]
            k["from_x"], k["from_y"], k["to_x"], k["to_y"] = quad[s:e].T
            keyed.append(k)
        keyed = np.concatenate(keyed) if keyed else np.zeros(0, CANON_FILL_DT)
        out["batches"].append(dict(tiles=_sorted(ct), fills=_sorted(keyed), clips=_sorted(clips),
                                   z=np.asarray(b["z"], "<u4").reshape(-1)))
    return out


def dense_tile_coords(batch):
    """(local path index, tile_x, tile_y) of every dense tile of a batch, in dense order."""
    meta = batch["propagate_metadata"]
    n = int(batch["info"][2])
    path = np.zeros(n, "<u4")
    tx = np.zeros(n, "<i4")
    ty = np.zeros(n, "<i4")
    for p, m in enumerate(meta):
        x0, y0, x1, y1 = (int(v) for v in m["rect"])
        w, h = x1 - x0, y1 - y0
        if w <= 0 or h <= 0:
            continue
        o = int(m["tile_offset"])
        path[o:o + w * h] = p
        tx[o:o + w * h] = np.tile(np.arange(x0, x1), h)
        ty[o:o + w * h] = np.repeat(np.arange(y0, y1), w)
    return path, tx, ty


def canonical_from_taps(scene, batch, tiles, fills, clip_batch=None, clip_tiles=None, clip_fills=None):
    """Same canonical form from the dense taps that both the C oracle and the CUDA path expose.

    tiles: per dense tile records with fields alpha_tile_id, clip_alpha_tile_id, fill_count, backdrop,
    backdrop_d3d9, listed. fills: records with tile_index, from_x.. to_y.
    Returns dict(tiles, fills, clips, group_hashes) for one batch.
    """
    path, tx, ty = dense_tile_coords(batch)
    gid = batch["dice_metadata"]["global_path_id"][path]
    tpi = batch["tile_path_info"]
    listed = tiles["listed"] != 0
    # the hybrid builder keeps a tile iff it has an alpha tile or a non-zero backdrop (d3d9/scene_builder.cpp:57-61)
    ct = np.zeros(int(listed.sum()), CANON_TILE_DT)
    ct["path"], ct["tile_x"], ct["tile_y"] = gid[listed], tx[listed], ty[listed]
    ct["ctrl"], ct["paint"] = tpi["ctrl"][path[listed]], tpi["color"][path[listed]]
    ct["backdrop"] = tiles["backdrop_d3d9"][listed]
    ct["has_alpha"] = tiles["alpha_tile_id"][listed] >= 0
    quad = np.stack([fills["from_x"], fills["from_y"], fills["to_x"], fills["to_y"]], axis=1)
    ti = fills["tile_index"]
    order = np.argsort(ti, kind="stable")
    ti, quad = ti[order], quad[order]
    utile, starts = np.unique(ti, return_index=True)
    ends = np.append(starts[1:], len(ti))
    # the GPU-driven builder does not clip a clip path's tile rect to the view box (d3d11/scene_builder.cpp:80 "TODO"),
    # the hybrid tiler does (tiler.cpp:327): tiles outside the view box exist only on the GPU-driven side (they feed the
    # column backdrops, which ARE compared) and are left out of the fill-group comparison
    vb = np.asarray(scene["view_box"], "f4")
    vx0, vy0 = int(np.floor(vb[0] / 16)), int(np.floor(vb[1] / 16))
    vx1, vy1 = int(np.ceil(vb[2] / 16)), int(np.ceil(vb[3] / 16))
    hashes = np.array([_group_hash(quad[s:e]) for t, s, e in zip(utile, starts, ends)
                       if vx0 <= tx[t] < vx1 and vy0 <= ty[t] < vy1], "<u8")
    # keyed fills: tiles that own a mask of their own and are listed
    first_own = None
    own = (tiles["alpha_tile_id"] >= 0) & (tiles["fill_count"] > 0) & listed
    sel = own[ti]
    k = np.zeros(int(sel.sum()), CANON_FILL_DT)
    k["path"], k["tile_x"], k["tile_y"] = gid[ti[sel]], tx[ti[sel]], ty[ti[sel]]
    k["from_x"], k["from_y"], k["to_x"], k["to_y"] = quad[sel].T
    clips = np.zeros(0, CANON_CLIP_DT)
    if clip_batch is not None:
        has_clip = tiles["clip_alpha_tile_id"] >= 0
        cpath, ctx, cty = dense_tile_coords(clip_batch)
        # map clip alpha id -> clip dense tile
        cidx = {int(a): i for i, a in enumerate(clip_tiles["alpha_tile_id"]) if a >= 0 and clip_tiles["fill_count"][i] > 0}
        cq = np.stack([clip_fills["from_x"], clip_fills["from_y"], clip_fills["to_x"], clip_fills["to_y"]], axis=1)
        cti = clip_fills["tile_index"]
        clips = np.zeros(int(has_clip.sum()), CANON_CLIP_DT)
        for j, i in enumerate(np.nonzero(has_clip)[0]):
            clips[j]["path"], clips[j]["tile_x"], clips[j]["tile_y"] = gid[i], tx[i], ty[i]
            clips[j]["dest_backdrop"] = tiles["backdrop"][i]
            c = cidx[int(tiles["clip_alpha_tile_id"][i])]
            clips[j]["src_backdrop"] = clip_tiles["backdrop_d3d9"][c]
            clips[j]["src_group"] = _group_hash(cq[cti == c])
        # solid draw tile x alpha clip tile: the draw tile points straight at the clip's mask (tiler.cpp:417-422), so
        # the reference form keys the clip tile's fills by the draw tile when that mask has exactly one such user and
        # is not also the source of a mask combine in this batch
        borrowed = np.nonzero((tiles["alpha_tile_id"] >= 0) & (tiles["fill_count"] == 0) & listed)[0]
        if len(borrowed):
            ids, cnt = np.unique(tiles["alpha_tile_id"][listed & (tiles["alpha_tile_id"] >= 0)], return_counts=True)
            users = dict(zip(ids.tolist(), cnt.tolist()))
            combine_src = set(tiles["clip_alpha_tile_id"][has_clip].tolist())
            extra = []
            for i in borrowed:
                a = int(tiles["alpha_tile_id"][i])
                if users.get(a, 0) != 1 or a in combine_src or a not in cidx:
                    continue
                q = cq[cti == cidx[a]]
                e = np.zeros(len(q), CANON_FILL_DT)
                e["path"], e["tile_x"], e["tile_y"] = gid[i], tx[i], ty[i]
                e["from_x"], e["from_y"], e["to_x"], e["to_y"] = q.T
                extra.append(e)
            if extra:
                k = np.concatenate([k] + extra)
    return dict(tiles=_sorted(ct), fills=_sorted(k), clips=_sorted(clips), group_hashes=np.sort(hashes))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------ synthetic scenes

class PCG32:
    """PCG-XSH-RR 64/32 (O'Neill 2014), the generator SURVEY.md section 8d names for config 4."""

    MULT = 6364136223846793005

    def __init__(self, seed, seq=0xda3e39cb94b95bdb):
        self.mask = (1 << 64) - 1
        self.state = 0
        self.inc = ((seq << 1) | 1) & self.mask
        self.next()
        self.state = (self.state + seed) & self.mask
        self.next()

    def next(self):
        old = self.state
        self.state = (old * self.MULT + self.inc) & self.mask
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def block(self, n):
        """The next n outputs at once (same values as n calls of next()): the LCG is jumped ahead in closed form,
        state_k = a^k * s + inc * (1 + a + ... + a^(k-1)), all in wrapping uint64 arithmetic."""
        with np.errstate(over="ignore"):
            a_pow = np.cumprod(np.concatenate(([np.uint64(1)], np.full(n, self.MULT, np.uint64))), dtype=np.uint64)
            geo = np.concatenate(([np.uint64(0)], np.cumsum(a_pow[:-1], dtype=np.uint64)))
            states = a_pow * np.uint64(self.state) + np.uint64(self.inc) * geo
        old = states[:-1]
        self.state = int(states[-1])
        xorshifted = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)) & np.uint64(0xFFFFFFFF)
        rot = old >> np.uint64(59)
        out = (xorshifted >> rot) | (xorshifted << ((np.uint64(32) - rot) & np.uint64(31)))
        return (out & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _half_bits(x):
    return np.asarray(x, "<f4").astype("<f2").view("<u2")


def solid_metadata(colors_rgba8):
    """RGBA16F metadata rows for solid paints, laid out as core/renderer.cpp:185-236 does:
    10 texels per paint, 128 paints per 1280-texel row; identity colour transform, base colour, ctrl = 0."""
    n = len(colors_rgba8)
    rows = max((n + 127) // 128, 1)
    md = np.zeros((rows, 1280 * 4), "<u2")
    one = _half_bits(1.0)
    col = _half_bits(np.asarray(colors_rgba8, "<f4") / 255.0)
    for i in range(n):
        r, c = divmod(i, 128)
        base = c * 40
        md[r, base + 0] = one  # m11
        md[r, base + 3] = one  # m22
        md[r, base + 8:base + 12] = col[i]
    return md


def build_scene_from_outlines(width, height, paths, colors_rgba8, strip=None, drop_invisible=True):
    """Host-side mirror of SceneBuilderD3D11::build for pre-flattened outlines (one draw batch, no clips).

    paths: list of dicts {contours: [ (points (k,2) f32, flags (k,) u8 [0 on-curve, 1 ctrl0, 2 ctrl1]) ],
                          paint: int, fill_rule: 0 winding | 1 even-odd, opaque: bool}
    Restates SegmentsD3D11::add_path (core/d3d11/gpu_data.cpp:79-115), prepare_draw_path_for_gpu_binning
    (core/d3d11/scene_builder.cpp:26-58), BuiltPath tile bounds (core/data/built_path.cpp:8-34,
    core/data/data.h:53-55) and TileBatchDataD3D11::push (core/d3d11/gpu_data.cpp:24-77).
    strip = (y0, y1), y0 a multiple of 16: render only that horizontal strip. The geometry is NOT
    moved: the view box becomes [0, y0, W, y1] and the framebuffer (width x (y1 - y0)) starts at scene tile row y0 / 16
    ("origin_tiles"), so every float operation is the one the full-canvas frame performs and the strip is bit-identical
    to rows [y0, y1) of it.
    """
    y_top = 0.0
    origin_tiles = (0, 0)
    if strip is not None:
        assert int(strip[0]) % 16 == 0
        y_top = float(strip[0])
        origin_tiles = (0, int(strip[0]) // 16)
        height = int(strip[1] - strip[0])
    vb_right = float(int(np.ceil(width / 16.0)) * 16)  # Scene::set_view_box (core/scene.cpp:170-179)
    view_box = np.array([0.0, y_top, vb_right, y_top + float(height)], "<f4")
    points, indices = [], []
    n_points = 0
    backdrops, meta, dice, tpi = [], [], [], []
    tile_count = seg_count = 0
    for gid, p in enumerate(paths):
        first_seg = len(indices)
        lo = np.array([np.inf, np.inf], "<f4")
        hi = np.array([-np.inf, -np.inf], "<f4")
        for pts, flags in p["contours"]:
            pts = np.asarray(pts, "<f4")
            k = len(pts)
            fl = np.asarray(flags, "u1")
            nxt1 = np.append(fl[1:], 0)
            nxt2 = np.append(fl[2:], [0, 0])[:k]
            on_curve = np.nonzero(fl == 0)[0]
            seg_flag = np.where(nxt1[on_curve] == 1, np.where(nxt2[on_curve] == 2, CURVE_IS_CUBIC, CURVE_IS_QUADRATIC), 0)
            indices.extend(zip((n_points + on_curve).tolist(), seg_flag.tolist()))
            points.append(pts)
            points.append(pts[:1])
            n_points += k + 1
            ok = pts[np.isfinite(pts).all(axis=1)]  # test scenes may carry non-finite points; keep bounds finite
            if len(ok):
                lo = np.minimum(lo, ok.min(axis=0))
                hi = np.maximum(hi, ok.max(axis=0))
        n_seg = len(indices) - first_seg
        # outline.bounds ∩ view box (Rect::intersection, common/math/rect.h:160-173); {} when disjoint
        if not np.isfinite(lo).all() or lo[0] > view_box[2] or hi[0] < view_box[0] or lo[1] > view_box[3] or \
                hi[1] < view_box[1]:
            if drop_invisible:
                # prepare_draw_path_for_gpu_binning returns nullptr (core/d3d11/scene_builder.cpp:36-42): the path is
                # not part of the batch; its segments stay in the scene's point / index arrays, unreferenced
                continue
            bounds = np.zeros(4, "<f4")
        else:
            bounds = np.array([max(lo[0], view_box[0]), max(lo[1], view_box[1]), min(hi[0], view_box[2]),
                               min(hi[1], view_box[3])], "<f4")
        scaled = bounds * np.float32(1.0 / 16.0)
        rect = [int(np.floor(scaled[0])), int(np.floor(scaled[1])), int(np.ceil(scaled[2])), int(np.ceil(scaled[3]))]
        w, h = rect[2] - rect[0], rect[3] - rect[1]
        path_index = len(meta)
        ctrl = 0x2 if p.get("fill_rule", 0) == 1 else 0x1
        meta.append((rect, tile_count, path_index, 1 if p.get("opaque", True) else 0, NONE, len(backdrops), (0, 0, 0)))
        backdrops.extend((0, x, path_index) for x in range(w))
        dice.append((gid, first_seg, seg_count, 0))  # first_seg: global (scene) index, seg_count: batch-local
        tpi.append((rect[0], rect[1], rect[2], rect[3], tile_count, p["paint"], ctrl, 0))
        tile_count += w * h
        seg_count += n_seg
    batch = {
        "info": np.array([0, len(meta), tile_count, seg_count, len(backdrops), 0] + [NONE] * 10, "<u4"),
        "backdrops": np.array(backdrops, BACKDROP_DT) if backdrops else np.zeros(0, BACKDROP_DT),
        "propagate_metadata": np.array(meta, PROPAGATE_DT),
        "dice_metadata": np.array(dice, DICE_DT),
        "tile_path_info": np.array(tpi, TILE_PATH_INFO_DT),
        "transform": np.array([1, 0, 0, 1, 0, 0], "<f4"),
    }
    return {
        "width": int(width), "height": int(height), "view_box": view_box,
        "draw_points": np.concatenate(points).astype("<f4") if points else np.zeros((0, 2), "<f4"),
        "draw_indices": np.array(indices, "<u4").reshape(-1, 2),
        "clip_points": np.zeros((0, 2), "<f4"), "clip_indices": np.zeros((0, 2), "<u4"),
        "draw_batches": [batch], "clip_batches": [],
        "metadata": solid_metadata(colors_rgba8), "pages": {},
        "origin_tiles": origin_tiles,
    }


def synthetic_paths(n_paths, size, seed=0x5EED5EED, n_colors=4096):
    """closed blobs of K~U{8..16} cubic segments, radius U[6,20] px.
    Draw order per path: K; centre x, y; radius; K radial jitters; then per segment two control points, each two
    Box-Muller normals (u1, u2 per normal). The stream is generated in blocks and consumed with a cursor."""
    rng = PCG32(seed)
    draws = rng.block(n_paths * (4 + 9 * 16))
    pos = 0
    paths = []
    two_pi = 2.0 * np.pi
    for i in range(n_paths):
        k = 8 + int(draws[pos]) % 9
        u = draws[pos + 1:pos + 4 + 9 * k].astype(np.float64) / 4294967296.0
        pos += 4 + 9 * k
        cx, cy = 32.0 + u[0] * (size - 64.0), 32.0 + u[1] * (size - 64.0)
        r = 6.0 + u[2] * 14.0
        ang = two_pi * np.arange(k) / k
        rr = r * (0.6 + 0.8 * u[3:3 + k])
        on = np.stack([cx + rr * np.cos(ang), cy + rr * np.sin(ang)], axis=1)
        nrm = u[3 + k:].reshape(k, 4, 2)  # per segment: (c1.x, c1.y, c2.x, c2.y) x (u1, u2)
        g = np.sqrt(-2.0 * np.log(np.maximum(nrm[..., 0], 1e-12))) * np.cos(two_pi * nrm[..., 1]) * (0.25 * r)
        p0, p3 = on, np.roll(on, -1, axis=0)
        c1 = p0 + (p3 - p0) / 3.0 + g[:, 0:2]
        c2 = p0 + (p3 - p0) * 2.0 / 3.0 + g[:, 2:4]
        pts = np.empty((3 * k + 1, 2), "<f4")
        pts[0:3 * k:3], pts[1:3 * k:3], pts[2:3 * k:3], pts[3 * k] = p0, c1, c2, on[0]
        flags = np.zeros(3 * k + 1, "u1")
        flags[1:3 * k:3], flags[2:3 * k:3] = 1, 2
        paths.append({"contours": [(pts, flags)], "paint": i % n_colors, "fill_rule": 0, "opaque": True})
    crng = PCG32(seed ^ 0xC0105)
    colors = np.concatenate([crng.block(3 * n_colors).reshape(n_colors, 3) & 255,
                             np.full((n_colors, 1), 255, np.uint32)], axis=1).astype("u1")
    return paths, colors


def synthetic_scene(n_paths, size, seed=0x5EED5EED, strip=None):
    paths, colors = synthetic_paths(n_paths, size, seed)
    return build_scene_from_outlines(size, size, paths, colors, strip=strip)


CONFIG4 = (200000, 8192)  # BASELINE.json configs[3] / SURVEY.md section 8d config 4: paths, canvas size


def tile_sums(px):
    """Channel sums of every 16 x 16 tile of an RGBA8 frame: (tiles_y, tiles_x, 4) u16 (256 * 255 < 65536). The pixel half
    of the fixtures that are too large to commit as frames (tests/golden/synthetic_200k_8192.npz)."""
    h, w, _ = px.shape
    return px.reshape(h // 16, 16, w // 16, 16, 4).astype(np.uint32).sum(axis=(1, 3)).astype("<u2")


# ---- paints_512 with its image paint switched to one of the two colour filters upstream never emits

def _half_bits(values):
    return np.asarray(values, "<f4").astype("<f2").view("<u2")


def refiltered_paints_scene(ctrl, params):
    """paints_512 with its image-pattern paint (entry 23: SRC_IN over pattern page 1) switched to another colour filter: ctrl
    = composite << 10 | combine << 8 | filter << 4 (tile.comp:96-121), params = filterParams0 .. 4 (texels 3 .. 7 of the
    paint's metadata entry, tile.comp:707-717)."""
    scene, _ = load_scene(golden_path("paints_512"))
    scene = dict(scene)
    md = np.array(scene["metadata"], "<u2", copy=True).reshape(-1, 1280, 4)
    entry = 23
    assert int(md[0, entry * 10 + 8].view("<f2")[0]) == 0x100
    for k, v in enumerate(params):
        md[0, entry * 10 + 3 + k] = _half_bits(v)
    md[0, entry * 10 + 8, 0] = _half_bits([float(ctrl)])[0]
    scene["metadata"] = md.reshape(scene["metadata"].shape)
    return scene


TEXT_KERNEL_WIDE = [0.033165660, 0.102074051, 0.221434336, 0.286651906]  # a 9-tap kernel (kernel.x > 0), defringing on
TEXT_KERNEL_NARROW = [0.0, 0.031372549, 0.301960784, 0.337254902]
FILTER_CASES = {
    # mat4 columns + offset: channels swapped and scaled, some of the alpha mixed in
    "color_matrix": (0x140, [[0.1, 0.7, 0.0, 0.0], [0.8, 0.1, 0.2, 0.0], [0.0, 0.2, 0.6, 0.0], [0.05, 0.0, 0.1, 0.9],
                             [0.02, 0.0, 0.05, 0.1]]),
    "text_defringe_wide": (0x120, [TEXT_KERNEL_WIDE, [0.9, 0.85, 0.8, 0.0], [0.1, 0.15, 0.3, 0.0], [0.0] * 4, [0.0] * 4]),
    "text_defringe_narrow": (0x120, [TEXT_KERNEL_NARROW, [1.0, 1.0, 1.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0] * 4, [0.0] * 4]),
    "text_plain": (0x120, [[0.0, 0.0, 0.0, 0.0], [0.2, 0.3, 0.4, 0.0], [1.0, 0.9, 0.1, 0.0], [0.0] * 4, [0.0] * 4]),
}

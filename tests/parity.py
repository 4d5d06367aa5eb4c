"""Canonical-form comparison used for BOTH the C oracle (pinning it to the reference) and the CUDA path."""
import numpy as np

import scenes


def batch_order(scene):
    """(kind, index) in submission order: clip batches reversed, then draw batches (d3d11/renderer.cpp:318-332)."""
    order = [("clip", i) for i in reversed(range(len(scene["clip_batches"])))
             if int(scene["clip_batches"][i]["info"][1]) > 0]
    return order + [("draw", i) for i in range(len(scene["draw_batches"]))]


def canonical_scene(scene, taps):
    """taps: {batch_id: (tiles, fills)} dense taps per prepared batch. Returns the reference-comparable form."""
    by_id = {}
    for kind in ("clip", "draw"):
        for b in scene[kind + "_batches"]:
            by_id[int(b["info"][0])] = b
    out = {"batches": [], "group_hashes": []}
    for b in scene["draw_batches"]:
        bid = int(b["info"][0])
        tiles, fills = taps[bid]
        cb_id = int(b["info"][6])
        if cb_id != scenes.NONE and cb_id in taps:
            ct, cf = taps[cb_id]
            c = scenes.canonical_from_taps(scene, b, tiles, fills, by_id[cb_id], ct, cf)
        else:
            c = scenes.canonical_from_taps(scene, b, tiles, fills)
        out["batches"].append(c)
        out["group_hashes"].append(c["group_hashes"])
    for b in scene["clip_batches"]:
        bid = int(b["info"][0])
        if bid in taps:
            tiles, fills = taps[bid]
            out["group_hashes"].append(scenes.canonical_from_taps(scene, b, tiles, fills)["group_hashes"])
    out["group_hashes"] = np.sort(np.concatenate(out["group_hashes"])) if out["group_hashes"] else np.zeros(0, "<u8")
    return out


def reference_from_extra(extra):
    ref = {"group_hashes": extra["ref_group_hashes"], "batches": []}
    for i in range(int(extra["ref_n_batches"])):
        ref["batches"].append({k: extra["ref%d_%s" % (i, k)] for k in ("tiles", "fills", "clips", "z")})
    return ref


def assert_canonical_equal(mine, ref, what=""):
    assert len(mine["batches"]) == len(ref["batches"]), what
    assert np.array_equal(mine["group_hashes"], ref["group_hashes"]), what + ": per-tile fill multisets differ"
    for i, (m, r) in enumerate(zip(mine["batches"], ref["batches"])):
        for k in ("tiles", "fills", "clips"):
            assert len(m[k]) == len(r[k]), "%s batch %d: %s count %d != %d" % (what, i, k, len(m[k]), len(r[k]))
            assert np.array_equal(m[k], r[k]), "%s batch %d: %s differ" % (what, i, k)


def canonical_digest(c, z_list):
    parts = [c["group_hashes"]]
    for b, z in zip(c["batches"], z_list):
        parts += [b["tiles"], b["fills"], b["clips"], np.asarray(z, "<u4").reshape(-1)]
    return scenes.digest(*parts)

#!/usr/bin/env python
"""Generates the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Needs /root/reference (this container only): oracle/Makefile compiles the unmodified reference CPU sources
into oracle/_ref/libpfref.so; this script drives it through oracle/pfref.py and stores

  <scene>.npz   inputs  = what SceneBuilderD3D11::build hands to the renderer (core/d3d11/scene_builder.cpp:209-223)
                outputs = SceneBuilderD3D9::build's fills / tiles / clips / z-buffers (core/d3d9/scene_builder.cpp:97-109)
                          in the order- and id-independent canonical form of tests/scenes.py
  area_lut.npz  the 256 x 256 RGBA8 area LUT as decoded by the reference (core/renderer.cpp:17-21)
  digests.json  sha256 of the canonical outputs of the configurations too large to commit (tiger 4096)
  synthetic_200k_8192.npz  BASELINE.json configs[3] at its STATED size (200,000 cubic blobs at 8192 x 8192): sha256 of every
                geometry tap of oracle/pf_oracle.c (lines, fills, tiles, z, sorted lists -- the restatement is pinned to the
                reference's tiler on the fixtures above) + the channel sums of every 16 x 16 tile of the oracle's frame
                (`python tests/golden/make_golden.py config4` regenerates only this one: 2 minutes of CPU)
  shader_frames.npz  RGBA8 frames of the 512^2 fixtures as the reference's OWN compute shaders render them: fill.comp and
                tile.comp, read where they lie, compiled by g++ through oracle/ref_harness/glsl_shim.h and run on the CPU
                (oracle/_ref/libpfshader.so, oracle/pfshader.py)

Usage: python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pfref  # noqa: E402
import scenes  # noqa: E402

# name -> (asset, size, native size of the SVG view box)
SVG_SCENES = {
    "tiger_512": ("tiger.svg", 512, 900.0),
    "tiger_1024": ("tiger.svg", 1024, 900.0),      # BASELINE.json configs[0]
    "features_2048": ("features.svg", 2048, 720.0),  # BASELINE.json configs[1], SVG half
}
DIGEST_ONLY = {
    "tiger_4096": ("tiger.svg", 4096, 900.0),      # BASELINE.json configs[2]
}
# The incremental-frame fixture: tiger 512 with ONE draw path moved (Outline::transform, core/data/path.h:20-21), built by
# the reference front end like the others; inputs only (its pixels are compared with a full CUDA render of the same inputs,
# which the other tests pin). name -> (base scene, draw path index, dx, dy)
MOVED_SCENES = {"tiger_512_moved": ("tiger_512", 145, 23.25, -17.5)}

# demo/common/app.cpp:21-101 blocks: 1 rect, 2 clip circle, 4 blurred shadow, 8 image, 16 gradient stroke, 32 render
# target pattern. 0x3f = the whole primitives scene of the demo (shadow = two blur passes through render targets).
# name -> (size, scale, feature mask)
DEMO_SCENES = {
    "demo_clip_512": (512, 1.0, 1 | 2 | 16),
    "demo_full_512": (512, 1.0, 0x3f),
    "demo_full_2048": (2048, 2048 / 720.0, 0x3f),  # BASELINE.json configs[1], demo-primitives half
}


# every BlendMode over a gradient, radial gradients, a repeating unsmoothed image pattern (ref_harness.cpp
# pfref_scene_paints): the tile.comp branches no SVG / demo scene reaches. name -> (size, scale)
PAINT_SCENES = {"paints_512": (512, 1.0)}

# scenes whose frame, rendered by the reference's own fill.comp + tile.comp on the CPU, is committed (shader_frames.npz)
SHADER_FRAMES = ["tiger_512", "demo_clip_512", "demo_full_512", "paints_512"]


def canonical_extra(ref):
    extra = {"ref_group_hashes": ref["group_hashes"], "ref_n_batches": np.int32(len(ref["batches"]))}
    for i, b in enumerate(ref["batches"]):
        for k in ("tiles", "fills", "clips", "z"):
            extra["ref%d_%s" % (i, k)] = b[k]
    return extra


def canonical_digest(ref):
    parts = [ref["group_hashes"]]
    for b in ref["batches"]:
        parts += [b["tiles"], b["fills"], b["clips"], b["z"]]
    return scenes.digest(*parts)


def geometry_digest(ref):
    """The same without the z-buffers (the GPU-driven path keeps z in dense-tile terms: parity.py maps tiles, not z)."""
    parts = [ref["group_hashes"]]
    for b in ref["batches"]:
        parts += [b["tiles"], b["fills"], b["clips"]]
    return scenes.digest(*parts)


def tap_digests(fr, slot):
    """sha256 of every geometry tap of one oracle batch, in the canonical forms tests/test_gpu_parity.py compares."""
    lines = fr.clipped_lines(slot)
    w = np.frombuffer(lines.tobytes(), "<u4").reshape(len(lines), -1)
    lines_sorted = w[np.lexsort(w.T[::-1])]
    tiles = fr.tiles(slot)
    off, lst = fr.tile_lists(slot)
    out = {"lines": scenes.digest(lines_sorted), "fills": scenes.digest(fr.fills(slot)),
           "z": scenes.digest(fr.z(slot)[0]), "tile_lists": scenes.digest(off, lst)}
    for f in ("fill_count", "backdrop", "backdrop_delta", "backdrop_d3d9", "listed"):
        out["tiles_" + f] = scenes.digest(tiles[f])
    out["tiles_has_alpha"] = scenes.digest(tiles["alpha_tile_id"] >= 0)
    return out


def config4():
    """BASELINE.json configs[3] at full size through the C restatement (no reference front end involved: the scene is
    built by tests/scenes.py from the PCG32 stream SURVEY.md names)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    import pforacle
    n_paths, size = scenes.CONFIG4
    scene = scenes.synthetic_scene(n_paths, size)
    lut = np.load(os.path.join(HERE, "area_lut.npz"))["lut"]
    fr = pforacle.Frame(scene, lut)
    px = fr.render()
    slot = fr.slots[int(scene["draw_batches"][0]["info"][0])]
    dig = tap_digests(fr, slot)
    counts = fr.counts(slot)
    fr.close()
    np.savez_compressed(os.path.join(HERE, "synthetic_%dk_%d.npz" % (n_paths // 1000, size)), tile_sums=scenes.tile_sums(px),
                        digests=np.array(json.dumps(dig)), counts=np.array(json.dumps(counts)),
                        frame_sha256=np.array(scenes.digest(px)))
    print("config4", counts, dig)


def moved_scene(name, base, index, dx, dy):
    asset, size, native = SVG_SCENES[base]
    s = pfref.RefScene.from_svg(pfref.asset(asset), size, size, size / native)
    s.translate_draw_path(index, dx, dy)
    scenes.save_scene(scenes.golden_path(name), s.build_d3d11())
    s.close()
    print(name, "path", index, "moved by", (dx, dy))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "config4":
        config4()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "moved":
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
        for name, (base, index, dx, dy) in MOVED_SCENES.items():
            moved_scene(name, base, index, dx, dy)
        return
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    digests = {}
    lut = None
    for name, (asset, size, native) in {**SVG_SCENES, **DIGEST_ONLY}.items():
        s = pfref.RefScene.from_svg(pfref.asset(asset), size, size, size / native)
        scene = s.build_d3d11()
        ref = scenes.canonical_from_reference(s.build_d3d9())
        if lut is None:
            lut = s.area_lut()
        digests[name] = {"canonical_sha256": canonical_digest(ref), "geometry_sha256": geometry_digest(ref),
                         "counts": s.counts(),
                         "fills": int(sum(len(b["fills"]) for b in ref["batches"])),
                         "tiles": int(sum(len(b["tiles"]) for b in ref["batches"])),
                         "alpha_tiles": int(len(ref["group_hashes"]))}
        if name in SVG_SCENES:
            scenes.save_scene(scenes.golden_path(name), scene, canonical_extra(ref))
        else:
            # inputs only (what bench.py renders); the outputs are pinned by the digest
            scenes.save_scene(scenes.golden_path(name + "_scene"), scene)
        s.close()
        print(name, digests[name])
    for name, (base, index, dx, dy) in MOVED_SCENES.items():
        moved_scene(name, base, index, dx, dy)
    for name, (size, scale, features) in DEMO_SCENES.items():
        s = pfref.RefScene.demo(size, size, scale, pfref.asset("sea.png"), features)
        scene = s.build_d3d11()
        ref = scenes.canonical_from_reference(s.build_d3d9())
        digests[name] = {"canonical_sha256": canonical_digest(ref), "counts": s.counts()}
        scenes.save_scene(scenes.golden_path(name), scene, canonical_extra(ref))
        s.close()
        print(name, digests[name])
    for name, (size, scale) in PAINT_SCENES.items():
        s = pfref.RefScene.paints(size, size, scale, pfref.asset("sea.png"))
        scene = s.build_d3d11()
        ref = scenes.canonical_from_reference(s.build_d3d9())
        digests[name] = {"canonical_sha256": canonical_digest(ref), "counts": s.counts()}
        scenes.save_scene(scenes.golden_path(name), scene, canonical_extra(ref))
        s.close()
        print(name, digests[name])
    np.savez_compressed(os.path.join(HERE, "area_lut.npz"), lut=lut)
    # The pixel half: the reference's OWN compute shaders (fill.comp, tile.comp) run on the CPU through
    # oracle/_ref/libpfshader.so on the (pinned) geometry of every 512^2 fixture: what the GPU-driven mode renders
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "shaders", "port"])
    import pforacle
    import pfshader
    frames = {}
    for name in SHADER_FRAMES:
        scene, _ = scenes.load_scene(scenes.golden_path(name))
        fr = pforacle.Frame(scene, lut)
        fr.render()
        dest, pages, _ = pfshader.render_frame(scene, fr, lut)
        fr.close()
        frames[name] = dest
        digests[name]["shader_frame_sha256"] = scenes.digest(dest)
        print(name, "shader frame", digests[name]["shader_frame_sha256"])
    np.savez_compressed(os.path.join(HERE, "shader_frames.npz"), **frames)
    with open(os.path.join(HERE, "digests.json"), "w") as fp:
        json.dump(digests, fp, indent=1, sort_keys=True)
    config4()


if __name__ == "__main__":
    main()

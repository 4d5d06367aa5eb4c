"""GPU tests (pytest -m gpu): the CUDA path, called through the C-ABI, against the reference fixtures and the oracle.

* dice / bin / propagate outputs: BIT-EXACT against the reference's hybrid CPU tiler (tests/golden fixtures) and
  against the oracle's taps (lines, fills, tiles, z, sorted lists);
* pixels: within 1/255 per channel of the oracle's restatement of fill.comp + tile.comp.
"""
import numpy as np
import pytest

import parity
import scenes

pytestmark = pytest.mark.gpu

GOLDEN = ["tiger_512", "tiger_1024", "features_2048", "demo_clip_512", "demo_full_512", "demo_full_2048", "paints_512"]
PIXEL_TOL = 1  # 1/255 per channel, BASELINE.json north_star


@pytest.fixture(scope="module")
def renderer(area_lut):
    import pfcu

    r = pfcu.Renderer(0, area_lut)
    yield r
    r.close()


def cuda_taps(r, scene):
    taps = {}
    for kind in ("clip", "draw"):
        for b in scene[kind + "_batches"]:
            if int(b["info"][1]) > 0 or kind == "draw":
                bid = int(b["info"][0])
                taps[bid] = (r.tiles(bid), r.fills(bid))
    return taps


def oracle_frame(scene, lut):
    import pforacle

    fr = pforacle.Frame(scene, lut)
    px = fr.render()
    return fr, px


def raw_sorted(records):
    """Rows of a structured array as raw uint32 words, lexicographically sorted (bit-exact multiset compare)."""
    w = np.frombuffer(records.tobytes(), "<u4").reshape(len(records), -1)
    return w[np.lexsort(w.T[::-1])]


@pytest.mark.parametrize("name", GOLDEN)
def test_geometry_bit_exact_vs_reference_fixture(renderer, name):
    scene, extra = scenes.load_scene(scenes.golden_path(name))
    renderer.set_scene(scene)
    stats = renderer.draw(clear=True)
    mine = parity.canonical_scene(scene, cuda_taps(renderer, scene))
    parity.assert_canonical_equal(mine, parity.reference_from_extra(extra), name)
    assert stats["kernel_launches"] > 0


@pytest.mark.parametrize("name", GOLDEN)
def test_taps_bit_exact_vs_oracle(renderer, area_lut, name):
    scene, _ = scenes.load_scene(scenes.golden_path(name))
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    fr, _ = oracle_frame(scene, area_lut)
    for kind in ("clip", "draw"):
        for b in scene[kind + "_batches"]:
            if int(b["info"][1]) == 0:
                continue
            bid = int(b["info"][0])
            slot = fr.slots[bid]
            # dice: same multiset of flattened + view-box-clipped lines, bit for bit (raw 20-byte records)
            lo, lc = fr.clipped_lines(slot), renderer.lines(bid)
            assert len(lo) == len(lc), "line count %d != %d" % (len(lc), len(lo))
            ro, rc = raw_sorted(lo), raw_sorted(lc)
            assert np.array_equal(ro, rc), "flattened lines differ"
            # bin: identical CSR fills (both canonical-sorted)
            assert np.array_equal(fr.fills(slot), renderer.fills(bid)), "fills differ"
            # propagate: backdrops, deltas, list membership, hybrid-equivalent backdrop
            to, tc = fr.tiles(slot), renderer.tiles(bid)
            for f in ("fill_count", "backdrop", "backdrop_delta", "backdrop_d3d9", "listed"):
                assert np.array_equal(to[f], tc[f]), f
            assert np.array_equal(to["alpha_tile_id"] >= 0, tc["alpha_tile_id"] >= 0)
            assert np.array_equal(to["clip_alpha_tile_id"] >= 0, tc["clip_alpha_tile_id"] >= 0)
            # z-buffer and sorted, culled lists
            assert np.array_equal(fr.z(slot)[0], renderer.z(bid)), "z differs"
            oo, ot = fr.tile_lists(slot)
            co, ct = renderer.tile_lists(bid)
            assert np.array_equal(oo, co) and np.array_equal(ot, ct), "tile lists differ"
    fr.close()


@pytest.mark.parametrize("name", GOLDEN)
def test_pixels_within_tolerance_of_oracle(renderer, area_lut, name):
    """The rasterizing end (fill + tile kernels) against the restatement of fill.comp / tile.comp."""
    scene, _ = scenes.load_scene(scenes.golden_path(name))
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    got = renderer.pixels()
    fr, want = oracle_frame(scene, area_lut)
    diff = np.abs(got.astype(int) - want.astype(int))
    print("%s: pixels off by one: %d of %d" % (name, int((diff.max(axis=2) > 0).sum()), diff.shape[0] * diff.shape[1]))
    assert diff.max() <= PIXEL_TOL, "max diff %d at %s" % (diff.max(), np.unravel_index(diff.argmax(), diff.shape))
    fr.close()


def test_masks_match_oracle(renderer, area_lut):
    """fill stage alone: every sampled 16 x 16 coverage mask within 1/255 of fill.comp's restatement."""
    scene, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    fr, _ = oracle_frame(scene, area_lut)
    bid = int(scene["draw_batches"][0]["info"][0])
    to = fr.tiles(fr.slots[bid])
    own = np.nonzero((to["alpha_tile_id"] >= 0) & (to["fill_count"] > 0))[0]
    try:
        # every mask, as fill.comp rasterizes them; then the default: masks of z-culled tiles are skipped, so only tiles
        # that made it into the sorted lists are compared (and some tile must actually have been culled)
        for fill_culled in (True, False):
            renderer.set_fill_culled_tiles(fill_culled)
            renderer.set_scene(scene)
            renderer.draw(clear=True)
            tc = renderer.tiles(bid)
            sel = own
            if not fill_culled:
                kept = np.zeros(len(tc), bool)
                kept[renderer.tile_lists(bid)[1]] = True
                assert (~kept[own]).any()
                sel = own[kept[own]]
            worst = 0
            for i in sel[:: max(1, len(sel) // 400)]:
                a = fr.mask(int(to["alpha_tile_id"][i])).astype(int)
                b = renderer.mask(int(tc["alpha_tile_id"][i])).astype(int)
                worst = max(worst, np.abs(a - b).max())
            assert worst <= 1
    finally:
        renderer.set_fill_culled_tiles(False)
    fr.close()


@pytest.mark.parametrize("name", GOLDEN + ["tiger_4096_scene"])
def test_fused_fill_is_byte_identical_to_separate_fill(renderer, name):
    """PFCU_OPT_FUSED_FILL: the tile kernel rasterizing its own masks in shared memory and the separate fill kernel (the
    default) writing them to device memory run the same mask arithmetic -- the frames must be equal byte for byte, on every
    fixture (clip masks, even-odd, render-target passes, LOAD frames included)."""
    scene, _ = scenes.load_scene(scenes.golden_path(name))
    frames = []
    try:
        for fused in (True, False):
            renderer.set_fused_fill(fused)
            renderer.set_scene(scene)
            renderer.draw(clear=True)
            st = renderer.draw(clear=True)
            frames.append(renderer.pixels().copy())
            # the separate mode launches one fill kernel more per draw batch
            frames.append(st["kernel_launches"])
    finally:
        renderer.set_fused_fill(False)
    assert frames[1] < frames[3], (frames[1], frames[3])
    assert np.array_equal(frames[0], frames[2]), "%d bytes differ" % int((frames[0] != frames[2]).sum())


@pytest.mark.parametrize("name", ["tiger_512", "features_2048", "demo_full_512", "paints_512", "tiger_4096_scene"])
def test_ordered_tile_groups_change_nothing_but_the_order(renderer, name):
    """PFCU_OPT_ORDER_TILE_GROUPS (default on): the tile kernel's CTAs take the groups of 16 tiles by descending cost and find
    their list headers in that order. Pixels, z-buffer and the z-culled lists must be what the grid-order frame gives."""
    scene, _ = scenes.load_scene(scenes.golden_path(name))
    got = []
    try:
        for ordered in (True, False):
            renderer.set_order_tile_groups(ordered)
            renderer.set_scene(scene)
            renderer.draw(clear=True)
            st = renderer.draw(clear=True)
            taps = []
            for b in scene["draw_batches"]:
                bid = int(b["info"][0])
                taps.append((renderer.z(bid), renderer.tile_lists(bid)))
            got.append((renderer.pixels().copy(), taps, st["kernel_launches"]))
    finally:
        renderer.set_order_tile_groups(True)
    assert np.array_equal(got[0][0], got[1][0]), "%d bytes differ" % int((got[0][0] != got[1][0]).sum())
    for (za, (oa, ta)), (zb, (ob, tb)) in zip(got[0][1], got[1][1]):
        assert np.array_equal(za, zb) and np.array_equal(oa, ob) and np.array_equal(ta, tb)


def test_device_timeline_of_the_trace_build(area_lut):
    """pfcu_set_timeline (lib/libpfcu_trace.so, the build with -DPFCU_TIMELINE): one record per CTA of every kernel of the
    frame, stamps in order, and the same pixels as the product library; the product library refuses the call."""
    import os
    import subprocess
    import sys

    import pfcu

    r = pfcu.Renderer(0, area_lut)
    with pytest.raises(pfcu.PfcuError):
        r.set_timeline(1000)
    scene, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    r.set_scene(scene)
    r.draw(clear=True)
    want = r.pixels().copy()
    r.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import sys, numpy as np
sys.path[:0] = [%r, %r]
import pfcu, scenes
lut = np.load(%r)["lut"]
r = pfcu.Renderer(0, lut)
r.set_retain_frame_graph(False)
r.set_timeline(100000)
scene, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
r.set_scene(scene)
st = r.draw(clear=True)
rec = r.read_timeline()
assert len(rec) > 1000 and st["retries"] == 0, (len(rec), st)
assert (rec["t_placed_ns"] <= rec["t_start_ns"]).all() and (rec["t_start_ns"] <= rec["t_end_ns"]).all()
stages = set(int(x) & 0xff for x in rec["stage"])
assert stages == set(range(10)), stages
for s in stages:  # every CTA of every kernel reported exactly once
    q = rec[(rec["stage"] & 0xff) == s]
    for sub in set(int(x) for x in q["stage"]):
        k = q[q["stage"] == sub]
        assert len(k) == int(k["n_ctas"][0]) and len(set(int(c) for c in k["cta"])) == len(k), (sub, len(k))
# the tile kernel starts after everything it depends on has ended
comp = rec[rec["stage"] == 9]
assert comp["t_start_ns"].min() >= rec[rec["stage"] == 8]["t_end_ns"].max()
assert comp["t_start_ns"].min() >= rec[rec["stage"] == 7]["t_end_ns"].max()
assert len(r.read_timeline()) == 0
np.save(sys.argv[1], r.pixels())
""" % (os.path.join(root, "pathfinder-cpp_b200"), os.path.join(root, "tests"), os.path.join(root, "tests", "golden", "area_lut.npz"))
    out = os.path.join(root, "gpurun_out", "timeline_test_pixels.npy")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    env = dict(os.environ, PFCU_LIB=os.path.join(root, "pathfinder-cpp_b200", "lib", "libpfcu_trace.so"))
    subprocess.run([sys.executable, "-c", code, out], check=True, env=env, timeout=300)
    assert np.array_equal(np.load(out), want)


@pytest.mark.parametrize("name", ["demo_full_512", "demo_full_2048", "demo_clip_512", "paints_512", "features_2048", "tiger_512"])
def test_concurrent_batches_render_the_same_frame(area_lut, name):
    """PFCU_OPT_CONCURRENT_BATCHES: the batches of a frame prepared side by side on four stream pairs (clip batches first,
    tile passes in order). Same pixels and same render-target pages as the batch-after-batch frame, eagerly and from the
    retained frame graph (frames 3 .. 5 are graph launches)."""
    import pfcu

    scene, _ = scenes.load_scene(scenes.golden_path(name))
    frames = []
    for concurrent in (False, True):
        r = pfcu.Renderer(0, area_lut)
        r.set_concurrent_batches(concurrent)
        r.set_scene(scene)
        got = []
        for _ in range(5):
            st = r.draw(clear=True)
            assert st["overflow_flags"] == 0
            got.append(r.pixels().copy())
        r.close()
        for g in got[1:]:
            assert np.array_equal(g, got[0]), "frames of one renderer differ (concurrent=%s)" % concurrent
        frames.append(got[0])
    assert np.array_equal(frames[0], frames[1]), "%d bytes differ" % int((frames[0] != frames[1]).sum())


def compare_with_oracle(renderer, area_lut, scene, what):
    """Geometry taps bit-exact + pixels within tolerance for a scene that has no reference fixture."""
    renderer.set_scene(scene)
    st = renderer.draw(clear=True)
    fr, want = oracle_frame(scene, area_lut)
    for b in scene["draw_batches"]:
        bid = int(b["info"][0])
        slot = fr.slots[bid]
        assert np.array_equal(raw_sorted(fr.clipped_lines(slot)), raw_sorted(renderer.lines(bid))), what + ": lines"
        assert np.array_equal(fr.fills(slot), renderer.fills(bid)), what + ": fills"
        to, tc = fr.tiles(slot), renderer.tiles(bid)
        for f in ("fill_count", "backdrop", "backdrop_delta", "backdrop_d3d9", "listed"):
            assert np.array_equal(to[f], tc[f]), what + ": " + f
        oo, ot = fr.tile_lists(slot)
        co, ct = renderer.tile_lists(bid)
        assert np.array_equal(oo, co) and np.array_equal(ot, ct), what + ": tile lists"
    diff = np.abs(renderer.pixels().astype(int) - want.astype(int))
    assert diff.max() <= PIXEL_TOL, what + ": pixels off by %d" % diff.max()
    fr.close()
    return st


def test_long_lines_bit_exact_vs_oracle(renderer, area_lut):
    """Lines that cross hundreds of tiles take the warp-per-line path of bin (merge of the two crossing sequences):
    axis-aligned, diagonal, nearly-horizontal, nearly-vertical, reversed, starting above / left of the view box."""
    def poly(pts, paint, rule=0):
        pts = np.array(pts, "<f4")
        return {"contours": [(pts, np.zeros(len(pts), "u1"))], "paint": paint, "fill_rule": rule, "opaque": False}

    W = 3000
    paths = [
        poly([[0, 0], [W, 0], [W, W], [0, W]], 0),                                  # full-canvas rectangle
        poly([[3.25, 7.5], [2990.75, 2801.125], [2989.5, 2810.0], [1.0, 19.0]], 1),  # thin diagonal sliver
        poly([[10, 1500.3], [2995, 1503.9], [2995, 1504.4], [10, 1500.9]], 2),       # nearly horizontal
        poly([[1501.3, 5], [1504.2, 2996], [1504.9, 2996], [1501.8, 5]], 3),         # nearly vertical
        poly([[2999, 2999], [0.5, 0.5], [40.0, 0.5]], 1, rule=1),                   # reversed diagonal, even-odd
        poly([[-500, -800], [2500, 2900], [2400, 2950]], 2),                         # starts above and left of the view
        poly([[16, 16], [2992, 16], [2992, 2992], [16, 2992]], 3),                   # edges exactly on tile boundaries
        poly([[100, 3500], [2900, -300], [2950, -250]], 0),                          # crosses top and bottom
    ]
    colors = np.array([[255, 255, 255, 255], [255, 0, 0, 128], [0, 255, 0, 200], [0, 0, 255, 77]], "u1")
    scene = scenes.build_scene_from_outlines(W, W, paths, colors)
    st = compare_with_oracle(renderer, area_lut, scene, "long lines")
    assert st["overflow_flags"] == 0


def test_synthetic_blobs_bit_exact_vs_oracle(renderer, area_lut):
    """BASELINE.json config 4 in miniature: 1500 random cubic blobs at 1024 x 1024 (many tiny paths, deep lists)."""
    scene = scenes.synthetic_scene(1500, 1024)
    compare_with_oracle(renderer, area_lut, scene, "synthetic blobs")


@pytest.mark.parametrize("size", [(1000, 700), (333, 517), (40, 1999)])
def test_ragged_canvas_sizes_bit_exact_vs_oracle(renderer, area_lut, size):
    """Canvases whose sides are not multiples of the 16-pixel tile and whose tile count is not a multiple of the 16-tile
    groups of the tile kernel (the last group is partial; edge tiles store pixel by pixel): geometry taps bit-exact, pixels
    within tolerance, with the tile groups in order of cost (the default) and in grid order."""
    w, h = size
    paths, colors = scenes.synthetic_paths(600, max(w, h))
    scene = scenes.build_scene_from_outlines(w, h, paths, colors)
    try:
        for ordered in (True, False):
            renderer.set_order_tile_groups(ordered)
            compare_with_oracle(renderer, area_lut, scene, "ragged %dx%d ordered=%s" % (w, h, ordered))
            assert renderer.pixels().shape == (h, w, 4)
    finally:
        renderer.set_order_tile_groups(True)


def test_repeat_frames_are_identical_and_reuse_buffers(renderer):
    scene, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    a = renderer.pixels()
    st = renderer.draw(clear=True, upload=True)
    b = renderer.pixels()
    assert st["retries"] == 0
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1


def test_empty_and_degenerate_scenes(renderer):
    """Empty scene, zero-area path, non-finite control points, a path entirely outside the view box."""
    paths = [
        {"contours": [(np.array([[10, 10], [10, 10], [10, 10]], "<f4"), np.array([0, 0, 0], "u1"))], "paint": 0},
        {"contours": [(np.array([[5, 5], [np.nan, 7], [9, np.inf], [20, 20]], "<f4"), np.array([0, 1, 2, 0], "u1"))],
         "paint": 0},
        {"contours": [(np.array([[-500, -500], [-400, -500], [-400, -400]], "<f4"), np.array([0, 0, 0], "u1"))],
         "paint": 0},
        {"contours": [(np.array([[8, 8], [40, 8], [40, 40], [8, 40]], "<f4"), np.array([0, 0, 0, 0], "u1"))], "paint": 1},
    ]
    colors = np.array([[255, 0, 0, 255], [0, 0, 255, 128]], "u1")
    scene = scenes.build_scene_from_outlines(64, 48, paths, colors)
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    px = renderer.pixels()
    assert px.shape == (48, 64, 4)
    assert px[20, 20, 3] == 128 and px[2, 2, 3] == 0
    empty = scenes.build_scene_from_outlines(32, 32, [], np.array([[0, 0, 0, 255]], "u1"))
    renderer.set_scene(empty)
    st = renderer.draw(clear=True)
    assert st["fills"] == 0 and renderer.pixels().max() == 0


def test_tiger_4096_geometry_digest(renderer):
    """BASELINE.json configs[2] at full size: fills / tiles / backdrops of tiger.svg @ 4096 x 4096 hash to the digest of
    the REFERENCE's own output (tests/golden/digests.json, written by make_golden.py from SceneBuilderD3D9::build)."""
    import json
    import os

    with open(os.path.join(scenes.GOLDEN, "digests.json")) as fp:
        want = json.load(fp)["tiger_4096"]
    scene, _ = scenes.load_scene(scenes.golden_path("tiger_4096_scene"))
    renderer.set_scene(scene)
    st = renderer.draw(clear=True)
    mine = parity.canonical_scene(scene, cuda_taps(renderer, scene))
    parts = [mine["group_hashes"]]
    for b in mine["batches"]:
        parts += [b["tiles"], b["fills"], b["clips"]]
    assert scenes.digest(*parts) == want["geometry_sha256"]
    assert st["fills"] == want["fills"] and len(mine["group_hashes"]) == want["alpha_tiles"]


def stacked_scene(n_layers, size=96):
    """n_layers translucent rectangles (a few even-odd, a few with a hole) over the same tiles: lists far deeper than the
    tile kernel's shared-memory staging (256 entries per 16 tiles), nothing for the z-buffer to cull."""
    rng = np.random.RandomState(7)
    paths = []
    for i in range(n_layers):
        x0, y0 = rng.uniform(2, 30, 2)
        x1, y1 = rng.uniform(50, size - 2, 2)
        pts = np.array([[x0, y0], [x1, y0 + rng.uniform(-1, 1)], [x1, y1], [x0 + rng.uniform(-1, 1), y1]], "<f4")
        contours = [(pts, np.zeros(4, "u1"))]
        if i % 7 == 3:  # a hole (opposite winding) inside
            hole = np.array([[36, 36], [36, 44], [44, 44], [44, 36]], "<f4")
            contours.append((hole, np.zeros(4, "u1")))
        paths.append({"contours": contours, "paint": i % 5, "fill_rule": 1 if i % 11 == 5 else 0, "opaque": False})
    colors = np.array([[255, 40, 40, 9], [40, 255, 40, 14], [40, 40, 255, 5], [250, 250, 30, 11], [20, 20, 20, 7]], "u1")
    return scenes.build_scene_from_outlines(size, size, paths, colors)


@pytest.mark.parametrize("n_layers", [40, 300, 700])
def test_deep_lists_match_oracle(renderer, area_lut, n_layers):
    """Framebuffer tiles with 40 / 300 / 700 surviving layers: lists in one staging round, in several rounds, and a
    single tile whose list exceeds the staging (walked by selection from global memory)."""
    compare_with_oracle(renderer, area_lut, stacked_scene(n_layers), "%d stacked layers" % n_layers)


def test_clear_colour_and_load_action(renderer, area_lut):
    """LOAD_ACTION_CLEAR with a non-zero clear colour, then the same batch again with LOAD_ACTION_LOAD over the result
    (tile.comp:743-749): empty tiles keep the destination, the others blend over it."""
    import pforacle

    scene, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    cc = (0.25, 0.5, 0.75, 1.0)
    renderer.set_scene(scene)
    renderer.draw(clear=True, clear_color=cc)
    first = renderer.pixels()
    fr = pforacle.Frame(scene, area_lut)
    want = fr.render(clear=True, clear_color=cc)
    assert np.abs(first.astype(int) - want.astype(int)).max() <= PIXEL_TOL
    assert tuple(first[0, 0]) == (64, 128, 191, 255)
    # second pass without clearing: the oracle draws over its own first result as well
    renderer.draw(clear=False)
    second = renderer.pixels()
    want2 = fr.render(clear=False)
    # the destination is re-read as RGBA8: one more quantisation step on both sides
    assert np.abs(second.astype(int) - want2.astype(int)).max() <= PIXEL_TOL + 1
    assert tuple(second[0, 0]) == (64, 128, 191, 255)
    fr.close()


def test_strips_of_the_full_size_canvas_are_bit_identical(renderer):
    """Size-independent property at a BASELINE size: a horizontal strip of the synthetic canvas rendered
    with the framebuffer-origin mechanism equals the same rows of the full frame, byte for byte."""
    full_scene = scenes.synthetic_scene(20000, 4096)
    renderer.set_scene(full_scene)
    renderer.draw(clear=True)
    full = renderer.pixels()
    for y0, y1 in ((0, 1024), (1536, 2560), (3072, 4096)):
        strip_scene = scenes.synthetic_scene(20000, 4096, strip=(y0, y1))
        renderer.set_scene(strip_scene)
        renderer.draw(clear=True)
        assert np.array_equal(renderer.pixels(), full[y0:y1]), "strip %d..%d differs" % (y0, y1)


def test_retained_frame_graph_serves_identical_frames_only(renderer, area_lut):
    """PFCU_OPT_RETAIN_FRAME_GRAPH: the third identical frame onwards is one graph launch issued by pfcu_end_frame. It
    must render the same bytes as the kernel-by-kernel path, pick up re-uploaded inputs, and step aside (and come back)
    when the frame changes."""
    tiger, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    clip, _ = scenes.load_scene(scenes.golden_path("demo_clip_512"))
    renderer.set_retain_frame_graph(False)
    renderer.set_scene(tiger)
    renderer.draw(clear=True)
    want_tiger = renderer.pixels()
    renderer.set_scene(clip)
    renderer.draw(clear=True)
    want_clip = renderer.pixels()
    renderer.set_retain_frame_graph(True)
    try:
        renderer.set_scene(tiger)
        for i in range(5):  # frames 0, 1 enqueue kernels (1 also captures), 2.. are graph launches
            st = renderer.draw(clear=True, upload=(i == 3))
            assert st["retries"] == 0 and st["fills"] == 27200
            assert np.array_equal(renderer.pixels(), want_tiger), "frame %d" % i
        bid = int(tiger["draw_batches"][0]["info"][0])
        assert len(renderer.fills(bid)) == 27200  # the taps see the graph-launched frame's buffers
        renderer.set_scene(clip)  # a different frame: the retained graph must not be used
        for i in range(4):
            renderer.draw(clear=True)
            assert np.array_equal(renderer.pixels(), want_clip), "clip frame %d" % i
        renderer.set_scene(tiger)
        renderer.draw(clear=True, clear_color=(0.0, 0.0, 0.0, 1.0))  # same batches, other clear colour: other frame
        assert renderer.pixels()[0, 0, 3] == 255
        renderer.draw(clear=True)
        assert np.array_equal(renderer.pixels(), want_tiger)
    finally:
        renderer.set_retain_frame_graph(True)


def test_many_small_paths_bit_exact_vs_oracle(renderer, area_lut):
    """More than 65 536 segments in one batch: dice takes its 4-CTAs-per-SM configuration and propagate its
    lane-per-column walk for glyph-sized paths (BASELINE.json config 4 at a size the oracle finishes in seconds)."""
    scene = scenes.synthetic_scene(8000, 2048)
    assert sum(int(b["info"][3]) for b in scene["draw_batches"]) > 65536
    compare_with_oracle(renderer, area_lut, scene, "8000 blobs")


def test_frames_in_flight_on_two_contexts(renderer, area_lut):
    """pfcu_submit_frame / pfcu_wait_frame: two contexts each keep a frame in flight (uploads included); every frame is
    byte-identical to the blocking pfcu_end_frame render, and the call-sequence rules are enforced."""
    import pfcu

    tiger, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    demo, _ = scenes.load_scene(scenes.golden_path("demo_full_512"))
    want = {}
    for name, scene in (("tiger", tiger), ("demo", demo)):
        renderer.set_scene(scene)
        renderer.draw(clear=True)
        want[name] = renderer.pixels()
    a, b = pfcu.Renderer(0, area_lut), pfcu.Renderer(0, area_lut)
    try:
        a.set_scene(tiger)
        b.set_scene(demo)
        for i in range(6):
            assert a.draw(clear=True, upload=True, wait=False) is None
            assert b.draw(clear=True, upload=True, wait=False) is None
            if i == 2:  # nothing but pfcu_wait_frame is accepted while the frame is pending
                with pytest.raises(RuntimeError):
                    a.draw(clear=True, wait=False)
                with pytest.raises(RuntimeError):
                    a.upload_segments(tiger)
            sa, sb = a.wait(), b.wait()
            assert sa["fills"] == 27200 and sa["retries"] == 0 or i == 0
            assert np.array_equal(a.pixels(), want["tiger"]), "tiger frame %d" % i
            assert np.array_equal(b.pixels(), want["demo"]), "demo frame %d" % i
        with pytest.raises(RuntimeError):
            a.wait()  # no frame submitted
    finally:
        a.close()
        b.close()


def test_region_readback_matches_the_full_frame(renderer):
    """pfcu_read_target_region = CommandEncoder::read_texture with a region (gpu/command_encoder.cpp:317-355)."""
    scene, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    full = renderer.pixels()
    for x, y, w, h in ((0, 0, 512, 512), (17, 33, 100, 7), (500, 500, 12, 12), (3, 0, 1, 512)):
        assert np.array_equal(renderer.pixels_region(x, y, w, h), full[y:y + h, x:x + w])
    for bad in ((-1, 0, 4, 4), (0, 0, 0, 4), (510, 0, 4, 4), (0, 510, 4, 4)):
        with pytest.raises(RuntimeError):
            renderer.pixels_region(*bad)


@pytest.mark.parametrize("name", ["tiger_512", "demo_clip_512", "demo_full_512", "paints_512"])
def test_pixels_against_reference_shader_frames(renderer, name):
    """CUDA pixels against the frames the reference's OWN fill.comp + tile.comp render (tests/golden/shader_frames.npz,
    produced by running the shader text on the CPU: oracle/pfshader.py). The north star's bound: 1/255 per channel."""
    import os

    want = np.load(os.path.join(scenes.GOLDEN, "shader_frames.npz"))[name]
    scene, _ = scenes.load_scene(scenes.golden_path(name))
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    d = np.abs(renderer.pixels().astype(int) - want.astype(int))
    print("%s: %d pixels differ from the reference shaders' frame, %d by more than 1" %
          (name, int((d.max(axis=2) > 0).sum()), int((d.max(axis=2) > 1).sum())))
    assert d.max() <= PIXEL_TOL, "max diff %d at %s" % (d.max(), np.unravel_index(d.argmax(), d.shape))


def test_tiger_4096_pixels_against_live_reference_shaders(renderer, area_lut):
    """The headline configuration (BASELINE.json configs[2]) against the reference's own fill.comp + tile.comp, run on the
    host CPU of this box through the prebuilt oracle/_ref/libpfshader.so (about 10 s; skipped where the library did not
    travel). 16.7 M pixels within 1/255."""
    pfshader = pytest.importorskip("pfshader")
    if not pfshader.available():
        pytest.skip("oracle/_ref/libpfshader.so not present")
    import pforacle

    scene, _ = scenes.load_scene(scenes.golden_path("tiger_4096_scene"))
    fr = pforacle.Frame(scene, area_lut)
    fr.render()  # geometry taps (pinned bit-exact to the reference's tiler) for the shaders' input buffers
    want, _, _ = pfshader.render_frame(scene, fr, area_lut)
    fr.close()
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    d = np.abs(renderer.pixels().astype(np.int16) - want.astype(np.int16)).max(axis=2)
    print("tiger 4096: %d of %d pixels differ from the reference shaders' frame" % (int((d > 0).sum()), d.size))
    assert d.max() <= PIXEL_TOL


FILTER_CASES = scenes.FILTER_CASES
refiltered_paints_scene = scenes.refiltered_paints_scene
TEXT_KERNEL_WIDE = scenes.TEXT_KERNEL_WIDE


@pytest.mark.parametrize("case", sorted(FILTER_CASES))
def test_text_and_color_matrix_filters_against_the_reference_shader(renderer, area_lut, case):
    """tile.comp's two remaining colour filters (filterText :136-227 without gamma correction, filterColorMatrix :394-404).
    Upstream never emits them (paint/palette.cpp:65-67), so no scene of the reference reaches them: the metadata of a fixture's
    image paint is rewritten and the frame compared with the reference's own tile.comp run on the CPU."""
    pfshader = pytest.importorskip("pfshader")
    if not pfshader.available():
        pytest.skip("oracle/_ref/libpfshader.so not present")
    import pforacle

    ctrl, params = FILTER_CASES[case]
    scene = refiltered_paints_scene(ctrl, params)
    plain, _ = scenes.load_scene(scenes.golden_path("paints_512"))
    fr = pforacle.Frame(plain, area_lut)
    fr.render()  # geometry taps only: they do not depend on the paints
    want, _, _ = pfshader.render_frame(scene, fr, area_lut)
    base, _, _ = pfshader.render_frame(plain, fr, area_lut)
    fr.close()
    assert (np.abs(want.astype(int) - base.astype(int)).max(axis=2) > 8).mean() > 0.02, "the filter must change the picture"
    renderer.set_scene(scene)
    renderer.draw(clear=True)
    d = np.abs(renderer.pixels().astype(np.int16) - want.astype(np.int16)).max(axis=2)
    print("%s: %d of %d pixels differ from the reference shader's frame" % (case, int((d > 0).sum()), d.size))
    assert d.max() <= PIXEL_TOL, "max diff %d at %s" % (d.max(), np.unravel_index(d.argmax(), d.shape))


def test_text_filter_with_gamma_correction_is_refused(renderer):
    """The reference binds a 1 x 1 dummy as the text filter's gamma LUT (d3d11/renderer.cpp:262-266): no defined result."""
    import pfcu

    scene = refiltered_paints_scene(0x120, [TEXT_KERNEL_WIDE, [1.0, 1.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0], [0.0] * 4, [0.0] * 4])
    with pytest.raises(pfcu.PfcuError, match="gamma"):
        renderer.set_scene(scene)
    renderer.set_scene(scenes.load_scene(scenes.golden_path("paints_512"))[0])


def test_config4_at_its_stated_size_against_the_oracle_fixture(renderer):
    """BASELINE.json configs[3] at FULL size -- 200,000 cubic blobs at 8192 x 8192 (16.7 M fills, 1.39 M masks: the
    only configuration that stresses the 24-bit tile ids and the multi-megabyte scans) -- against
    tests/golden/synthetic_200k_8192.npz, made by tests/golden/make_golden.py from oracle/pf_oracle.c (99 s of CPU):
    every geometry tap bit-exact (sha256 of lines, fills, tiles, z, sorted lists), pixels through the channel sums of
    every 16 x 16 tile (a pixel may differ by 1/255, so a tile's sum may move by a few units, never by a layer)."""
    import json
    import os

    fx = np.load(os.path.join(scenes.GOLDEN, "synthetic_200k_8192.npz"))
    want = json.loads(str(fx["digests"]))
    counts = json.loads(str(fx["counts"]))
    n_paths, size = scenes.CONFIG4
    scene = scenes.synthetic_scene(n_paths, size)
    renderer.set_scene(scene)
    st = renderer.draw(clear=True)
    assert st["overflow_flags"] == 0
    assert st["lines"] == counts["lines"] and st["fills"] == counts["fills"] and st["alpha_tiles"] == counts["alpha_tiles"]
    bid = int(scene["draw_batches"][0]["info"][0])
    tiles = renderer.tiles(bid)
    got = {"lines": scenes.digest(raw_sorted(renderer.lines(bid))), "fills": scenes.digest(renderer.fills(bid)),
           "z": scenes.digest(renderer.z(bid)), "tile_lists": scenes.digest(*renderer.tile_lists(bid)),
           "tiles_has_alpha": scenes.digest(tiles["alpha_tile_id"] >= 0)}
    for f in ("fill_count", "backdrop", "backdrop_delta", "backdrop_d3d9", "listed"):
        got["tiles_" + f] = scenes.digest(tiles[f])
    for k in sorted(want):
        assert got[k] == want[k], "config 4 @ %d^2: %s differs from the oracle" % (size, k)
    assert st["listed_after_cull"] == counts["listed_culled"] and st["max_list_len"] == counts["max_list"]
    px = renderer.pixels()
    d = np.abs(scenes.tile_sums(px).astype(np.int32) - fx["tile_sums"].astype(np.int32))
    print("config 4: %d of %d tiles differ in a channel sum, worst by %d" % (int((d.max(axis=2) > 0).sum()), d.shape[0] * d.shape[1], d.max()))
    # a 16 x 16 tile holds 256 pixels, each within 1/255: the bound is 256 per channel; what is observed is far below it
    assert d.max() <= 32, "a tile's channel sum is off by %d" % d.max()
    assert (d.max(axis=2) > 0).mean() < 0.02


def test_async_readback_into_pinned_memory(renderer, area_lut):
    """pfcu_read_target_async / pfcu_wait_read (= CommandEncoder::read_texture without the blocking fence wait,
    gpu/command_encoder.cpp:317-355): the copy is enqueued behind a SUBMITTED frame and lands in page-locked memory while
    another context renders; a new frame on the same context does not overwrite the target before the copy has read it."""
    import pfcu

    tiger, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    demo, _ = scenes.load_scene(scenes.golden_path("demo_full_512"))
    want = {}
    for name, scene in (("tiger", tiger), ("demo", demo)):
        renderer.set_scene(scene)
        renderer.draw(clear=True)
        want[name] = renderer.pixels()
    a, b = pfcu.Renderer(0, area_lut), pfcu.Renderer(0, area_lut)
    try:
        a.set_scene(tiger)
        b.set_scene(demo)
        bufs = {"a": [a.pinned_frame(), a.pinned_frame()], "b": [b.pinned_frame(), b.pinned_frame()]}
        for i in range(6):
            for r, key, name in ((a, "a", "tiger"), (b, "b", "demo")):
                buf = bufs[key][i & 1]
                buf[:] = 0x55
                r.draw(clear=True, clear_color=(1.0, 1.0, 1.0, 1.0) if i == 4 else (0.0, 0.0, 0.0, 0.0), upload=True, wait=False)
                r.read_async(buf)       # behind the frame that was just submitted
            for r, key, name in ((a, "a", "tiger"), (b, "b", "demo")):
                r.wait()
                r.wait_read()
                if i != 4:
                    assert np.array_equal(bufs[key][i & 1], want[name]), "%s frame %d" % (name, i)
                elif name == "tiger":
                    assert tuple(bufs[key][i & 1][0, 0]) == (255, 255, 255, 255)
        # read-back of a finished frame, then a new frame at once: the copy must see the OLD frame
        a.draw(clear=True)
        buf = bufs["a"][0]
        a.read_async(buf)
        a.draw(clear=True, clear_color=(1.0, 0.0, 0.0, 1.0))
        a.wait_read()
        assert np.array_equal(buf, want["tiger"])
        assert tuple(a.pixels()[0, 0]) == (255, 0, 0, 255)
        with pytest.raises(RuntimeError):
            a.L.pfcu_begin_frame(a.h) and None
            a.read_async(buf)  # a frame is open and not submitted
    finally:
        a.close()
        b.close()


def test_incremental_frames_redice_only_the_moved_path(renderer, area_lut):
    """PFCU_OPT_INCREMENTAL_DICE + pfcu_update_scene_range (what SceneEpoch / LastSceneInfo::draw_segment_ranges are for,
    core/scene.h:32-49): one tiger path moves (tests/golden/tiger_512_moved.npz, built by the reference front end from the
    same SVG with Outline::transform on draw path 145). The animated frame uploads and dices that path's 168 segments only,
    and its pixels, fills and tiles are those of a full render of the moved scene, bit for bit -- in both directions, and
    again after the path has moved back."""
    import pfcu

    base, _ = scenes.load_scene(scenes.golden_path("tiger_512"))
    moved, _ = scenes.load_scene(scenes.golden_path("tiger_512_moved"))
    bid = int(base["draw_batches"][0]["info"][0])
    want = {}
    for name, scene in (("base", base), ("moved", moved)):
        renderer.set_scene(scene)
        st = renderer.draw(clear=True)
        want[name] = (renderer.pixels(), renderer.fills(bid), renderer.tiles(bid), st)
    assert not np.array_equal(want["base"][0], want["moved"][0])
    full = want["base"][3]
    assert full["diced_segments"] == full["segments"] == 4399
    r = pfcu.Renderer(0, area_lut)
    try:
        r.set_incremental_dice(True)
        r.set_scene(base)
        st = r.draw(clear=True)                      # the base frame: everything is diced
        assert st["diced_segments"] == 4399
        st = r.draw(clear=True)                      # nothing changed: nothing is diced, nothing but metadata uploaded
        assert st["diced_segments"] == 0 and st["lines"] == full["lines"] and st["fills"] == full["fills"]
        assert np.array_equal(r.pixels(), want["base"][0])
        seq = [("moved", moved), ("base", base), ("moved", moved)]
        for i, (name, scene) in enumerate(seq):
            first_seg, n_seg = r.update_scene(scene)
            assert n_seg == 168
            st = r.draw(clear=True)
            assert st["diced_segments"] == 168, st
            assert st["uploaded_bytes"] < 0.6 * full["uploaded_bytes"]  # 4 KB of points instead of 121 KB of segments
            assert st["fills"] == want[name][3]["fills"] and st["alpha_tiles"] == want[name][3]["alpha_tiles"]
            assert np.array_equal(r.fills(bid), want[name][1]), "fills of frame %d" % i
            got_t, want_t = r.tiles(bid), want[name][2]
            for f in ("fill_count", "backdrop", "backdrop_delta", "listed"):
                assert np.array_equal(got_t[f], want_t[f]), f
            assert np.array_equal(r.pixels(), want[name][0]), "pixels of frame %d (%s)" % (i, name)
        # a full upload invalidates what was retained: the next frame dices everything again
        r.upload_segments(moved)
        st = r.draw(clear=True)
        assert st["diced_segments"] == 4399
        assert np.array_equal(r.pixels(), want["moved"][0])
    finally:
        r.close()

"""GPU tests of pfcu_stroke_to_fill against the REFERENCE's own stroker:
OutlineStrokeToFill::offset (pathfinder/core/stroke.cpp:124-167), run live through the prebuilt oracle/_ref/libpfref.so
(entry point pfref_stroke_outline in oracle/ref_harness/ref_harness.cpp). Every output point and flag must equal the
reference's BIT FOR BIT, for every cap x join combination, open and closed contours, lines, quadratics and cubics,
degenerate input included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CAPS = {"butt": 0, "square": 1, "round": 2}
JOINS = {"miter": 0, "bevel": 1, "round": 2}


def _contours(seed, n, size=600.0):
    """Random contours: polylines, quadratic and cubic splines, open and closed, a few degenerate ones."""
    rng = np.random.RandomState(seed)
    pts, flags, first, closed = [], [], [0], []
    for i in range(n):
        k = rng.randint(2, 9)
        c = rng.uniform(40, size - 40, 2)
        r = rng.uniform(5, 120)
        kind = i % 4
        p, f = [], []
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        on = np.stack([c[0] + r * np.cos(ang) * rng.uniform(0.5, 1.2, k), c[1] + r * np.sin(ang) * rng.uniform(0.5, 1.2, k)], 1)
        for j in range(k):
            p.append(on[j]); f.append(0)
            if j == k - 1:
                break
            if kind == 1:    # quadratic to the next point
                p.append((on[j] + on[j + 1]) * 0.5 + rng.normal(0, 0.3 * r, 2)); f.append(1)
            elif kind >= 2:  # cubic
                d = on[j + 1] - on[j]
                p.append(on[j] + d / 3 + rng.normal(0, 0.25 * r, 2)); f.append(1)
                p.append(on[j] + d * 2 / 3 + rng.normal(0, 0.25 * r, 2)); f.append(2)
        if i % 11 == 5:      # coincident points, a zero-length segment
            p.insert(1, p[0].copy()); f.insert(1, 0)
        if i % 13 == 7:      # a closed contour whose last point repeats the first (no closing line)
            p.append(p[0].copy()); f.append(0)
        pts.extend(p); flags.extend(f)
        first.append(len(pts))
        closed.append(1 if (i % 3 != 0) else 0)
    return (np.array(pts, "<f4"), np.array(flags, "u1"), np.array(first, "<u4"), np.array(closed, "u1"))


@pytest.fixture(scope="module")
def pfref():
    pfref = pytest.importorskip("pfref")
    if not pfref.available():
        pytest.skip("oracle/_ref/libpfref.so not present")
    return pfref


@pytest.fixture(scope="module")
def renderer(area_lut):
    import pfcu

    r = pfcu.Renderer(0, area_lut)
    yield r
    r.close()


def _equal_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view("<u4"), b.view("<u4"))


@pytest.mark.parametrize("join", sorted(JOINS))
@pytest.mark.parametrize("cap", sorted(CAPS))
def test_stroke_matches_the_reference_stroker_bit_for_bit(renderer, pfref, cap, join):
    pts, flags, first, closed = _contours(1234 + CAPS[cap] * 3 + JOINS[join], 160)
    for width, miter in ((1.0, 10.0), (7.5, 4.0), (31.0, 1.5)):
        want = pfref.stroke_outline(pts, flags, first, closed, width, CAPS[cap], JOINS[join], miter)
        got = renderer.stroke_to_fill(pts, flags, first, closed, np.zeros(len(closed), "<u4"), [(width, CAPS[cap], JOINS[join], miter)])
        assert np.array_equal(got[2], want[2]), "contour layout differs (width %g)" % width
        assert np.array_equal(got[1], want[1]), "point flags differ"
        assert _equal_bits(got[0], want[0]), "points differ: %d of %d" % (int((got[0].view("<u4") != want[0].view("<u4")).any(axis=1).sum()), len(want[0]))


def test_stroke_styles_per_contour_and_degenerate_input(renderer, pfref):
    """One batch, a different style per contour; empty and single-point contours; non-finite points."""
    pts, flags, first, closed = _contours(99, 40)
    # an empty contour, a single point, a contour with a NaN control point
    extra_pts = np.array([[5, 5], [50, 50], [60, np.nan], [70, 40], [90, 50]], "<f4")
    extra_flags = np.array([0, 0, 1, 2, 0], "u1")
    n0 = len(pts)
    pts = np.concatenate([pts, extra_pts]); flags = np.concatenate([flags, extra_flags])
    first = np.concatenate([first, [n0, n0 + 1, n0 + 5]]).astype("<u4")   # empty, single point, NaN cubic
    closed = np.concatenate([closed, [1, 0, 0]]).astype("u1")
    styles = [(2.0, 0, 0, 10.0), (9.0, 2, 2, 10.0), (14.0, 1, 1, 10.0), (5.0, 2, 0, 2.0)]
    idx = (np.arange(len(closed)) % len(styles)).astype("<u4")
    got = renderer.stroke_to_fill(pts, flags, first, closed, idx, styles)
    at_c = 0
    for i in range(len(closed)):
        w, cap, join, miter = styles[idx[i]]
        lo, hi = int(first[i]), int(first[i + 1])
        want = pfref.stroke_outline(pts[lo:hi], flags[lo:hi], [0, hi - lo], closed[i:i + 1], w, cap, join, miter)
        n_c = 2 if closed[i] else 1
        a, b = int(got[2][at_c]), int(got[2][at_c + n_c])
        assert np.array_equal(got[2][at_c:at_c + n_c + 1] - a, want[2]), "contour %d layout" % i
        assert np.array_equal(got[1][a:b], want[1]) and _equal_bits(got[0][a:b], want[0]), "contour %d" % i
        at_c += n_c
    assert at_c + 1 == len(got[2])


def test_stroked_outline_renders_like_the_cpu_stroked_one(renderer, pfref, area_lut):
    """End to end: blobs stroked on the GPU, pushed as fill paths (what Canvas::stroke_path does after the stroker,
    core/canvas.cpp:296-300) and rendered: the frame equals the one built from the reference stroker's outlines."""
    import scenes

    pts, flags, first, closed = _contours(7, 60, size=480.0)
    closed[:] = 1
    colors = np.array([[200, 30, 30, 255], [30, 160, 60, 200], [40, 60, 220, 140]], "u1")

    def scene_from(op, of, oc):
        paths = []
        for i in range(len(closed)):  # a closed contour -> outer + inner contour of one path
            contours = [(op[oc[2 * i + k]:oc[2 * i + k + 1]], of[oc[2 * i + k]:oc[2 * i + k + 1]]) for k in (0, 1)]
            contours = [c for c in contours if len(c[0])]
            if contours:
                paths.append({"contours": contours, "paint": i % 3, "fill_rule": 0, "opaque": False})
        return scenes.build_scene_from_outlines(512, 512, paths, colors)

    style = (6.0, 0, 2, 10.0)
    g = renderer.stroke_to_fill(pts, flags, first, closed, np.zeros(len(closed), "<u4"), [style])
    w = pfref.stroke_outline(pts, flags, first, closed, *style)
    frames = []
    for o in (g, w):
        renderer.set_scene(scene_from(o[0], o[1], o[2]))
        renderer.draw(clear=True)
        frames.append(renderer.pixels())
    assert frames[0].any() and np.array_equal(frames[0], frames[1])


def test_stroke_applies_the_canvas_transform_like_outline_transform(renderer, pfref):
    """pfcu_stroke_style::transform = the Outline::transform Canvas::push_path applies to the stroked outline
    (core/canvas.cpp:190, core/data/path.cpp:7-22): matrix * point + vector, one rounding per operation."""
    pts, flags, first, closed = _contours(5, 50)
    m = np.array([1.25, 0.5, -0.375, 0.8125, 17.5, -3.25], "<f4")
    want = pfref.stroke_outline(pts, flags, first, closed, 5.0, 2, 0, 10.0)
    got = renderer.stroke_to_fill(pts, flags, first, closed, np.zeros(len(closed), "<u4"), [(5.0, 2, 0, 10.0, tuple(m))])
    x, y = want[0][:, 0], want[0][:, 1]
    tx = (m[0] * x + m[2] * y) + m[4]   # Mat2 * Vec2F + vector (mat2.h:84-86, transform2.h:88-90), float32 throughout
    ty = (m[1] * x + m[3] * y) + m[5]
    assert _equal_bits(got[0], np.stack([tx, ty], 1).astype("<f4"))
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])


def test_dash_matches_the_reference_dasher_bit_for_bit(renderer, pfref):
    """pfcu_dash_outlines against OutlineDash (core/dash.cpp): several outlines of several contours each (the dash state
    runs on across the contours of an outline), different patterns and phases; then dash -> stroke, as Canvas::stroke_path
    chains them (core/canvas.cpp:286-296)."""
    rng = np.random.RandomState(3)
    pts, flags, first, closed = _contours(21, 90)
    n_c = len(closed)
    cuts = np.sort(rng.choice(np.arange(1, n_c), 19, replace=False))
    outline_first = np.concatenate([[0], cuts, [n_c]]).astype("<u4")
    patterns = [[10.0, 5.0], [3.0, 3.0, 12.0, 4.0], [25.0, 1.0], [0.5, 0.75], [40.0, 10.0, 5.0, 10.0]]
    dashes, dash_first, offsets = [], [0], []
    for o in range(len(outline_first) - 1):
        pat = patterns[o % len(patterns)]
        dashes.extend(pat)
        dash_first.append(len(dashes))
        offsets.append(0.0 if o % 3 else float(rng.uniform(0, sum(pat) * 0.9)))
    got = renderer.dash_outlines(pts, flags, first, closed, outline_first, dashes, dash_first, offsets)
    all_p, all_f, all_c = [], [], [0]
    for o in range(len(outline_first) - 1):
        c0, c1 = int(outline_first[o]), int(outline_first[o + 1])
        lo, hi = int(first[c0]), int(first[c1])
        want = pfref.dash_outline(pts[lo:hi], flags[lo:hi], first[c0:c1 + 1] - lo, closed[c0:c1],
                                  dashes[dash_first[o]:dash_first[o + 1]], offsets[o])
        g0, g1 = int(got[3][o]), int(got[3][o + 1])
        a, b = int(got[2][g0]), int(got[2][g1])
        assert np.array_equal(got[2][g0:g1 + 1] - a, want[2]), "outline %d: contour layout" % o
        assert np.array_equal(got[1][a:b], want[1]) and _equal_bits(got[0][a:b], want[0]), "outline %d" % o
        all_p.append(want[0]); all_f.append(want[1]); all_c.extend((want[2][1:] + all_c[-1]).tolist())
    assert got[2][-1] == len(got[0]) and len(got[0]) > 1000
    # dashed contours are open: stroke them with round caps, GPU against reference
    n_d = len(got[2]) - 1
    style = (2.5, 2, 1, 10.0)
    g = renderer.stroke_to_fill(got[0], got[1], got[2], np.zeros(n_d, "u1"), np.zeros(n_d, "<u4"), [style])
    w = pfref.stroke_outline(got[0], got[1], got[2], np.zeros(n_d, "u1"), *style)
    assert np.array_equal(g[2], w[2]) and np.array_equal(g[1], w[1]) and _equal_bits(g[0], w[0])


def _svg_stroke_chain_gpu(renderer, d, scale):
    """Canvas::stroke_path for every stroked shape of an SVG, on the GPU: dash (shapes with a dash array), stroke-to-fill,
    canvas transform. Returns per shape (points, flags, contour_first)."""
    ns = len(d["styles"])
    pts, fl, cf, cl, sf = d["points"], d["flags"], d["contour_first"], d["closed"], d["shape_first"]
    dashed = [s for s in range(ns) if d["dash_first"][s + 1] > d["dash_first"][s]]
    per_shape = {}
    if dashed:  # one batch of outlines = the dashed shapes
        sel_c = np.concatenate([np.arange(sf[s], sf[s + 1]) for s in dashed])
        lo = [int(cf[c]) for c in sel_c]
        hi = [int(cf[c + 1]) for c in sel_c]
        p = np.concatenate([pts[a:b] for a, b in zip(lo, hi)])
        f = np.concatenate([fl[a:b] for a, b in zip(lo, hi)])
        c_first = np.concatenate([[0], np.cumsum([b - a for a, b in zip(lo, hi)])]).astype("<u4")
        o_first = np.concatenate([[0], np.cumsum([sf[s + 1] - sf[s] for s in dashed])]).astype("<u4")
        dashes = np.concatenate([d["dashes"][d["dash_first"][s]:d["dash_first"][s + 1]] for s in dashed])
        d_first = np.concatenate([[0], np.cumsum([d["dash_first"][s + 1] - d["dash_first"][s] for s in dashed])]).astype("<u4")
        got = renderer.dash_outlines(p, f, c_first, cl[sel_c], o_first, dashes, d_first, [float(d["styles"][s][4]) for s in dashed])
        for k, s in enumerate(dashed):
            g0, g1 = int(got[3][k]), int(got[3][k + 1])
            a, b = int(got[2][g0]), int(got[2][g1])
            per_shape[s] = (got[0][a:b], got[1][a:b], got[2][g0:g1 + 1] - a, np.zeros(g1 - g0, "u1"))
    for s in range(ns):
        if s not in per_shape:
            c0, c1 = int(sf[s]), int(sf[s + 1])
            a, b = int(cf[c0]), int(cf[c1])
            per_shape[s] = (pts[a:b], fl[a:b], cf[c0:c1 + 1] - a, cl[c0:c1])
    # one stroke batch over every shape's contours, a style per shape
    all_p = np.concatenate([per_shape[s][0] for s in range(ns)])
    all_f = np.concatenate([per_shape[s][1] for s in range(ns)])
    counts = [len(per_shape[s][3]) for s in range(ns)]
    firsts, at = [0], 0
    for s in range(ns):
        firsts.extend((per_shape[s][2][1:] + at).tolist())
        at += len(per_shape[s][0])
    all_cl = np.concatenate([per_shape[s][3] for s in range(ns)])
    style_index = np.repeat(np.arange(ns), counts).astype("<u4")
    styles = [(float(st[0]), int(st[1]), int(st[2]), float(st[3]), (scale, 0, 0, scale, 0, 0)) for st in d["styles"]]
    g = renderer.stroke_to_fill(all_p, all_f, np.array(firsts, "<u4"), all_cl, style_index, styles)
    out, k = [], 0
    for s in range(ns):
        n_out = int(sum(2 if c else 1 for c in per_shape[s][3]))
        a, b = int(g[2][k]), int(g[2][k + n_out])
        out.append((g[0][a:b], g[1][a:b], g[2][k:k + n_out + 1] - a))
        k += n_out
    return out, g[3]


@pytest.mark.parametrize("asset,scale", [("tiger.svg", 4096 / 900.0), ("features.svg", 2048 / 720.0)])
def test_svg_strokes_from_unstroked_outlines(renderer, pfref, asset, scale):
    """Every stroked shape of tiger.svg / features.svg (caps, joins and a dashed path) as SvgScene hands it to
    Canvas::stroke_path (core/svg.cpp:155-196, oracle/ref_harness pfref_svg_stroke_inputs): the GPU chain dash -> stroke ->
    canvas transform equals OutlineDash -> OutlineStrokeToFill -> Outline::transform of the reference, bit for bit."""
    d = pfref.svg_stroke_inputs(pfref.asset(asset))
    got, gpu_ms = _svg_stroke_chain_gpu(renderer, d, np.float32(scale))
    s32 = np.float32(scale)
    n_points = 0
    for s in range(len(d["styles"])):
        c0, c1 = int(d["shape_first"][s]), int(d["shape_first"][s + 1])
        a, b = int(d["contour_first"][c0]), int(d["contour_first"][c1])
        p, f, cf, cl = d["points"][a:b], d["flags"][a:b], d["contour_first"][c0:c1 + 1] - a, d["closed"][c0:c1]
        dashes = d["dashes"][d["dash_first"][s]:d["dash_first"][s + 1]]
        if len(dashes):
            p, f, cf = pfref.dash_outline(p, f, cf, cl, dashes, float(d["styles"][s][4]))
            cl = np.zeros(len(cf) - 1, "u1")
        st = d["styles"][s]
        wp, wf, wc, _ = pfref.stroke_outline(p, f, cf, cl, float(st[0]), int(st[1]), int(st[2]), float(st[3]))
        x, y = wp[:, 0], wp[:, 1]
        wp = np.stack([(s32 * x + np.float32(0) * y) + np.float32(0), (np.float32(0) * x + s32 * y) + np.float32(0)], 1).astype("<f4")
        assert np.array_equal(got[s][2], wc) and np.array_equal(got[s][1], wf), "shape %d layout" % s
        assert _equal_bits(got[s][0], wp), "shape %d points" % s
        n_points += len(wp)
    print("%s: %d stroked shapes, %d output points, GPU stroke passes %.3f ms" % (asset, len(d["styles"]), n_points, gpu_ms))
    assert n_points > 1000

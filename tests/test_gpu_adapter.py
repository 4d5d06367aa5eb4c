"""GPU test of the whole drop-in chain in C++: the reference's own front end (SvgScene -> Canvas -> Palette ->
SceneBuilderD3D11, unmodified, compiled from /root/reference into oracle/_ref/libpfref_cuda.so) driving
pathfinder-cpp_b200/host/renderer_cuda.cpp, which calls the C-ABI -- i.e. exactly what Canvas::draw does with
RendererD3D11 upstream (core/canvas.cpp:557-567). The result must equal the frame the ctypes harness renders from
the committed fixture of the same scene (same inputs, same kernels) and sit within 1/255 of the oracle.

The library is prebuilt in the authoring container (it needs the reference sources) and travels to the GPU box;
the SVG bytes are linked into it as binary blobs, so nothing is read from /root/reference at run time.
"""
import ctypes as C
import os

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libpfref_cuda.so")


class FrameStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("batches", "segments", "lines", "fills", "alpha_tiles", "dense_tiles",
                                          "listed_tiles", "listed_after_cull", "fb_tiles", "max_list_len",
                                          "overflow_flags", "retries", "kernel_launches", "diced_segments",
                                          "uploaded_bytes")] + \
               [("reserved", C.c_uint32 * 1), ("gpu_ms", C.c_float)]


def _asset(lib, name):
    sym = "_binary_" + name.replace(".", "_")
    start = C.addressof(C.c_char.in_dll(lib, sym + "_start"))
    end = C.addressof(C.c_char.in_dll(lib, sym + "_end"))
    return C.string_at(start, end - start)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libpfref_cuda.so not built (needs /root/reference at build time)")
    L = C.CDLL(LIB)
    L.pfref_cuda_render_svg.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                        C.c_void_p, C.POINTER(FrameStats)]
    return L


@pytest.mark.parametrize("name,asset,size,native", [("tiger_512", "tiger.svg", 512, 900.0),
                                                    ("features_2048", "features.svg", 2048, 720.0)])
def test_reference_front_end_through_renderer_cuda(lib, area_lut, name, asset, size, native):
    import pfcu
    import pforacle

    svg = _asset(lib, asset)
    out = np.zeros((size, size, 4), "u1")
    st = FrameStats()
    rc = lib.pfref_cuda_render_svg(svg, len(svg), size, size, C.c_float(size / native), 0, 2,
                                   out.ctypes.data_as(C.c_void_p), C.byref(st))
    assert rc == 0
    assert st.retries == 0 and st.kernel_launches > 0  # second frame: steady state
    scene, _ = scenes.load_scene(scenes.golden_path(name))
    r = pfcu.Renderer(0, area_lut)
    r.set_scene(scene)
    stats = r.draw(clear=True)
    mine = r.pixels()
    r.close()
    for k in ("segments", "lines", "fills", "alpha_tiles", "dense_tiles"):
        assert getattr(st, k) == stats[k], k
    # same inputs, same kernels; the only freedom is the order of a tile's fills (scatter by atomic cursor), i.e. the
    # order of a float sum: last-bit differences may flip the rounding of a mask byte on a handful of pixels
    diff = np.abs(out.astype(int) - mine.astype(int))
    print("%s: adapter vs harness: %d channel values differ, max %d" % (name, int((diff > 0).sum()), int(diff.max())))
    assert diff.max() <= 1 and (diff > 0).sum() <= 1e-4 * diff.size, "C++ adapter and ctypes harness disagree"
    fr = pforacle.Frame(scene, area_lut)
    want = fr.render()
    fr.close()
    assert np.abs(out.astype(int) - want.astype(int)).max() <= 1


def _demo(lib, size, scale, features, frames, load_last):
    lib.pfref_cuda_render_demo.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.POINTER(FrameStats)]
    out = np.zeros((size, size, 4), "u1")
    st = FrameStats()
    rc = lib.pfref_cuda_render_demo(size, size, C.c_float(scale), features, 0, frames, load_last,
                                    out.ctypes.data_as(C.c_void_p), C.byref(st))
    assert rc == 0
    return out, st


def test_demo_primitives_through_renderer_cuda(lib, area_lut):
    """The demo's whole primitives scene (demo/common/app.cpp:21-101) through the C++ adapter: clip batches prepared in
    reverse, two blur passes through render-target pages, the image pattern's page + color_texture_info, the render-target
    pattern -- every branch of host/renderer_cuda.cpp that an SVG never takes."""
    import pfcu
    import pforacle

    out, st = _demo(lib, 512, 1.0, 0x3f, 2, 0)
    assert st.retries == 0 and st.kernel_launches > 0
    scene, _ = scenes.load_scene(scenes.golden_path("demo_full_512"))
    r = pfcu.Renderer(0, area_lut)
    r.set_scene(scene)
    stats = r.draw(clear=True)
    mine = r.pixels()
    r.close()
    for k in ("batches", "segments", "lines", "fills", "alpha_tiles", "dense_tiles"):
        assert getattr(st, k) == stats[k], k
    diff = np.abs(out.astype(int) - mine.astype(int))
    print("demo_full_512: adapter vs harness: %d channel values differ, max %d" % (int((diff > 0).sum()), int(diff.max())))
    assert diff.max() <= 1 and (diff > 0).sum() <= 1e-3 * diff.size
    fr = pforacle.Frame(scene, area_lut)
    want = fr.render()
    fr.close()
    assert np.abs(out.astype(int) - want.astype(int)).max() <= 1


def test_renderer_cuda_load_action_keeps_the_destination(lib, area_lut):
    """Renderer::draw(builder, clear_dst_texture = false) (core/renderer.h:83; LOAD_ACTION_LOAD, d3d11/renderer.cpp:382-386):
    the second frame blends over what the first left in the destination instead of starting from a zeroed target."""
    import pfcu

    out, _ = _demo(lib, 512, 1.0, 1 | 2 | 16, 2, 1)
    scene, _ = scenes.load_scene(scenes.golden_path("demo_clip_512"))
    r = pfcu.Renderer(0, area_lut)
    r.set_scene(scene)
    r.draw(clear=True)
    once = r.pixels()
    r.draw(clear=False)
    twice = r.pixels()
    r.close()
    assert np.abs(once.astype(int) - twice.astype(int)).max() > 8  # translucent layers: drawing twice must show
    assert np.abs(out.astype(int) - twice.astype(int)).max() <= 1

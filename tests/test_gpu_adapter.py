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
                                          "overflow_flags", "retries", "kernel_launches")] + \
               [("reserved", C.c_uint32 * 3), ("gpu_ms", C.c_float)]


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

// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// C API over the UNMODIFIED reference (compiled from /root/reference where it lies, see
// oracle/Makefile) running headless on the null device in null_device.h. It is used to
//   * run the reference front end (Canvas / SvgScene, core/svg.cpp:128-206) and its GPU-driven
//     scene builder (SceneBuilderD3D11::build, core/d3d11/scene_builder.cpp:209-223) to produce
//     exactly the inputs the CUDA path receives,
//   * run the reference hybrid CPU tiler (SceneBuilderD3D9::build, core/d3d9/scene_builder.cpp:97-109)
//     whose fills / tiles / z-buffers are the parity truth for dice / bin / propagate,
//   * time that CPU builder (bench.py --impl reference, cpu_baseline.kind == "reference").
// Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
// reference legs may load the resulting oracle/_ref/libpfref.so.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <nanosvg.h>

#include "demo_scene.h"
#include "null_device.h"
#include "pathfinder/common/io.h"
#include "pathfinder/common/logger.h"
#include "pathfinder/core/canvas.h"
#include "pathfinder/core/d3d11/scene_builder.h"
#include "pathfinder/core/d3d9/scene_builder.h"
#include "pathfinder/core/renderer.h"
#include "pathfinder/core/scene.h"
#include "pathfinder/core/dash.h"
#include "pathfinder/core/stroke.h"
#include "pathfinder/core/svg.h"

using namespace Pathfinder;

namespace {

/// Minimal concrete Renderer: the base class does all the palette-facing work
/// (core/renderer.cpp:13-251); the virtuals are never reached by the scene builders.
class CaptureRenderer : public Renderer {
public:
    CaptureRenderer(const std::shared_ptr<Device> &d, const std::shared_ptr<Queue> &q) : Renderer(d, q) {}
    void set_up_pipelines() override {}
    std::shared_ptr<Texture> get_dest_texture() override { return dest; }
    void set_dest_texture(const std::shared_ptr<Texture> &t) override { dest = t; }
    void draw(const std::shared_ptr<SceneBuilder> &, bool) override {}

    pfref::NullTexture *metadata_texture() {
        return static_cast<pfref::NullTexture *>(allocator->get_texture(metadata_texture_id).get());
    }
    pfref::NullTexture *area_lut_texture() {
        return static_cast<pfref::NullTexture *>(allocator->get_texture(area_lut_texture_id).get());
    }
    size_t page_count() const { return pattern_texture_pages.size(); }
    pfref::NullTexture *page_texture(size_t page) {
        if (page >= pattern_texture_pages.size() || !pattern_texture_pages[page]) return nullptr;
        return static_cast<pfref::NullTexture *>(allocator->get_texture(pattern_texture_pages[page]->texture_id_).get());
    }
    const std::vector<TextureLocation> &rt_locations() const { return render_target_locations; }

protected:
    TextureFormat mask_texture_format() const override { return TextureFormat::Rgba8Unorm; }
    std::shared_ptr<Texture> dest;
};

struct Handle {
    std::shared_ptr<pfref::NullDevice> device;
    std::shared_ptr<pfref::NullQueue> queue;
    std::shared_ptr<Canvas> canvas;
    std::shared_ptr<Scene> scene;
    std::shared_ptr<CaptureRenderer> renderer;
    std::shared_ptr<SceneBuilderD3D11> b11;
    std::shared_ptr<SceneBuilderD3D9> b9;
    int width = 0, height = 0;
};

Handle *new_handle(int width, int height) {
    Logger::set_global_level(Logger::Level::Error);
    auto *h = new Handle;
    h->width = width;
    h->height = height;
    h->device = std::make_shared<pfref::NullDevice>();
    h->queue = std::make_shared<pfref::NullQueue>();
    h->canvas = std::make_shared<Canvas>(Vec2I(width, height), h->device, h->queue, RenderMode::Hybrid);
    h->renderer = std::make_shared<CaptureRenderer>(h->device, h->queue);
    return h;
}

/// Palette::paints is private; count through the public, throwing accessor (paint/palette.cpp:98-104).
size_t paint_count(const Scene &scene) {
    size_t n = 0;
    try {
        for (;; n++) scene.palette.get_paint((uint32_t)n);
    } catch (const std::runtime_error &) {
    }
    return n;
}

template <typename T>
size_t copy_out(const std::vector<T> &v, void *out) {
    if (out && !v.empty()) memcpy(out, v.data(), v.size() * sizeof(T));
    return v.size();
}

const TileBatchDataD3D11 *get_batch11(Handle *h, int kind, int i, const DrawTileBatchD3D11 **draw = nullptr) {
    if (!h->b11) return nullptr;
    if (kind == 0) {
        if (i < 0 || i >= (int)h->b11->tile_batches.size()) return nullptr;
        if (draw) *draw = &h->b11->tile_batches[i];
        return &h->b11->tile_batches[i].tile_batch_data;
    }
    auto &pb = h->b11->clip_batches_d3d11->prepare_batches;
    if (i < 0 || i >= (int)pb.size()) return nullptr;
    return &pb[i];
}

} // namespace

extern "C" {

/// Scene = `svg` parsed by the reference front end under Canvas::set_transform(scale), view box width x height.
void *pfref_scene_from_svg(const char *svg, size_t len, int width, int height, float scale) {
    auto *h = new_handle(width, height);
    h->canvas->set_transform(Transform2::from_scale(Vec2F(scale, scale)));
    SvgScene svg_scene(std::string(svg, svg + len), *h->canvas);
    h->scene = svg_scene.get_scene();
    if (!h->scene) {
        delete h;
        return nullptr;
    }
    return h;
}

/// The non-SVG half of the reference demo (demo/common/app.cpp:21-101), every coordinate scaled by `scale`:
/// view-box clipped rect, clip circle, blurred shadow, image, gradient stroke, render-target pattern.
/// `features` is a bit mask selecting which of those blocks to include (1 rect, 2 clip, 4 shadow,
/// 8 image, 16 gradient stroke, 32 render target).
void *pfref_scene_demo(int width, int height, float scale, const char *img, size_t img_len, int features) {
    auto *h = new_handle(width, height);
    auto &canvas = h->canvas;
    pfref::draw_demo_scene(canvas, width, height, scale, img, img_len, features);
    h->scene = canvas->get_scene();
    return h;
}

/// Paint coverage the SVG / demo scenes lack: a linear-gradient background, one translucent rectangle per BlendMode
/// (effects.h:56-122; the separable and HSL modes go through tile.comp's composite(), :459-582), two radial gradients
/// (filterRadialGradient, :319-347) and a small image used as a repeating, unsmoothed pattern (REPEAT + NEAREST sampler
/// flags). `scale` multiplies every coordinate.
void *pfref_scene_paints(int width, int height, float scale, const char *img, size_t img_len) {
    auto *h = new_handle(width, height);
    auto &canvas = h->canvas;
    const float s = scale;
    {
        Path2d path;
        path.add_rect(RectF(Vec2F(0, 0), Vec2F(512 * s, 512 * s)));
        auto gradient = Gradient::linear(LineSegmentF({0, 0}, {512.0f * s, 512.0f * s}));
        gradient.add_color_stop(ColorU(230, 40, 60, 255), 0);
        gradient.add_color_stop(ColorU(40, 200, 90, 255), 0.5f);
        gradient.add_color_stop(ColorU(30, 60, 220, 255), 1);
        canvas->set_fill_paint(Paint::from_gradient(gradient));
        canvas->fill_path(path, FillRule::Winding);
    }
    const BlendMode modes[] = {BlendMode::SrcOver, BlendMode::Multiply, BlendMode::Screen, BlendMode::Overlay,
                               BlendMode::Darken, BlendMode::Lighten, BlendMode::ColorDodge, BlendMode::ColorBurn,
                               BlendMode::HardLight, BlendMode::SoftLight, BlendMode::Difference, BlendMode::Exclusion,
                               BlendMode::Hue, BlendMode::Saturation, BlendMode::Color, BlendMode::Luminosity,
                               BlendMode::Lighter, BlendMode::Xor, BlendMode::DestOver, BlendMode::SrcAtop};
    int k = 0;
    for (BlendMode mode : modes) {
        const float x = (20 + (k % 5) * 96) * s, y = (20 + (k / 5) * 70) * s;
        Path2d path;
        path.add_circle(Vec2F(x + 40 * s, y + 28 * s), 30 * s);
        path.add_rect(RectF(Vec2F(x + 30 * s, y + 10 * s), Vec2F(x + 90 * s, y + 50 * s)));
        canvas->set_global_composite_operation(mode);
        canvas->set_fill_paint(Paint::from_color(ColorU((uint8_t)(40 + 37 * k), (uint8_t)(220 - 29 * k), (uint8_t)(90 + 53 * k), (uint8_t)(k % 3 ? 200 : 255))));
        canvas->fill_path(path, k % 2 ? FillRule::EvenOdd : FillRule::Winding);
        k++;
    }
    canvas->set_global_composite_operation(BlendMode::SrcOver);
    for (int r = 0; r < 2; r++) {
        Path2d path;
        const Vec2F c((120 + 260 * r) * s, 400 * s);
        path.add_circle(c, 90 * s);
        auto gradient = Gradient::radial(LineSegmentF(c - Vec2F(20 * s * r, 0), c + Vec2F(30 * s * r, 10 * s * r)), Vec2F(10 * s, 90 * s));
        gradient.add_color_stop(ColorU(255, 255, 255, 255), 0);
        gradient.add_color_stop(ColorU(255, 180, 0, 160), 0.6f);
        gradient.add_color_stop(ColorU(0, 0, 0, 0), 1);
        canvas->set_fill_paint(Paint::from_gradient(gradient));
        canvas->fill_path(path, FillRule::Winding);
    }
    if (img && img_len) {
        auto image_buffer = ImageBuffer::from_memory(std::vector<char>(img, img + img_len), false);
        if (image_buffer) {
            auto image = std::make_shared<Image>(image_buffer->get_size(), image_buffer->to_rgba_pixels());
            auto pattern = Pattern::from_image(image);
            pattern.set_repeat_x(true);
            pattern.set_repeat_y(true);
            pattern.set_smoothing_enabled(false);
            pattern.apply_transform(Transform2::from_scale(Vec2F(0.11f * s, 0.11f * s)));
            Path2d path;
            path.add_rect(RectF(Vec2F(300 * s, 300 * s), Vec2F(500 * s, 500 * s)));
            canvas->set_fill_paint(Paint::from_pattern(pattern));
            canvas->fill_path(path, FillRule::Winding);
        }
    }
    h->scene = canvas->get_scene();
    return h;
}

void pfref_scene_free(void *p) { delete static_cast<Handle *>(p); }

/// Outline::transform on one draw path of the scene (an animated path: the incremental-frame fixtures).
/// xform = the six floats of Transform2(float[6]).
int pfref_scene_transform_draw_path(void *p, uint32_t index, const float xform[6]) {
    auto *h = static_cast<Handle *>(p);
    if (!h->scene || index >= h->scene->draw_paths.size()) return -1;
    float m[6] = {xform[0], xform[1], xform[2], xform[3], xform[4], xform[5]};
    h->scene->draw_paths[index].outline.transform(Transform2(m));
    return 0;
}

/// The same as a translation.
int pfref_scene_translate_draw_path(void *p, uint32_t index, float dx, float dy) {
    auto *h = static_cast<Handle *>(p);
    if (!h->scene || index >= h->scene->draw_paths.size()) return -1;
    h->scene->draw_paths[index].outline.transform(Transform2::from_translation(Vec2F(dx, dy)));
    return 0;
}

void pfref_scene_counts(void *p, uint32_t out[4]) {
    auto *h = static_cast<Handle *>(p);
    out[0] = (uint32_t)h->scene->draw_paths.size();
    out[1] = (uint32_t)h->scene->clip_paths.size();
    out[2] = (uint32_t)paint_count(*h->scene);
    out[3] = (uint32_t)h->scene->display_list.size();
}

void pfref_view_box(void *p, float out[4]) {
    auto vb = static_cast<Handle *>(p)->scene->get_view_box();
    out[0] = vb.left;
    out[1] = vb.top;
    out[2] = vb.right;
    out[3] = vb.bottom;
}

// ---------------------------------------------------------------- GPU-driven builder (inputs of the CUDA path)

int pfref_build_d3d11(void *p) {
    auto *h = static_cast<Handle *>(p);
    h->b11 = std::make_shared<SceneBuilderD3D11>();
    h->b11->build(h->scene.get(), h->renderer.get());
    return (int)h->b11->tile_batches.size();
}

size_t pfref_d3d11_points(void *p, int which, float *out) {
    auto *h = static_cast<Handle *>(p);
    auto &s = which == 0 ? h->b11->built_segments.draw_segments : h->b11->built_segments.clip_segments;
    return copy_out(s.points, out);
}

size_t pfref_d3d11_indices(void *p, int which, uint32_t *out) {
    auto *h = static_cast<Handle *>(p);
    auto &s = which == 0 ? h->b11->built_segments.draw_segments : h->b11->built_segments.clip_segments;
    return copy_out(s.indices, out);
}

/// kind 0 = draw batches in draw order; kind 1 = clip prepare batches in STORAGE order (the renderer
/// submits them in reverse, core/d3d11/renderer.cpp:318-327).
int pfref_d3d11_num_batches(void *p, int kind) {
    auto *h = static_cast<Handle *>(p);
    if (!h->b11) return 0;
    return kind == 0 ? (int)h->b11->tile_batches.size()
                     : (int)h->b11->clip_batches_d3d11->prepare_batches.size();
}

/// info: [0] batch_id [1] path_count [2] tile_count [3] segment_count [4] n_backdrops [5] path_source
/// [6] clip_batch_id or ~0 [7] color page or ~0 [8] sampling flags [9] composite op [10] render target or ~0
/// [11] render target page or ~0, [12..15] render target rect.
int pfref_d3d11_batch_info(void *p, int kind, int i, uint32_t info[16]) {
    auto *h = static_cast<Handle *>(p);
    const DrawTileBatchD3D11 *draw = nullptr;
    auto *b = get_batch11(h, kind, i, &draw);
    if (!b) return -1;
    for (int k = 0; k < 16; k++) info[k] = ~0u;
    info[0] = b->batch_id;
    info[1] = b->path_count;
    info[2] = b->tile_count;
    info[3] = b->segment_count;
    info[4] = (uint32_t)b->prepare_info.backdrops.size();
    info[5] = b->path_source == PathSource::Draw ? 0 : 1;
    if (b->clipped_path_info) info[6] = b->clipped_path_info->clip_batch_id;
    if (draw) {
        if (draw->color_texture_info) {
            info[7] = draw->color_texture_info->page_id;
            info[8] = draw->color_texture_info->sampling_flags.value;
            info[9] = (uint32_t)draw->color_texture_info->composite_op;
        }
        if (draw->render_target_id) {
            info[10] = draw->render_target_id->render_target;
            auto &locs = h->renderer->rt_locations();
            if (info[10] < locs.size()) {
                info[11] = locs[info[10]].page;
                info[12] = locs[info[10]].rect.left;
                info[13] = locs[info[10]].rect.top;
                info[14] = locs[info[10]].rect.right;
                info[15] = locs[info[10]].rect.bottom;
            }
        }
    }
    return 0;
}

/// which: 0 backdrops (12 B), 1 propagate metadata (48 B), 2 dice metadata (16 B), 3 tile path info (16 B),
/// 4 transform (6 floats: m11 m21 m12 m22 tx ty). Returns the element count.
size_t pfref_d3d11_batch_array(void *p, int kind, int i, int which, void *out) {
    auto *h = static_cast<Handle *>(p);
    auto *b = get_batch11(h, kind, i);
    if (!b) return 0;
    auto &pi = b->prepare_info;
    switch (which) {
        case 0: return copy_out(pi.backdrops, out);
        case 1: return copy_out(pi.propagate_metadata, out);
        case 2: return copy_out(pi.dice_metadata, out);
        case 3: return copy_out(pi.tile_path_info, out);
        case 4: {
            if (out) {
                float *f = static_cast<float *>(out);
                auto &t = pi.transform;
                f[0] = t.m11();
                f[1] = t.m21();
                f[2] = t.m12();
                f[3] = t.m22();
                f[4] = t.m13();
                f[5] = t.m23();
            }
            return 6;
        }
    }
    return 0;
}

/// RGBA16F paint metadata texels as uploaded by Renderer::upload_texture_metadata (core/renderer.cpp:167-251).
/// Returns the number of halfs in the whole 1280 x 512 texture; `rows_used` = rows holding paints.
size_t pfref_metadata_texels(void *p, uint16_t *out, uint32_t *rows_used) {
    auto *h = static_cast<Handle *>(p);
    auto *t = h->renderer->metadata_texture();
    size_t n = t->bytes.size() / 2;
    if (out) memcpy(out, t->bytes.data(), t->bytes.size());
    if (rows_used) {
        size_t paints = paint_count(*h->scene);
        *rows_used = (uint32_t)((paints + 127) / 128);
    }
    return n;
}

void pfref_area_lut(void *p, uint8_t *out, int wh[2]) {
    auto *h = static_cast<Handle *>(p);
    auto *t = h->renderer->area_lut_texture();
    wh[0] = t->get_size().x;
    wh[1] = t->get_size().y;
    if (out) memcpy(out, t->bytes.data(), t->bytes.size());
}

int pfref_num_pages(void *p) { return (int)static_cast<Handle *>(p)->renderer->page_count(); }

int pfref_page(void *p, int page, uint8_t *out, int wh[2]) {
    auto *t = static_cast<Handle *>(p)->renderer->page_texture(page);
    if (!t) return -1;
    wh[0] = t->get_size().x;
    wh[1] = t->get_size().y;
    if (out) memcpy(out, t->bytes.data(), t->bytes.size());
    return 0;
}

// ---------------------------------------------------------------- hybrid CPU builder (the parity truth)

int pfref_build_d3d9(void *p) {
    auto *h = static_cast<Handle *>(p);
    h->b9 = std::make_shared<SceneBuilderD3D9>();
    h->b9->build(h->scene.get(), h->renderer.get());
    return (int)h->b9->tile_batches.size();
}

/// Fill = {u16 from_x, from_y, to_x, to_y; u32 alpha tile id} (core/d3d9/data/gpu_data.h:11-16), 12 B.
size_t pfref_d3d9_fills(void *p, void *out) {
    static_assert(sizeof(Fill) == 12, "Fill layout");
    return copy_out(static_cast<Handle *>(p)->b9->pending_fills, out);
}

int pfref_d3d9_num_batches(void *p) { return (int)static_cast<Handle *>(p)->b9->tile_batches.size(); }

/// TileObjectPrimitive (core/d3d9/data/gpu_data.h:19-27), 16 B.
size_t pfref_d3d9_batch_tiles(void *p, int i, void *out) {
    static_assert(sizeof(TileObjectPrimitive) == 16, "TileObjectPrimitive layout");
    return copy_out(static_cast<Handle *>(p)->b9->tile_batches[i].tiles, out);
}

/// Clip (core/d3d9/data/gpu_data.h:30-35), 16 B.
size_t pfref_d3d9_batch_clips(void *p, int i, void *out) {
    static_assert(sizeof(Clip) == 16, "Clip layout");
    return copy_out(static_cast<Handle *>(p)->b9->tile_batches[i].clips, out);
}

size_t pfref_d3d9_batch_z(void *p, int i, uint32_t *out, int rect[4]) {
    auto &z = static_cast<Handle *>(p)->b9->tile_batches[i].z_buffer_data;
    rect[0] = z.rect.left;
    rect[1] = z.rect.top;
    rect[2] = z.rect.right;
    rect[3] = z.rect.bottom;
    return copy_out(z.data, out);
}

/// info: [0] color page or ~0 [1] sampling flags [2] composite op [3] render target or ~0.
void pfref_d3d9_batch_info(void *p, int i, uint32_t info[4]) {
    auto &b = static_cast<Handle *>(p)->b9->tile_batches[i];
    for (int k = 0; k < 4; k++) info[k] = ~0u;
    if (b.color_texture_info) {
        info[0] = b.color_texture_info->page_id;
        info[1] = b.color_texture_info->sampling_flags.value;
        info[2] = (uint32_t)b.color_texture_info->composite_op;
    }
    if (b.render_target_id) info[3] = b.render_target_id->render_target;
}

/// Wall time (ms) of `iters` calls of SceneBuilderD3D9::build, each written to out_ms[i].
/// Threads: 4, hard-coded by the reference (core/d3d9/scene_builder.cpp:13).
void pfref_time_d3d9_build(void *p, int iters, double *out_ms) {
    auto *h = static_cast<Handle *>(p);
    SceneBuilderD3D9 b;
    for (int i = 0; i < iters; i++) {
        auto t0 = std::chrono::steady_clock::now();
        b.build(h->scene.get(), h->renderer.get());
        auto t1 = std::chrono::steady_clock::now();
        out_ms[i] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
}

/// Same for the GPU-driven builder's CPU part (segment packing + batch metadata).
void pfref_time_d3d11_build(void *p, int iters, double *out_ms) {
    auto *h = static_cast<Handle *>(p);
    SceneBuilderD3D11 b;
    for (int i = 0; i < iters; i++) {
        auto t0 = std::chrono::steady_clock::now();
        b.build(h->scene.get(), h->renderer.get());
        auto t1 = std::chrono::steady_clock::now();
        out_ms[i] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
}

// ---------------------------------------------------------------- stroke-to-fill (the oracle of pfcu_stroke_to_fill)

/// OutlineStrokeToFill::offset (core/stroke.cpp:124-167) on an outline given as arrays: points (x, y), flags (PointFlag:
/// 0 on-curve, 1 control point 0, 2 control point 1), contour i = points [contour_first[i], contour_first[i + 1]), closed[i].
/// One style for the whole outline. Query the sizes with null outputs: returns the number of output points and stores the
/// number of output contours; out_contour_first gets n_out_contours + 1 entries. Returns the CPU time in out_ms (optional).
size_t pfref_stroke_outline(const float *points, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                            uint32_t n_contours, float line_width, int line_cap, int line_join, float miter_limit,
                            float *out_points, uint8_t *out_flags, uint32_t *out_contour_first, uint32_t *n_out_contours,
                            double *out_ms) {
    Logger::set_global_level(Logger::Level::Silence);
    Outline outline;
    for (uint32_t i = 0; i < n_contours; i++) {
        Contour contour;
        for (uint32_t k = contour_first[i]; k < contour_first[i + 1]; k++)
            contour.push_point(Vec2F(points[2 * k], points[2 * k + 1]), (PointFlag)flags[k], true);
        contour.closed = closed[i] != 0;
        outline.contours.push_back(contour);  // (Outline::push_contour drops empty contours; the batch API keeps them)
    }
    StrokeStyle style;
    style.line_width = line_width;
    style.line_cap = (LineCap)line_cap;
    style.line_join = (LineJoin)line_join;
    style.miter_limit = miter_limit;
    const auto t0 = std::chrono::steady_clock::now();
    OutlineStrokeToFill stroker(outline, style);
    stroker.offset();
    const Outline &result = stroker.output;
    if (out_ms) *out_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    size_t n = 0;
    uint32_t c = 0;
    for (const auto &contour : result.contours) {
        if (out_contour_first) out_contour_first[c] = (uint32_t)n;
        for (size_t k = 0; k < contour.points.size(); k++) {
            if (out_points) {
                out_points[2 * (n + k)] = contour.points[k].x;
                out_points[2 * (n + k) + 1] = contour.points[k].y;
            }
            if (out_flags) out_flags[n + k] = (uint8_t)contour.flags[k];
        }
        n += contour.points.size();
        c++;
    }
    if (out_contour_first) out_contour_first[c] = (uint32_t)n;
    if (n_out_contours) *n_out_contours = c;
    return n;
}

/// OutlineDash::dash + into_outline (core/dash.cpp:49-65) on an outline given as arrays (as pfref_stroke_outline).
size_t pfref_dash_outline(const float *points, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                          uint32_t n_contours, const float *dashes, uint32_t n_dashes, float offset, float *out_points,
                          uint8_t *out_flags, uint32_t *out_contour_first, uint32_t *n_out_contours) {
    Logger::set_global_level(Logger::Level::Silence);
    Outline outline;
    for (uint32_t i = 0; i < n_contours; i++) {
        Contour contour;
        for (uint32_t k = contour_first[i]; k < contour_first[i + 1]; k++)
            contour.push_point(Vec2F(points[2 * k], points[2 * k + 1]), (PointFlag)flags[k], true);
        contour.closed = closed[i] != 0;
        outline.contours.push_back(contour);
    }
    OutlineDash dasher(outline, std::vector<float>(dashes, dashes + n_dashes), offset);
    dasher.dash();
    const Outline result = dasher.into_outline();
    size_t n = 0;
    uint32_t c = 0;
    for (const auto &contour : result.contours) {
        if (out_contour_first) out_contour_first[c] = (uint32_t)n;
        for (size_t k = 0; k < contour.points.size(); k++) {
            if (out_points) {
                out_points[2 * (n + k)] = contour.points[k].x;
                out_points[2 * (n + k) + 1] = contour.points[k].y;
            }
            if (out_flags) out_flags[n + k] = (uint8_t)contour.flags[k];
        }
        n += contour.points.size();
        c++;
    }
    if (out_contour_first) out_contour_first[c] = (uint32_t)n;
    if (n_out_contours) *n_out_contours = c;
    return n;
}

/// What SvgScene hands to Canvas::stroke_path for every shape with a visible stroke (core/svg.cpp:155-196): the shape's
/// outline BEFORE dashing, stroking and the canvas transform, and its stroke style. Shapes in document order. Two-call
/// pattern: null outputs to query. counts[4] = {shapes, contours, points, dash values}. style rows: line_width, line_cap,
/// line_join, miter_limit, dash_offset.
int pfref_svg_stroke_inputs(const char *svg, size_t len, uint32_t counts[4], float *points, uint8_t *flags, uint32_t *contour_first,
                            uint8_t *closed, uint32_t *shape_first, float *styles, float *dashes, uint32_t *dash_first) {
    std::string copy(svg, svg + len);
    NSVGimage *image = nsvgParse((char *)copy.data(), "px", 96);
    if (!image) return -1;
    uint32_t n_shapes = 0, n_contours = 0, n_points = 0, n_dashes = 0;
    if (contour_first) contour_first[0] = 0;
    if (shape_first) shape_first[0] = 0;
    if (dash_first) dash_first[0] = 0;
    for (NSVGshape *shape = image->shapes; shape != nullptr; shape = shape->next) {
        // Canvas::stroke_path strokes when the stroke paint is visible and the width positive (canvas.cpp:281)
        if (shape->stroke.type == NSVG_PAINT_NONE || !(shape->strokeWidth > 0)) continue;
        if (shape->stroke.type == NSVG_PAINT_COLOR && ColorU(shape->stroke.color).a_ == 0) continue;
        Path2d path;
        for (NSVGpath *p = shape->paths; p != nullptr; p = p->next) {
            path.move_to(p->pts[0], p->pts[1]);
            for (int i = 0; i < p->npts - 3; i += 3) {
                float *q = &p->pts[i * 2];
                path.cubic_to(q[2], q[3], q[4], q[5], q[6], q[7]);
            }
            if (p->closed) path.close_path();
        }
        const Outline outline = path.into_outline();
        for (const auto &contour : outline.contours) {
            for (size_t k = 0; k < contour.points.size(); k++) {
                if (points) {
                    points[2 * (n_points + k)] = contour.points[k].x;
                    points[2 * (n_points + k) + 1] = contour.points[k].y;
                }
                if (flags) flags[n_points + k] = (uint8_t)contour.flags[k];
            }
            n_points += (uint32_t)contour.points.size();
            if (closed) closed[n_contours] = contour.closed ? 1 : 0;
            n_contours++;
            if (contour_first) contour_first[n_contours] = n_points;
        }
        if (styles) {
            float *st = styles + 5 * n_shapes;
            st[0] = shape->strokeWidth;
            st[1] = shape->strokeLineCap == NSVG_CAP_ROUND ? 2.f : shape->strokeLineCap == NSVG_CAP_SQUARE ? 1.f : 0.f;
            st[2] = shape->strokeLineJoin == NSVG_JOIN_ROUND ? 2.f : shape->strokeLineJoin == NSVG_JOIN_BEVEL ? 1.f : 0.f;
            st[3] = shape->miterLimit;
            st[4] = shape->strokeDashOffset;
        }
        for (int k = 0; k < shape->strokeDashCount; k++) {
            if (dashes) dashes[n_dashes] = shape->strokeDashArray[k];
            n_dashes++;
        }
        n_shapes++;
        if (shape_first) shape_first[n_shapes] = n_contours;
        if (dash_first) dash_first[n_shapes] = n_dashes;
    }
    nsvgDelete(image);
    counts[0] = n_shapes;
    counts[1] = n_contours;
    counts[2] = n_points;
    counts[3] = n_dashes;
    return 0;
}

// ---------------------------------------------------------------- embedded assets (linked from /root/reference/assets)

extern const char _binary_tiger_svg_start[], _binary_tiger_svg_end[];
extern const char _binary_features_svg_start[], _binary_features_svg_end[];
extern const char _binary_sea_png_start[], _binary_sea_png_end[];

const char *pfref_asset(const char *name, size_t *len) {
    std::string n(name);
    if (n == "tiger.svg") {
        *len = _binary_tiger_svg_end - _binary_tiger_svg_start;
        return _binary_tiger_svg_start;
    }
    if (n == "features.svg") {
        *len = _binary_features_svg_end - _binary_features_svg_start;
        return _binary_features_svg_start;
    }
    if (n == "sea.png") {
        *len = _binary_sea_png_end - _binary_sea_png_start;
        return _binary_sea_png_start;
    }
    *len = 0;
    return nullptr;
}

} // extern "C"

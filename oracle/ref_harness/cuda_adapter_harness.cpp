// TEST INFRASTRUCTURE. End-to-end check of the drop-in boundary: the UNMODIFIED reference front end
// (SvgScene / Canvas / Palette / SceneBuilderD3D11, compiled from /root/reference where it lies) driving the product's
// C++ adapter (pathfinder-cpp_b200/host/renderer_cuda.cpp) exactly as Canvas::draw does (core/canvas.cpp:557-567):
//     scene_builder->build(scene, renderer);  renderer->draw(scene_builder, clear);
// Built into oracle/_ref/libpfref_cuda.so together with the reference objects; links libpfcu.so. Used only by
// tests/test_gpu_adapter.py.
#include <cstring>
#include <memory>
#include <string>

#include "demo_scene.h"
#include "host_device.h"
#include "pathfinder/common/logger.h"
#include "pathfinder/core/canvas.h"
#include "pathfinder/core/d3d11/scene_builder.h"
#include "pathfinder/core/svg.h"
#include "renderer_cuda.h"

using namespace Pathfinder;

extern "C" {

/// Renders `svg` scaled by `scale` into width x height RGBA8 (out, tightly packed) on CUDA device `cuda_device`.
/// `frames` > 1 redraws the same scene (steady state: no re-allocation). stats: pfcu_frame_stats of the last frame.
/// Returns 0, or a negative value on failure.
int pfref_cuda_render_svg(const char *svg, size_t len, int width, int height, float scale, int cuda_device, int frames,
                          uint8_t *out, pfcu_frame_stats *stats) {
    try {
        Logger::set_global_level(Logger::Level::Error);
        auto device = std::make_shared<HostDevice>();
        auto queue = std::make_shared<HostQueue>();
        // The front end: a Canvas records the scene (its own renderer is never asked to draw).
        auto canvas = std::make_shared<Canvas>(Vec2I(width, height), device, queue, RenderMode::Hybrid);
        canvas->set_transform(Transform2::from_scale(Vec2F(scale, scale)));
        SvgScene svg_scene(std::string(svg, svg + len), *canvas);
        auto scene = svg_scene.get_scene();
        if (!scene) return -1;

        // What Canvas::Canvas does for RenderMode::GpuDriven (core/canvas.cpp:164-175), with the CUDA renderer.
        auto renderer = std::make_shared<RendererCuda>(device, queue, cuda_device);
        auto scene_builder = std::make_shared<SceneBuilderD3D11>();
        renderer->set_up_pipelines();
        auto dest = device->create_texture({Vec2I(width, height), TextureFormat::Rgba8Unorm}, "dest texture");
        renderer->set_dest_texture(dest);

        for (int f = 0; f < (frames > 0 ? frames : 1); f++) {
            // Canvas::draw (core/canvas.cpp:557-567)
            scene_builder->build(scene.get(), renderer.get());
            renderer->draw(scene_builder, true);
            renderer->reset();
        }
        renderer->read_dest_texture();
        if (out) memcpy(out, static_cast<HostTexture *>(dest.get())->bytes.data(), (size_t)width * height * 4);
        if (stats) *stats = renderer->last_frame_stats();
        return 0;
    } catch (const std::exception &e) {
        Logger::error(std::string("pfref_cuda_render_svg: ") + e.what());
        return -2;
    }
}

/// The demo's primitives scene (demo_scene.h: clip circle, blurred shadow through two render-target passes, image pattern,
/// gradient stroke, render-target pattern) through RendererCuda: clip batches in reverse, pattern pages, color_texture_info.
/// `frames` frames are drawn; when `load_last` is set, the last one is drawn with clear_dst_texture = false (the reference's
/// LOAD_ACTION_LOAD: it blends over what the previous frame left in the destination, d3d11/renderer.cpp:382-386).
int pfref_cuda_render_demo(int width, int height, float scale, int features, int cuda_device, int frames, int load_last,
                           uint8_t *out, pfcu_frame_stats *stats) {
    try {
        Logger::set_global_level(Logger::Level::Error);
        auto device = std::make_shared<HostDevice>();
        auto queue = std::make_shared<HostQueue>();
        auto canvas = std::make_shared<Canvas>(Vec2I(width, height), device, queue, RenderMode::Hybrid);
        extern const char _binary_sea_png_start[], _binary_sea_png_end[];
        pfref::draw_demo_scene(canvas, width, height, scale, _binary_sea_png_start,
                               (size_t)(_binary_sea_png_end - _binary_sea_png_start), features);
        auto scene = canvas->get_scene();
        auto renderer = std::make_shared<RendererCuda>(device, queue, cuda_device);
        auto scene_builder = std::make_shared<SceneBuilderD3D11>();
        renderer->set_up_pipelines();
        auto dest = device->create_texture({Vec2I(width, height), TextureFormat::Rgba8Unorm}, "dest texture");
        renderer->set_dest_texture(dest);
        const int n = frames > 0 ? frames : 1;
        for (int f = 0; f < n; f++) {
            scene_builder->build(scene.get(), renderer.get());
            renderer->draw(scene_builder, !(load_last && f == n - 1));
            renderer->reset();
        }
        renderer->read_dest_texture();
        if (out) memcpy(out, static_cast<HostTexture *>(dest.get())->bytes.data(), (size_t)width * height * 4);
        if (stats) *stats = renderer->last_frame_stats();
        return 0;
    } catch (const std::exception &e) {
        Logger::error(std::string("pfref_cuda_render_demo: ") + e.what());
        return -2;
    }
}

}  // extern "C"

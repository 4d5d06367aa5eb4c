// TEST INFRASTRUCTURE. The non-SVG half of the reference demo (demo/common/app.cpp:21-101) drawn through the reference's
// own Canvas API, every coordinate scaled by `scale`: view-box clipped rect, clip circle, blurred shadow, image, gradient
// stroke, render-target pattern. `features` is a bit mask selecting which of those blocks to include (1 rect, 2 clip,
// 4 shadow, 8 image, 16 gradient stroke, 32 render target). Shared by ref_harness.cpp (fixtures) and
// cuda_adapter_harness.cpp (the drop-in test of RendererCuda).
#pragma once

#include <memory>
#include <vector>

#include "pathfinder/common/io.h"
#include "pathfinder/core/canvas.h"
#include "pathfinder/core/scene.h"

namespace pfref {

using namespace Pathfinder;

inline void draw_demo_scene(const std::shared_ptr<Canvas> &canvas, int width, int height, float scale, const char *img,
                            size_t img_len, int features) {
    float s = scale;
    auto canvas_size = Vec2I(width, height);

    canvas->save_state();
    if (features & 1) {
        Path2d path;
        path.add_rect(RectF(Vec2F(400 * s, 400 * s), canvas_size.to_f32() + Vec2F(100 * s)));
        canvas->set_fill_paint(Paint::from_color(ColorU::red()));
        canvas->fill_path(path, FillRule::Winding);
    }
    if (features & 2) {
        Path2d path;
        path.add_circle(Vec2F(180.0f * s, 180.0f * s), 180 * s);
        canvas->clip_path(path, FillRule::Winding);
    }
    if (features & 4) {
        canvas->set_shadow_color(ColorU::white());
        canvas->set_shadow_blur(16 * s);
        canvas->set_shadow_offset({8 * s, 8 * s});
        canvas->set_shadow_strength(2.0);
    }
    if ((features & 8) && img && img_len) {
        auto image_buffer = ImageBuffer::from_memory(std::vector<char>(img, img + img_len), false);
        if (image_buffer) {
            auto image = std::make_shared<Image>(image_buffer->get_size(), image_buffer->to_rgba_pixels());
            Vec2F pos = {10 * s, 20 * s};
            canvas->draw_image(image, RectF(pos, pos + image->size.to_f32() * s));
        }
    }
    if (features & 16) {
        Path2d path;
        path.move_to(260.0f * s, 260.0f * s);
        path.line_to(460.0f * s, 260.0f * s);
        path.line_to(460.0f * s, 460.0f * s);
        path.line_to(260.0f * s, 460.0f * s);
        path.close_path();
        canvas->set_line_width(10.0f * s);
        auto gradient = Gradient::linear(LineSegmentF({260.0f * s, 260.0f * s}, {460.0f * s, 460.0f * s}));
        gradient.add_color_stop(ColorU::red(), 0);
        gradient.add_color_stop(ColorU::transparent_black(), 1);
        gradient.add_color_stop(ColorU::blue(), 0.5);
        gradient.add_color_stop(ColorU::green(), 0.25);
        canvas->set_stroke_paint(Paint::from_gradient(gradient));
        canvas->stroke_path(path);
    }
    canvas->restore_state();
    if (features & 32) {
        auto render_target_size = Vec2I(int(400 * s), int(300 * s));
        auto render_target_desc = RenderTargetDesc{render_target_size, "sub render target"};
        auto render_target_id = canvas->get_scene()->push_render_target(render_target_desc);
        Path2d path;
        path.add_circle({200 * s, 150 * s}, 50 * s);
        path.add_line({}, {200 * s, 150 * s});
        canvas->set_line_width(10.0f * s);
        canvas->set_stroke_paint(Paint::from_color(ColorU::red()));
        canvas->stroke_path(path);
        canvas->get_scene()->pop_render_target();
        auto pos = Vec2F(100 * s, 50 * s);
        canvas->draw_render_target(render_target_id, {pos, pos + render_target_size.to_f32()});
    }
}

}  // namespace pfref

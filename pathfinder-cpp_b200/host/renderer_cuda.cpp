#include "renderer_cuda.h"

#include <stdexcept>
#include <string>

#include "host_device.h"
#include "pathfinder/common/logger.h"
#include "pathfinder/core/scene.h"

namespace Pathfinder {

namespace {

// The reference's records and the C-ABI's are the same PODs (include/pfcu.h cites each one).
static_assert(sizeof(BackdropInfoD3D11) == sizeof(pfcu_backdrop_info), "BackdropInfoD3D11");
static_assert(sizeof(PropagateMetadataD3D11) == sizeof(pfcu_propagate_metadata), "PropagateMetadataD3D11");
static_assert(sizeof(DiceMetadataD3D11) == sizeof(pfcu_dice_metadata), "DiceMetadataD3D11");
static_assert(sizeof(TilePathInfoD3D11) == sizeof(pfcu_tile_path_info), "TilePathInfoD3D11");
static_assert(sizeof(SegmentIndicesD3D11) == 8 && sizeof(Vec2F) == 8, "segment records");

bool check(int rc, const char *what) {
    if (rc == PFCU_OK) return true;
    Logger::error(std::string(what) + ": " + pfcu_last_error());
    return false;
}

HostTexture *host(const std::shared_ptr<Texture> &t) { return static_cast<HostTexture *>(t.get()); }

} // namespace

RendererCuda::RendererCuda(const std::shared_ptr<Device> &device, const std::shared_ptr<Queue> &queue, int cuda_device)
    : Renderer(device, queue) {
    if (!std::dynamic_pointer_cast<HostDevice>(device)) {
        throw std::runtime_error("RendererCuda needs a HostDevice (pathfinder-cpp_b200/host/host_device.h)");
    }
    if (pfcu_create(cuda_device, &ctx_) != PFCU_OK) {
        throw std::runtime_error(std::string("pfcu_create: ") + pfcu_last_error());  // no CPU fallback
    }
}

RendererCuda::~RendererCuda() { pfcu_destroy(ctx_); }

void RendererCuda::set_up_pipelines() {
    auto *lut = host(allocator->get_texture(area_lut_texture_id));
    const auto size = lut->get_size();
    lut_uploaded_ = check(pfcu_set_area_lut(ctx_, lut->bytes.data(), size.x, size.y), "area LUT");
}

std::shared_ptr<Texture> RendererCuda::get_dest_texture() { return dest_texture_; }

void RendererCuda::set_dest_texture(const std::shared_ptr<Texture> &new_texture) { dest_texture_ = new_texture; }

void RendererCuda::upload_paint_state() {
    // Palette::build_paint_info wrote these through the base class (core/renderer.cpp:46-114,167-251) during
    // SceneBuilderD3D11::build; only what changed since the last frame crosses PCIe.
    auto *metadata = host(allocator->get_texture(metadata_texture_id));
    if (metadata->version != metadata_version_) {
        const uint32_t rows = (uint32_t)metadata->get_size().y;
        if (check(pfcu_upload_paint_metadata(ctx_, reinterpret_cast<const uint16_t *>(metadata->bytes.data()), rows),
                  "paint metadata")) {
            metadata_version_ = metadata->version;
        }
    }
    for (size_t page = 0; page < pattern_texture_pages.size(); page++) {
        if (!pattern_texture_pages[page]) continue;
        auto *tex = host(allocator->get_texture(pattern_texture_pages[page]->texture_id_));
        if (!tex) continue;
        auto it = page_versions_.find(page);
        if (it != page_versions_.end() && it->second == tex->version) continue;
        const auto size = tex->get_size();
        if (it == page_versions_.end() && !check(pfcu_alloc_page(ctx_, (uint32_t)page, size.x, size.y), "pattern page")) continue;
        if (tex->get_format() == TextureFormat::Rgba8Unorm &&
            check(pfcu_upload_page_region(ctx_, (uint32_t)page, 0, 0, size.x, size.y, tex->bytes.data()), "pattern page")) {
            page_versions_[page] = tex->version;
        }
    }
}

bool RendererCuda::prepare(const TileBatchDataD3D11 &batch) {
    const auto &info = batch.prepare_info;
    pfcu_batch_desc d{};
    d.batch_id = batch.batch_id;
    d.path_count = batch.path_count;
    d.tile_count = batch.tile_count;
    d.segment_count = batch.segment_count;
    d.column_count = (uint32_t)info.backdrops.size();
    d.path_source = batch.path_source == PathSource::Clip ? 1 : 0;
    d.clip_batch_id = batch.clipped_path_info ? (int32_t)batch.clipped_path_info->clip_batch_id : -1;
    d.backdrops = reinterpret_cast<const pfcu_backdrop_info *>(info.backdrops.data());
    d.propagate_metadata = reinterpret_cast<const pfcu_propagate_metadata *>(info.propagate_metadata.data());
    d.dice_metadata = reinterpret_cast<const pfcu_dice_metadata *>(info.dice_metadata.data());
    d.tile_path_info = reinterpret_cast<const pfcu_tile_path_info *>(info.tile_path_info.data());
    const auto &tr = info.transform;
    const float t[6] = {tr.m11(), tr.m21(), tr.m12(), tr.m22(), tr.m13(), tr.m23()};
    memcpy(d.transform, t, sizeof(t));
    return check(pfcu_prepare_batch(ctx_, &d), "prepare tiles");
}

void RendererCuda::draw(const std::shared_ptr<SceneBuilder> &_scene_builder, bool _clear_dst_texture) {
    clear_dest_texture = _clear_dst_texture;
    auto *scene_builder = static_cast<SceneBuilderD3D11 *>(_scene_builder.get());
    if (scene_builder->built_segments.draw_segments.points.empty()) return;  // renderer.cpp:309-311
    if (!dest_texture_) {
        Logger::error("RendererCuda: no destination texture");
        return;
    }
    if (!lut_uploaded_) set_up_pipelines();

    const auto size = dest_texture_->get_size();
    const RectF vb = scene_builder->get_scene()->get_view_box();
    const float view_box[4] = {vb.left, vb.top, vb.right, vb.bottom};
    if (!check(pfcu_set_target(ctx_, size.x, size.y, nullptr, 0, view_box), "set target")) return;
    upload_paint_state();

    // RenderCommand::UploadSceneD3D11 (renderer.cpp:314, 346-350)
    auto &segs = scene_builder->built_segments;
    if (!check(pfcu_upload_scene(ctx_, 0, reinterpret_cast<const float *>(segs.draw_segments.points.data()),
                                 (uint32_t)segs.draw_segments.points.size(),
                                 reinterpret_cast<const uint32_t *>(segs.draw_segments.indices.data()),
                                 (uint32_t)segs.draw_segments.indices.size()), "upload scene") ||
        !check(pfcu_upload_scene(ctx_, 1, reinterpret_cast<const float *>(segs.clip_segments.points.data()),
                                 (uint32_t)segs.clip_segments.points.size(),
                                 reinterpret_cast<const uint32_t *>(segs.clip_segments.indices.data()),
                                 (uint32_t)segs.clip_segments.indices.size()), "upload scene")) {
        return;
    }
    if (!check(pfcu_begin_frame(ctx_), "begin frame")) return;

    // Prepare clip tiles, last batch first (renderer.cpp:318-327).
    if (scene_builder->clip_batches_d3d11) {
        auto &prepare_batches = scene_builder->clip_batches_d3d11->prepare_batches;
        for (auto iter = prepare_batches.rbegin(); iter != prepare_batches.rend(); ++iter) {
            if (iter->path_count > 0 && !prepare(*iter)) break;
        }
    }
    // prepare_and_draw_tiles (renderer.cpp:352-448)
    const float clear_color[4] = {0.f, 0.f, 0.f, 0.f};
    for (auto &batch : scene_builder->tile_batches) {
        if (!prepare(batch.tile_batch_data)) break;
        int color_page = -1;
        uint32_t sampling_flags = 0;
        if (batch.color_texture_info) {
            color_page = (int)batch.color_texture_info->page_id;
            sampling_flags = batch.color_texture_info->sampling_flags.value;
        }
        int rc;
        if (batch.render_target_id == nullptr) {
            rc = pfcu_draw_batch(ctx_, batch.tile_batch_data.batch_id, -1, color_page, sampling_flags,
                                 clear_dest_texture ? 1 : 0, clear_color);
            clear_dest_texture = false;
        } else {
            // a render target is a region of a pattern page (core/renderer.cpp:65-89); pfcu renders whole pages
            const auto location = get_render_target_location(*batch.render_target_id);
            rc = pfcu_draw_batch(ctx_, batch.tile_batch_data.batch_id, (int)location.page, color_page, sampling_flags, 1,
                                 clear_color);
        }
        if (!check(rc, "draw tiles")) break;
    }
    if (pfcu_end_frame(ctx_, &stats_) == PFCU_ERR_OVERFLOW) {
        Logger::error("Ran out of space for the frame after retries!");  // renderer.cpp:551,575
    }
}

void *RendererCuda::device_pixels(size_t *pitch_bytes) const { return pfcu_target_device_ptr(ctx_, pitch_bytes); }

void RendererCuda::read_dest_texture() {
    if (!dest_texture_) return;
    auto *tex = host(dest_texture_);
    if (check(pfcu_read_target(ctx_, tex->bytes.data()), "read target")) tex->version++;
}

} // namespace Pathfinder

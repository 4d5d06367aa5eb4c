// RendererCuda: the reference's Renderer interface (pathfinder/core/renderer.h:57-140) implemented on libpfcu.so.
//
// It is what RendererD3D11 (pathfinder/core/d3d11/renderer.h, renderer.cpp:113-1079) is in the reference: the object
// Canvas hands its SceneBuilderD3D11 to once per frame (core/canvas.cpp:557-567). Same names, same argument meaning,
// same error behaviour (Logger::error + early return; std::runtime_error for API misuse). Everything above it --
// Canvas, Path2d, Paint, Scene, SceneBuilderD3D11 and their data contracts (core/d3d11/gpu_data.h:54-205) -- is used
// unchanged; everything below it (pathfinder/gpu, gpu_mem, shaders/d3d11) is replaced by the C-ABI in include/pfcu.h.
#pragma once

#include <cstdint>
#include <memory>
#include <unordered_map>
#include <vector>

#include "pathfinder/core/d3d11/scene_builder.h"
#include "pathfinder/core/renderer.h"

extern "C" {
#include "pfcu.h"
}

namespace Pathfinder {

class RendererCuda : public Renderer {
public:
    /// `device` / `queue` serve the base class's texture bookkeeping only and must be a HostDevice / HostQueue
    /// (host_device.h); `cuda_device` is the ordinal the frame is rendered on.
    RendererCuda(const std::shared_ptr<Device> &device, const std::shared_ptr<Queue> &queue, int cuda_device = 0);

    ~RendererCuda() override;

    /// RendererD3D11::set_up_pipelines (renderer.cpp:129-300) has nothing to compile here: it hands the area LUT the base
    /// class decoded (core/renderer.cpp:17-21) to the CUDA context.
    void set_up_pipelines() override;

    std::shared_ptr<Texture> get_dest_texture() override;

    /// The texture only carries the size here; pixels live in device memory (device_pixels()) and are copied into the
    /// texture's host bytes by read_dest_texture().
    void set_dest_texture(const std::shared_ptr<Texture> &new_texture) override;

    /// RendererD3D11::draw (renderer.cpp:302-336): upload scene, prepare clip batches in reverse, prepare + draw every
    /// draw batch. One CUDA stream, no mid-frame read-back; returns after the frame's counters came back.
    void draw(const std::shared_ptr<SceneBuilder> &scene_builder, bool clear_dst_texture) override;

    /// Device pointer + pitch of the RGBA8 destination (zero-copy consumers: interop, NCCL, encoders).
    void *device_pixels(size_t *pitch_bytes) const;

    /// Copies the destination into the dest texture's host memory (what CommandEncoder::read_texture does upstream).
    void read_dest_texture();

    /// Counters of the last frame (segments, lines, fills, alpha tiles, ...).
    const pfcu_frame_stats &last_frame_stats() const { return stats_; }

    pfcu_ctx *context() const { return ctx_; }

protected:
    TextureFormat mask_texture_format() const override { return TextureFormat::Rgba8Unorm; }

private:
    void upload_paint_state();
    bool prepare(const TileBatchDataD3D11 &batch);

    pfcu_ctx *ctx_ = nullptr;
    std::shared_ptr<Texture> dest_texture_;
    pfcu_frame_stats stats_{};
    uint64_t metadata_version_ = ~0ull;
    std::unordered_map<uint64_t, uint64_t> page_versions_;  // pattern page id -> uploaded version
    bool lut_uploaded_ = false;
};

} // namespace Pathfinder

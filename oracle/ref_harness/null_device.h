// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// A host-memory "null" implementation of the reference's abstract GPU layer
// (pathfinder/gpu/device.h:24-147, command_encoder.h:158-273, queue.h:11-25), just enough
// to let the reference's own front end (Canvas / SvgScene / Palette) and both of its scene
// builders (core/d3d9/scene_builder.cpp, core/d3d11/scene_builder.cpp) run headless, with no
// Vulkan/GL. Texture writes are executed on the host so the harness can read back what the
// reference uploads (area LUT, RGBA16F paint metadata, gradient/image pages).
//
// This file is compiled only into oracle/_ref/libpfref.so, against the reference headers where
// they lie under /root/reference. Nothing here is copied from the reference.
#pragma once

#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "pathfinder/gpu/device.h"
#include "pathfinder/gpu/queue.h"

namespace pfref {

using namespace Pathfinder;

class NullBuffer : public Buffer {
public:
    explicit NullBuffer(const BufferDescriptor &desc) : Buffer(desc), bytes(desc.size) {}
    void upload_via_mapping(size_t data_size, size_t offset, const void *data) override {
        if (offset + data_size <= bytes.size()) memcpy(bytes.data() + offset, data, data_size);
    }
    void download_via_mapping(size_t data_size, size_t offset, void *data) override {
        if (offset + data_size <= bytes.size()) memcpy(data, bytes.data() + offset, data_size);
    }
    std::vector<uint8_t> bytes;
};

class NullTexture : public Texture {
public:
    explicit NullTexture(const TextureDescriptor &desc) : Texture(desc), bytes(desc.byte_size(), 0) {}
    std::string label() const { return label_; }
    std::vector<uint8_t> bytes;
    uint64_t serial = 0;
};

class NullSampler : public Sampler {
public:
    explicit NullSampler(const SamplerDescriptor &d) : Sampler(d) {}
};

class NullDescriptorSetLayout : public DescriptorSetLayout {
public:
    explicit NullDescriptorSetLayout(const std::vector<DescriptorLayout> &l) : DescriptorSetLayout(l) {}
};

class NullDescriptorSet : public DescriptorSet {
public:
    explicit NullDescriptorSet(const std::shared_ptr<DescriptorSetLayout> &l) : DescriptorSet(l) {}
};

class NullRenderPass : public RenderPass {
public:
    NullRenderPass(AttachmentLoadOp op, const std::string &label) {
        load_op_ = op;
        label_ = label;
    }
};

class NullFramebuffer : public Framebuffer {
public:
    explicit NullFramebuffer(const std::shared_ptr<Texture> &t) : Framebuffer(t) {}
};

class NullShaderModule : public ShaderModule {};

class NullRenderPipeline : public RenderPipeline {
public:
    NullRenderPipeline(const std::vector<VertexInputAttributeDescription> &a, const BlendState &b, std::string l)
        : RenderPipeline(a, b, std::move(l)) {}
};

class NullComputePipeline : public ComputePipeline {};

class NullCommandEncoder : public CommandEncoder {
public:
    NullCommandEncoder(const std::shared_ptr<Device> &device, const std::string &label) {
        device_ = device;
        label_ = label;
    }

    /// Execute the recorded host-visible writes (buffer + texture uploads) on the CPU.
    void execute_writes() {
        for (auto &cmd : commands_) {
            if (cmd.type == CommandType::WriteTexture) {
                auto &a = cmd.args.write_texture;
                auto *tex = static_cast<NullTexture *>(a.texture);
                auto *staging = static_cast<NullBuffer *>(a.staging_buffer);
                size_t px = get_pixel_size(tex->get_format());
                size_t tex_w = tex->get_size().x;
                const uint8_t *src = staging->bytes.data() + a.staging_offset;
                for (uint32_t row = 0; row < a.height; row++) {
                    memcpy(tex->bytes.data() + ((size_t)(a.offset_y + row) * tex_w + a.offset_x) * px,
                           src + (size_t)row * a.width * px,
                           (size_t)a.width * px);
                }
            } else if (cmd.type == CommandType::WriteBuffer) {
                auto &a = cmd.args.write_buffer;
                auto *buf = static_cast<NullBuffer *>(a.buffer);
                auto *staging = static_cast<NullBuffer *>(a.staging_buffer);
                if (buf && staging && a.offset + a.data_size <= buf->bytes.size()) {
                    memcpy(buf->bytes.data() + a.offset, staging->bytes.data() + a.staging_offset, a.data_size);
                }
            }
        }
        commands_.clear();
    }

protected:
    bool prepare() override { return true; }
};

class NullDevice : public Device {
public:
    NullDevice() : Device(1) { backend_type = BackendType::Vulkan; }

    std::shared_ptr<Framebuffer> create_framebuffer(const std::shared_ptr<RenderPass> &,
                                                    const std::shared_ptr<Texture> &texture,
                                                    const std::string &) override {
        return std::make_shared<NullFramebuffer>(texture);
    }
    std::shared_ptr<Buffer> create_buffer(const BufferDescriptor &desc, const std::string &) override {
        return std::make_shared<NullBuffer>(desc);
    }
    std::shared_ptr<Texture> create_texture(const TextureDescriptor &desc, const std::string &label) override {
        auto t = std::make_shared<NullTexture>(desc);
        t->set_label(label);
        t->serial = next_serial++;
        textures.push_back(t);
        return t;
    }
    std::shared_ptr<Sampler> create_sampler(SamplerDescriptor d) override { return std::make_shared<NullSampler>(d); }
    std::shared_ptr<CommandEncoder> create_command_encoder(const std::string &label) override {
        return std::make_shared<NullCommandEncoder>(shared_from_this(), label);
    }
    std::shared_ptr<DescriptorSetLayout> create_descriptor_set_layout(
        const std::vector<DescriptorLayout> &descriptors) override {
        return std::make_shared<NullDescriptorSetLayout>(descriptors);
    }
    std::shared_ptr<DescriptorSet> create_descriptor_set(std::shared_ptr<DescriptorSetLayout> layout) override {
        return std::make_shared<NullDescriptorSet>(layout);
    }
    std::shared_ptr<RenderPass> create_render_pass(TextureFormat, AttachmentLoadOp op, const std::string &l) override {
        return std::make_shared<NullRenderPass>(op, l);
    }
    std::shared_ptr<RenderPass> create_swap_chain_render_pass(TextureFormat, AttachmentLoadOp op) override {
        return std::make_shared<NullRenderPass>(op, "swap chain");
    }
    std::shared_ptr<ShaderModule> create_shader_module(const std::shared_ptr<Shader> &, const std::string &) override {
        return std::make_shared<NullShaderModule>();
    }
    std::shared_ptr<ShaderModule> create_shader_module(const std::vector<char> &,
                                                       ShaderStage,
                                                       const std::string &) override {
        return std::make_shared<NullShaderModule>();
    }
    std::shared_ptr<RenderPipeline> create_render_pipeline(const std::shared_ptr<ShaderModule> &,
                                                           const std::shared_ptr<ShaderModule> &,
                                                           const std::vector<VertexInputAttributeDescription> &a,
                                                           BlendState b,
                                                           const std::shared_ptr<DescriptorSetLayout> &,
                                                           TextureFormat,
                                                           const std::string &l) override {
        return std::make_shared<NullRenderPipeline>(a, b, l);
    }
    std::shared_ptr<ComputePipeline> create_compute_pipeline(const std::shared_ptr<ShaderModule> &,
                                                             const std::shared_ptr<DescriptorSetLayout> &,
                                                             const std::string &) override {
        return std::make_shared<NullComputePipeline>();
    }
    std::shared_ptr<Fence> create_fence(const std::string &label) override {
        auto f = std::make_shared<Fence>();
        f->label = label;
        return f;
    }
    void *map_staging(const StagingAllocation &allocation) override {
        auto *b = static_cast<NullBuffer *>(allocation.buffer.get());
        return b->bytes.data() + allocation.offset;
    }
    size_t get_aligned_uniform_size(size_t original_size) override { return (original_size + 255) & ~size_t(255); }

    /// Every texture ever created, in creation order (weak: the allocator owns them).
    std::vector<std::weak_ptr<NullTexture>> textures;
    uint64_t next_serial = 0;

protected:
    std::shared_ptr<Buffer> create_staging_buffer(size_t size) override {
        return std::make_shared<NullBuffer>(
            BufferDescriptor{BufferType::Storage, size, MemoryProperty::HostVisibleAndCoherent});
    }
};

class NullQueue : public Queue {
public:
    void submit(const std::shared_ptr<CommandEncoder> &encoder, const std::shared_ptr<Fence> &) override {
        auto *e = static_cast<NullCommandEncoder *>(encoder.get());
        e->execute_writes();
        e->invoke_callbacks();
    }
};

} // namespace pfref

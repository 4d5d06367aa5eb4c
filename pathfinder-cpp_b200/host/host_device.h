// Host-memory implementation of the reference's abstract GPU layer (pathfinder/gpu/device.h:24-147,
// command_encoder.h:158-273, queue.h:11-25) for applications that render through RendererCuda.
//
// On this path the GPU work does not go through pathfinder/gpu at all -- RendererCuda talks to libpfcu.so -- but the
// reference's front end still creates and fills a few textures through its Device: the area LUT, the RGBA16F paint
// metadata and the gradient / image pattern pages (core/renderer.cpp:13-251, core/paint/palette.cpp:120-179). This
// device keeps those in host memory, executes the recorded texture / buffer writes on submit, and counts a version per
// texture so that RendererCuda re-uploads only what changed. No windowing, no pipelines, no shaders.
//
// Compiled against the reference's headers (add pathfinder-cpp_b200/host/*.cpp to the pathfinder target and link
// libpfcu.so). Nothing here is copied from the reference.
#pragma once

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "pathfinder/gpu/device.h"
#include "pathfinder/gpu/queue.h"

namespace Pathfinder {

class HostBuffer : public Buffer {
public:
    explicit HostBuffer(const BufferDescriptor &desc) : Buffer(desc), bytes(desc.size) {}
    void upload_via_mapping(size_t data_size, size_t offset, const void *data) override {
        if (offset + data_size <= bytes.size()) memcpy(bytes.data() + offset, data, data_size);
    }
    void download_via_mapping(size_t data_size, size_t offset, void *data) override {
        if (offset + data_size <= bytes.size()) memcpy(data, bytes.data() + offset, data_size);
    }
    std::vector<uint8_t> bytes;
};

/// A texture whose texels live in host memory. `version` increases with every write.
class HostTexture : public Texture {
public:
    explicit HostTexture(const TextureDescriptor &desc) : Texture(desc), bytes(desc.byte_size(), 0) {}
    std::vector<uint8_t> bytes;
    uint64_t version = 0;
};

namespace host_detail {
class Sampler_ : public Sampler {
public:
    explicit Sampler_(const SamplerDescriptor &d) : Sampler(d) {}
};
class SetLayout_ : public DescriptorSetLayout {
public:
    explicit SetLayout_(const std::vector<DescriptorLayout> &l) : DescriptorSetLayout(l) {}
};
class Set_ : public DescriptorSet {
public:
    explicit Set_(const std::shared_ptr<DescriptorSetLayout> &l) : DescriptorSet(l) {}
};
class Pass_ : public RenderPass {
public:
    Pass_(AttachmentLoadOp op, const std::string &label) {
        load_op_ = op;
        label_ = label;
    }
};
class Framebuffer_ : public Framebuffer {
public:
    explicit Framebuffer_(const std::shared_ptr<Texture> &t) : Framebuffer(t) {}
};
class Module_ : public ShaderModule {};
class RenderPipeline_ : public RenderPipeline {
public:
    RenderPipeline_(const std::vector<VertexInputAttributeDescription> &a, const BlendState &b, std::string l)
        : RenderPipeline(a, b, std::move(l)) {}
};
class ComputePipeline_ : public ComputePipeline {};
} // namespace host_detail

class HostCommandEncoder : public CommandEncoder {
public:
    HostCommandEncoder(const std::shared_ptr<Device> &device, const std::string &label) {
        device_ = device;
        label_ = label;
    }

    /// Perform the recorded uploads (CommandEncoder::write_texture / write_buffer stage their data, command_encoder.cpp:205-315).
    void execute_writes() {
        for (auto &cmd : commands_) {
            if (cmd.type == CommandType::WriteTexture) {
                auto &a = cmd.args.write_texture;
                auto *tex = static_cast<HostTexture *>(a.texture);
                auto *staging = static_cast<HostBuffer *>(a.staging_buffer);
                const size_t px = get_pixel_size(tex->get_format());
                const size_t tex_w = (size_t)tex->get_size().x;
                const uint8_t *src = staging->bytes.data() + a.staging_offset;
                for (uint32_t row = 0; row < a.height; row++)
                    memcpy(tex->bytes.data() + ((size_t)(a.offset_y + row) * tex_w + a.offset_x) * px,
                           src + (size_t)row * a.width * px, (size_t)a.width * px);
                tex->version++;
            } else if (cmd.type == CommandType::WriteBuffer) {
                auto &a = cmd.args.write_buffer;
                auto *buf = static_cast<HostBuffer *>(a.buffer);
                auto *staging = static_cast<HostBuffer *>(a.staging_buffer);
                if (buf && staging && a.offset + a.data_size <= buf->bytes.size())
                    memcpy(buf->bytes.data() + a.offset, staging->bytes.data() + a.staging_offset, a.data_size);
            }
        }
        commands_.clear();
    }

protected:
    bool prepare() override { return true; }
};

class HostDevice : public Device {
public:
    HostDevice() : Device(1) { backend_type = BackendType::Vulkan; }

    std::shared_ptr<Framebuffer> create_framebuffer(const std::shared_ptr<RenderPass> &, const std::shared_ptr<Texture> &t,
                                                    const std::string &) override {
        return std::make_shared<host_detail::Framebuffer_>(t);
    }
    std::shared_ptr<Buffer> create_buffer(const BufferDescriptor &desc, const std::string &) override {
        return std::make_shared<HostBuffer>(desc);
    }
    std::shared_ptr<Texture> create_texture(const TextureDescriptor &desc, const std::string &label) override {
        auto t = std::make_shared<HostTexture>(desc);
        t->set_label(label);
        return t;
    }
    std::shared_ptr<Sampler> create_sampler(SamplerDescriptor d) override { return std::make_shared<host_detail::Sampler_>(d); }
    std::shared_ptr<CommandEncoder> create_command_encoder(const std::string &label) override {
        return std::make_shared<HostCommandEncoder>(shared_from_this(), label);
    }
    std::shared_ptr<DescriptorSetLayout> create_descriptor_set_layout(const std::vector<DescriptorLayout> &d) override {
        return std::make_shared<host_detail::SetLayout_>(d);
    }
    std::shared_ptr<DescriptorSet> create_descriptor_set(std::shared_ptr<DescriptorSetLayout> layout) override {
        return std::make_shared<host_detail::Set_>(layout);
    }
    std::shared_ptr<RenderPass> create_render_pass(TextureFormat, AttachmentLoadOp op, const std::string &l) override {
        return std::make_shared<host_detail::Pass_>(op, l);
    }
    std::shared_ptr<RenderPass> create_swap_chain_render_pass(TextureFormat, AttachmentLoadOp op) override {
        return std::make_shared<host_detail::Pass_>(op, "swap chain");
    }
    std::shared_ptr<ShaderModule> create_shader_module(const std::shared_ptr<Shader> &, const std::string &) override {
        return std::make_shared<host_detail::Module_>();
    }
    std::shared_ptr<ShaderModule> create_shader_module(const std::vector<char> &, ShaderStage, const std::string &) override {
        return std::make_shared<host_detail::Module_>();
    }
    std::shared_ptr<RenderPipeline> create_render_pipeline(const std::shared_ptr<ShaderModule> &,
                                                           const std::shared_ptr<ShaderModule> &,
                                                           const std::vector<VertexInputAttributeDescription> &a, BlendState b,
                                                           const std::shared_ptr<DescriptorSetLayout> &, TextureFormat,
                                                           const std::string &l) override {
        return std::make_shared<host_detail::RenderPipeline_>(a, b, l);
    }
    std::shared_ptr<ComputePipeline> create_compute_pipeline(const std::shared_ptr<ShaderModule> &,
                                                             const std::shared_ptr<DescriptorSetLayout> &,
                                                             const std::string &) override {
        return std::make_shared<host_detail::ComputePipeline_>();
    }
    std::shared_ptr<Fence> create_fence(const std::string &label) override {
        auto f = std::make_shared<Fence>();
        f->label = label;
        return f;
    }
    void *map_staging(const StagingAllocation &allocation) override {
        return static_cast<HostBuffer *>(allocation.buffer.get())->bytes.data() + allocation.offset;
    }
    size_t get_aligned_uniform_size(size_t original_size) override { return (original_size + 255) & ~size_t(255); }

protected:
    std::shared_ptr<Buffer> create_staging_buffer(size_t size) override {
        return std::make_shared<HostBuffer>(BufferDescriptor{BufferType::Storage, size, MemoryProperty::HostVisibleAndCoherent});
    }
};

class HostQueue : public Queue {
public:
    void submit(const std::shared_ptr<CommandEncoder> &encoder, const std::shared_ptr<Fence> &) override {
        auto *e = static_cast<HostCommandEncoder *>(encoder.get());
        e->execute_writes();
        e->invoke_callbacks();
    }
};

} // namespace Pathfinder

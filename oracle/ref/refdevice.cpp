// refdevice.cpp — "interp": a LuisaCompute backend that executes kernels with the AST interpreter (interp.cpp).
// TEST INFRASTRUCTURE; see oracle/ref/README.md.
//
// It implements LuisaCompute's DeviceInterface (include/luisa/runtime/rhi/device_interface.h) the way the reference's own
// backends do (src/compute/src/backends/cuda/cuda_device.cpp is the model for which calls must work): buffers are host
// memory, shaders keep the recorded AST, a dispatch runs the AST once per dispatch id on the host's cores, bindless
// arrays / meshes / accels are small host tables.  Built as `liblc-backend-interp.so` next to the reference's own plugins
// and CLI (oracle/_ref/bin), it is selected like any other backend: `luisa-render-cli -b interp scene.luisa` runs the
// UNMODIFIED reference renderer end to end.
//
// Ray / triangle intersection is the one third-party piece of the reference's path (OptiX / Embree); here it
// is a brute-force loop over all triangles with the Moeller-Trumbore arithmetic this repository's oracle and CUDA kernels
// use (oracle/oracle.cpp: trace_brute), so that hits are comparable bit for bit.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>

#include <luisa/ast/function_builder.h>
#include <luisa/core/logging.h>
#include <luisa/runtime/context.h>
#include <luisa/runtime/device.h>
#include <luisa/runtime/rhi/command.h>
#include <luisa/runtime/rhi/device_interface.h>
#include <luisa/runtime/command_list.h>

#include "interp.h"

namespace luisa::compute::interp {

namespace {

struct BufferObject {
    luisa::vector<std::byte> data;
    size_t stride{0u};
};

struct BindlessObject {
    struct Slot {
        uint64_t buffer{0u};
        size_t offset{0u};
        uint64_t tex2d{0u};
        Sampler sampler{};
        uint64_t tex3d{0u};
    };
    luisa::vector<Slot> slots;
};

struct TextureObject {
    PixelStorage storage{PixelStorage::FLOAT4};
    uint32_t width{0u}, height{0u}, depth{1u};
    luisa::vector<std::byte> texels;// level 0, rows tightly packed (slices back to back for a volume)
};

float half_bits_to_float(uint16_t h) noexcept {
    auto sign = static_cast<uint32_t>(h & 0x8000u) << 16u;
    auto exponent = (h >> 10u) & 0x1fu;
    auto mantissa = static_cast<uint32_t>(h & 0x3ffu);
    uint32_t bits;
    if (exponent == 0u) {
        if (mantissa == 0u) {
            bits = sign;
        } else {// subnormal
            auto e = 127u - 15u + 1u;
            while ((mantissa & 0x400u) == 0u) { mantissa <<= 1u; e--; }
            bits = sign | (e << 23u) | ((mantissa & 0x3ffu) << 13u);
        }
    } else if (exponent == 31u) {
        bits = sign | 0x7f800000u | (mantissa << 13u);
    } else {
        bits = sign | ((exponent + 127u - 15u) << 23u) | (mantissa << 13u);
    }
    float f;
    std::memcpy(&f, &bits, 4u);
    return f;
}

struct MeshObject {
    uint64_t vertex_buffer{0u};
    size_t vertex_offset{0u}, vertex_size{0u}, vertex_stride{0u};
    uint64_t triangle_buffer{0u};
    size_t triangle_offset{0u}, triangle_size{0u};
    // A plain median-split bounding-volume hierarchy over the triangles (object space), only to skip triangles the ray
    // cannot reach: every triangle whose box the ray enters is still tested with the expressions below, so the set of
    // crossings - and the closest one - is the brute-force loop's (up to exact ties in t, which keep the first found).
    struct Node {
        float lo[3], hi[3];
        uint32_t left, right;// inner: child indices; leaf: left = first, right = count | 0x80000000
    };
    luisa::vector<Node> nodes;
    luisa::vector<uint32_t> order;// triangle indices in leaf order
};

struct AccelObject {
    struct Instance {
        float affine[12]{1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};// row-major 3x4, object -> world
        float inverse[12]{1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
        uint64_t mesh{0u};
        uint32_t visibility{0xffu};
        bool opaque{true};
    };
    luisa::vector<Instance> instances;
};

struct ShaderObject {
    luisa::shared_ptr<const detail::FunctionBuilder> builder;
    luisa::vector<Argument> bound;
};

struct StreamObject {};
struct EventObject {};

template<typename T>
[[nodiscard]] T *object(uint64_t handle) noexcept { return reinterpret_cast<T *>(handle); }

void invert_affine(const float m[12], float out[12]) noexcept {// rows of the 3x4 matrix [A | t]; out = [A^-1 | -A^-1 t]
    auto a = [&](int r, int c) { return static_cast<double>(m[r * 4 + c]); };
    auto det = a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) -
               a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
               a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    auto inv_det = 1.0 / det;
    double inv[3][3];
    inv[0][0] = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) * inv_det;
    inv[0][1] = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * inv_det;
    inv[0][2] = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * inv_det;
    inv[1][0] = (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) * inv_det;
    inv[1][1] = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * inv_det;
    inv[1][2] = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * inv_det;
    inv[2][0] = (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) * inv_det;
    inv[2][1] = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * inv_det;
    inv[2][2] = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * inv_det;
    for (auto r = 0; r < 3; r++) {
        for (auto c = 0; c < 3; c++) { out[r * 4 + c] = static_cast<float>(inv[r][c]); }
        out[r * 4 + 3] = static_cast<float>(-(inv[r][0] * a(0, 3) + inv[r][1] * a(1, 3) + inv[r][2] * a(2, 3)));
    }
}

void build_mesh_bvh(MeshObject *m) {
    auto vbuf = object<BufferObject>(m->vertex_buffer)->data.data() + m->vertex_offset;
    auto tbuf = reinterpret_cast<const uint32_t *>(object<BufferObject>(m->triangle_buffer)->data.data() + m->triangle_offset);
    auto count = static_cast<uint32_t>(m->triangle_size / 12u);
    auto position = [&](uint32_t v) { return reinterpret_cast<const float *>(vbuf + v * m->vertex_stride); };
    m->order.resize(count);
    for (auto i = 0u; i < count; i++) { m->order[i] = i; }
    luisa::vector<float> centroid(static_cast<size_t>(count) * 3u), tlo(static_cast<size_t>(count) * 3u), thi(static_cast<size_t>(count) * 3u);
    for (auto i = 0u; i < count; i++) {
        for (auto a = 0; a < 3; a++) {
            auto x = position(tbuf[3u * i])[a], y = position(tbuf[3u * i + 1u])[a], z = position(tbuf[3u * i + 2u])[a];
            tlo[3u * i + a] = std::min({x, y, z});
            thi[3u * i + a] = std::max({x, y, z});
            centroid[3u * i + a] = (x + y + z) * (1.f / 3.f);
        }
    }
    m->nodes.clear();
    m->nodes.reserve(2u * count / 4u + 8u);
    struct Task { uint32_t node, first, count; };
    luisa::vector<Task> stack;
    m->nodes.emplace_back();
    stack.push_back({0u, 0u, count});
    while (!stack.empty()) {
        auto task = stack.back();
        stack.pop_back();
        MeshObject::Node node{};
        float clo[3], chi[3];
        for (auto a = 0; a < 3; a++) { node.lo[a] = clo[a] = std::numeric_limits<float>::max(); node.hi[a] = chi[a] = -std::numeric_limits<float>::max(); }
        for (auto j = task.first; j < task.first + task.count; j++) {
            auto tri = m->order[j];
            for (auto a = 0; a < 3; a++) {
                node.lo[a] = std::min(node.lo[a], tlo[3u * tri + a]);
                node.hi[a] = std::max(node.hi[a], thi[3u * tri + a]);
                clo[a] = std::min(clo[a], centroid[3u * tri + a]);
                chi[a] = std::max(chi[a], centroid[3u * tri + a]);
            }
        }
        for (auto a = 0; a < 3; a++) {// pad: the box must never reject a triangle the exact test would accept
            auto pad = 1e-4f * std::max({std::fabs(node.lo[a]), std::fabs(node.hi[a]), 1e-3f});
            node.lo[a] -= pad;
            node.hi[a] += pad;
        }
        auto axis = 0;
        for (auto a = 1; a < 3; a++) { if (chi[a] - clo[a] > chi[axis] - clo[axis]) { axis = a; } }
        if (task.count <= 4u || !(chi[axis] > clo[axis])) {
            node.left = task.first;
            node.right = task.count | 0x80000000u;
        } else {
            auto mid = task.first + task.count / 2u;
            std::nth_element(m->order.begin() + task.first, m->order.begin() + mid, m->order.begin() + task.first + task.count,
                             [&](uint32_t x, uint32_t y) { return centroid[3u * x + axis] < centroid[3u * y + axis]; });
            node.left = static_cast<uint32_t>(m->nodes.size());
            node.right = node.left + 1u;
            m->nodes.emplace_back();
            m->nodes.emplace_back();
            stack.push_back({node.left, task.first, mid - task.first});
            stack.push_back({node.right, mid, task.first + task.count - mid});
        }
        m->nodes[task.node] = node;
    }
}

}// namespace

class InterpDevice final : public DeviceInterface, public refinterp::DeviceResources {

private:
    unsigned _threads;

public:
    explicit InterpDevice(Context &&ctx) noexcept
        : DeviceInterface{std::move(ctx)}, _threads{std::max(1u, std::thread::hardware_concurrency())} {
        if (auto env = std::getenv("LUISA_INTERP_THREADS")) { _threads = std::max(1, std::atoi(env)); }
    }

    void *native_handle() const noexcept override { return const_cast<InterpDevice *>(this); }
    uint compute_warp_size() const noexcept override { return 1u; }

    // ---- buffers
    BufferCreationInfo create_buffer(const Type *element, size_t elem_count, void *) noexcept override {
        auto stride = element == Type::of<void>() ? size_t{1u} : (element->size() + element->alignment() - 1u) / element->alignment() * element->alignment();
        auto b = new BufferObject;
        b->stride = stride;
        b->data.resize(std::max<size_t>(stride * elem_count, 16u), std::byte{0});
        BufferCreationInfo info{};
        info.handle = reinterpret_cast<uint64_t>(b);
        info.native_handle = b->data.data();
        info.element_stride = stride;
        info.total_size_bytes = stride * elem_count;
        return info;
    }
    BufferCreationInfo create_buffer(const ir::CArc<ir::Type> *, size_t, void *) noexcept override {
        LUISA_ERROR_WITH_LOCATION("interp: IR buffers are not supported.");
    }
    void destroy_buffer(uint64_t handle) noexcept override { delete object<BufferObject>(handle); }

    // ---- textures: not needed by the paths exercised (constant textures live in a buffer, the film is a buffer)
    // ---- textures: 2D, level 0 (the image-texture plugin samples without LOD, src/textures/image.cpp:166)
    ResourceCreationInfo create_texture(PixelFormat format, uint dimension, uint width, uint height, uint depth, uint, bool, bool) noexcept override {
        // 2D images, and 3D volumes for point reads (the blue-noise volume of the PMJ02BN sampler, pmj02bn.cpp:91-101)
        LUISA_ASSERT(dimension == 2u || dimension == 3u, "interp: only 2D / 3D textures are implemented.");
        auto t = new TextureObject;
        t->storage = pixel_format_to_storage(format);
        t->width = width;
        t->height = height;
        t->depth = dimension == 3u ? depth : 1u;
        t->texels.resize(pixel_storage_size(t->storage, make_uint3(width, height, t->depth)), std::byte{0});
        return {reinterpret_cast<uint64_t>(t), t};
    }
    void destroy_texture(uint64_t handle) noexcept override { delete object<TextureObject>(handle); }

    // ---- bindless arrays
    ResourceCreationInfo create_bindless_array(size_t size) noexcept override {
        auto a = new BindlessObject;
        a->slots.resize(size);
        return {reinterpret_cast<uint64_t>(a), a};
    }
    void destroy_bindless_array(uint64_t handle) noexcept override { delete object<BindlessObject>(handle); }

    // ---- streams / events: everything executes synchronously inside dispatch()
    ResourceCreationInfo create_stream(StreamTag) noexcept override {
        auto s = new StreamObject;
        return {reinterpret_cast<uint64_t>(s), s};
    }
    void destroy_stream(uint64_t handle) noexcept override { delete object<StreamObject>(handle); }
    void synchronize_stream(uint64_t) noexcept override {}
    void set_stream_log_callback(uint64_t, const StreamLogCallback &) noexcept override {}
    ResourceCreationInfo create_event() noexcept override {
        auto e = new EventObject;
        return {reinterpret_cast<uint64_t>(e), e};
    }
    void destroy_event(uint64_t handle) noexcept override { delete object<EventObject>(handle); }
    void signal_event(uint64_t, uint64_t, uint64_t) noexcept override {}
    void wait_event(uint64_t, uint64_t, uint64_t) noexcept override {}
    bool is_event_completed(uint64_t, uint64_t) const noexcept override { return true; }
    void synchronize_event(uint64_t, uint64_t) noexcept override {}

    // ---- swapchains: none
    SwapchainCreationInfo create_swapchain(uint64_t, uint64_t, uint, uint, bool, bool, uint) noexcept override {
        LUISA_ERROR_WITH_LOCATION("interp: no swapchains.");
    }
    void destroy_swap_chain(uint64_t) noexcept override {}
    void present_display_in_stream(uint64_t, uint64_t, uint64_t) noexcept override {}

    // ---- shaders
    ShaderCreationInfo create_shader(const ShaderOption &, Function kernel) noexcept override {
        auto s = new ShaderObject;
        s->builder = kernel.shared_builder();
        for (auto &&binding : kernel.bound_arguments()) {// as cuda_device.cpp:619-646
            luisa::visit(
                [s]<typename T>(T b) noexcept {
                    Argument a{};
                    if constexpr (std::is_same_v<T, Function::BufferBinding>) {
                        a.tag = Argument::Tag::BUFFER;
                        a.buffer = {b.handle, b.offset, b.size};
                    } else if constexpr (std::is_same_v<T, Function::BindlessArrayBinding>) {
                        a.tag = Argument::Tag::BINDLESS_ARRAY;
                        a.bindless_array.handle = b.handle;
                    } else if constexpr (std::is_same_v<T, Function::AccelBinding>) {
                        a.tag = Argument::Tag::ACCEL;
                        a.accel.handle = b.handle;
                    } else if constexpr (std::is_same_v<T, Function::TextureBinding>) {
                        LUISA_ERROR_WITH_LOCATION("interp: texture bindings are not implemented.");
                    } else {
                        LUISA_ERROR_WITH_LOCATION("interp: unbound captured argument.");
                    }
                    s->bound.emplace_back(a);
                },
                binding);
        }
        ShaderCreationInfo info{};
        info.handle = reinterpret_cast<uint64_t>(s);
        info.native_handle = s;
        info.block_size = kernel.block_size();
        return info;
    }
    ShaderCreationInfo create_shader(const ShaderOption &, const ir::KernelModule *) noexcept override {
        LUISA_ERROR_WITH_LOCATION("interp: IR kernels are not supported.");
    }
    ShaderCreationInfo load_shader(luisa::string_view, luisa::span<const Type *const>) noexcept override {
        LUISA_ERROR_WITH_LOCATION("interp: AOT shaders are not supported.");
    }
    Usage shader_argument_usage(uint64_t handle, size_t index) noexcept override {
        Function f{object<ShaderObject>(handle)->builder.get()};
        return f.variable_usage(f.arguments()[index].uid());
    }
    void destroy_shader(uint64_t handle) noexcept override { delete object<ShaderObject>(handle); }

    // ---- ray tracing resources
    ResourceCreationInfo create_mesh(const AccelOption &) noexcept override {
        auto m = new MeshObject;
        return {reinterpret_cast<uint64_t>(m), m};
    }
    void destroy_mesh(uint64_t handle) noexcept override { delete object<MeshObject>(handle); }
    ResourceCreationInfo create_procedural_primitive(const AccelOption &) noexcept override {
        LUISA_ERROR_WITH_LOCATION("interp: procedural primitives are not implemented.");
    }
    void destroy_procedural_primitive(uint64_t) noexcept override {}
    ResourceCreationInfo create_accel(const AccelOption &) noexcept override {
        auto a = new AccelObject;
        return {reinterpret_cast<uint64_t>(a), a};
    }
    void destroy_accel(uint64_t handle) noexcept override { delete object<AccelObject>(handle); }
    void set_name(luisa::compute::Resource::Tag, uint64_t, luisa::string_view) noexcept override {}

    // ---- command execution
    void dispatch(uint64_t, CommandList &&list) noexcept override {
        struct Visitor final : CommandVisitor {
            InterpDevice *device;
            explicit Visitor(InterpDevice *d) noexcept : device{d} {}
            void visit(const BufferUploadCommand *c) noexcept override {
                auto b = object<BufferObject>(c->handle());
                std::memcpy(b->data.data() + c->offset(), c->data(), c->size());
            }
            void visit(const BufferDownloadCommand *c) noexcept override {
                auto b = object<BufferObject>(c->handle());
                std::memcpy(c->data(), b->data.data() + c->offset(), c->size());
            }
            void visit(const BufferCopyCommand *c) noexcept override {
                std::memmove(object<BufferObject>(c->dst_handle())->data.data() + c->dst_offset(),
                             object<BufferObject>(c->src_handle())->data.data() + c->src_offset(), c->size());
            }
            void visit(const BufferToTextureCopyCommand *) noexcept override { LUISA_ERROR_WITH_LOCATION("interp: textures."); }
            void visit(const TextureUploadCommand *c) noexcept override {
                auto t = object<TextureObject>(c->handle());
                if (c->level() != 0u) { return; }// mip levels are never sampled here
                auto size = c->size();
                LUISA_ASSERT(all(c->offset() == make_uint3(0u)) && size.x == t->width && size.y == t->height && size.z == t->depth,
                             "interp: partial texture uploads.");
                std::memcpy(t->texels.data(), c->data(), t->texels.size());
            }
            void visit(const TextureDownloadCommand *) noexcept override { LUISA_ERROR_WITH_LOCATION("interp: textures."); }
            void visit(const TextureCopyCommand *) noexcept override { LUISA_ERROR_WITH_LOCATION("interp: textures."); }
            void visit(const TextureToBufferCopyCommand *) noexcept override { LUISA_ERROR_WITH_LOCATION("interp: textures."); }
            void visit(const CurveBuildCommand *) noexcept override { LUISA_ERROR_WITH_LOCATION("interp: curves."); }
            void visit(const ProceduralPrimitiveBuildCommand *) noexcept override { LUISA_ERROR_WITH_LOCATION("interp: procedural primitives."); }
            void visit(const CustomCommand *) noexcept override { LUISA_ERROR_WITH_LOCATION("interp: custom commands."); }
            void visit(const MeshBuildCommand *c) noexcept override {
                auto m = object<MeshObject>(c->handle());
                m->vertex_buffer = c->vertex_buffer();
                m->vertex_offset = c->vertex_buffer_offset();
                m->vertex_size = c->vertex_buffer_size();
                m->vertex_stride = c->vertex_stride();
                m->triangle_buffer = c->triangle_buffer();
                m->triangle_offset = c->triangle_buffer_offset();
                m->triangle_size = c->triangle_buffer_size();
                build_mesh_bvh(m);
            }
            void visit(const AccelBuildCommand *c) noexcept override {
                auto a = object<AccelObject>(c->handle());
                a->instances.resize(c->instance_count());
                using M = AccelBuildCommand::Modification;
                for (auto &&m : c->modifications()) {
                    auto &inst = a->instances[m.index];
                    if (m.flags & M::flag_primitive) { inst.mesh = m.primitive; }
                    if (m.flags & M::flag_transform) {
                        std::memcpy(inst.affine, m.affine, sizeof(inst.affine));
                        invert_affine(inst.affine, inst.inverse);
                    }
                    if (m.flags & M::flag_visibility) { inst.visibility = m.vis_mask; }
                    if (m.flags & M::flag_opaque_on) { inst.opaque = true; }
                    if (m.flags & M::flag_opaque_off) { inst.opaque = false; }
                }
            }
            void visit(const BindlessArrayUpdateCommand *c) noexcept override {
                auto a = object<BindlessObject>(c->handle());
                using Op = BindlessArrayUpdateCommand::Modification::Operation;
                for (auto &&m : const_cast<BindlessArrayUpdateCommand *>(c)->steal_modifications()) {
                    if (m.buffer.op == Op::EMPLACE) { a->slots[m.slot] = {m.buffer.handle, m.buffer.offset_bytes}; }
                    else if (m.buffer.op == Op::REMOVE) { a->slots[m.slot] = {}; }
                    if (m.tex2d.op == Op::EMPLACE) {
                        a->slots[m.slot].tex2d = m.tex2d.handle;
                        a->slots[m.slot].sampler = m.tex2d.sampler;
                    } else if (m.tex2d.op == Op::REMOVE) {
                        a->slots[m.slot].tex2d = 0u;
                    }
                    if (m.tex3d.op == Op::EMPLACE) { a->slots[m.slot].tex3d = m.tex3d.handle; }
                    else if (m.tex3d.op == Op::REMOVE) { a->slots[m.slot].tex3d = 0u; }
                }
            }
            void visit(const ShaderDispatchCommand *c) noexcept override { device->run(c); }
        } visitor{this};
        for (auto &&cmd : list.commands()) { cmd->accept(visitor); }
        for (auto &&callback : list.callbacks()) { callback(); }
    }

    void run(const ShaderDispatchCommand *c) noexcept {
        auto shader = object<ShaderObject>(c->handle());
        Function kernel{shader->builder.get()};
        if (c->is_indirect() || c->is_multiple_dispatch()) { LUISA_ERROR_WITH_LOCATION("interp: indirect / batched dispatch."); }
        luisa::vector<Argument> all{shader->bound};
        for (auto &&a : c->arguments()) { all.emplace_back(a); }
        auto params = kernel.arguments();
        LUISA_ASSERT(params.size() == all.size(), "interp: kernel takes {} arguments, {} given.", params.size(), all.size());
        std::vector<refinterp::Arg> args(all.size());
        for (auto i = 0u; i < all.size(); i++) {
            auto &&a = all[i];
            switch (a.tag) {
                case Argument::Tag::BUFFER: {
                    auto b = object<BufferObject>(a.buffer.handle);
                    args[i].kind = refinterp::Arg::Kind::BUFFER;
                    args[i].buffer = {b->data.data() + a.buffer.offset, a.buffer.size};
                    break;
                }
                case Argument::Tag::UNIFORM: {
                    auto u = c->uniform(a.uniform);
                    args[i].bytes.assign(u.begin(), u.end());
                    args[i].bytes.resize(params[i].type()->size());
                    break;
                }
                case Argument::Tag::BINDLESS_ARRAY:
                    args[i].kind = refinterp::Arg::Kind::BINDLESS_ARRAY;
                    args[i].handle = a.bindless_array.handle;
                    break;
                case Argument::Tag::ACCEL:
                    args[i].kind = refinterp::Arg::Kind::ACCEL;
                    args[i].handle = a.accel.handle;
                    break;
                default: LUISA_ERROR_WITH_LOCATION("interp: texture arguments are not implemented.");
            }
        }
        auto size = c->dispatch_size();
        uint32_t s[3] = {size.x, size.y, size.z};
        try {
            refinterp::launch(kernel, args, s, this, _threads);
        } catch (const std::exception &e) {
            LUISA_ERROR_WITH_LOCATION("interp: kernel failed: {}", e.what());
        }
    }

    // ---- refinterp::DeviceResources
    refinterp::BufferArg bindless_buffer(uint64_t array, uint32_t slot) override {
        auto a = object<BindlessObject>(array);
        // reads through an unpopulated slot happen in never-taken `ite` operands (both sides are evaluated, as on a GPU, where
        // such a read returns garbage that is then discarded): give them zeros
        if (slot >= a->slots.size() || a->slots[slot].buffer == 0u) { return {nullptr, 0u}; }
        auto b = object<BufferObject>(a->slots[slot].buffer);
        return {b->data.data() + a->slots[slot].offset, b->data.size() - a->slots[slot].offset};
    }

    // The reference's software sampler: src/compute/src/rust/luisa_compute_backend_impl/src/cpu/codegen/cpu_texture.h
    // (:63 unorm conversion, :369-372 out-of-bounds reads return zero, :418-464 coordinates + bilinear, :489-493 point)
    static void read_texel(const TextureObject *t, uint32_t x, uint32_t y, float out[4], uint32_t z = 0u) {
        out[0] = out[1] = out[2] = out[3] = 0.f;
        if (!(x < t->width && y < t->height && z < t->depth)) { return; }
        auto channels = pixel_storage_channel_count(t->storage);
        auto index = (static_cast<size_t>(z) * t->height + y) * t->width + x;
        for (auto c = 0u; c < channels; c++) {
            switch (t->storage) {
                case PixelStorage::BYTE1:
                case PixelStorage::BYTE2:
                case PixelStorage::BYTE4: out[c] = static_cast<float>(reinterpret_cast<const uint8_t *>(t->texels.data())[index * channels + c]) / 255.f; break;
                case PixelStorage::SHORT1:
                case PixelStorage::SHORT2:
                case PixelStorage::SHORT4: out[c] = static_cast<float>(reinterpret_cast<const uint16_t *>(t->texels.data())[index * channels + c]) / 65535.f; break;
                case PixelStorage::HALF1:
                case PixelStorage::HALF2:
                case PixelStorage::HALF4: out[c] = half_bits_to_float(reinterpret_cast<const uint16_t *>(t->texels.data())[index * channels + c]); break;
                case PixelStorage::FLOAT1:
                case PixelStorage::FLOAT2:
                case PixelStorage::FLOAT4: out[c] = reinterpret_cast<const float *>(t->texels.data())[index * channels + c]; break;
                default: throw std::runtime_error("unsupported texture storage");
            }
        }
    }
    static float coord(Sampler::Address address, float uv, float s) {
        constexpr auto one_minus_epsilon = 0x1.fffffep-1f;
        switch (address) {
            case Sampler::Address::EDGE: return std::fmin(std::fmax(uv, 0.0f), one_minus_epsilon) * s;
            case Sampler::Address::REPEAT: return (uv - std::floor(uv)) * s;
            case Sampler::Address::MIRROR: {
                uv = std::fmod(std::fabs(uv), 2.0f);
                uv = uv < 1.f ? uv : 2.f - uv;
                return std::fmin(uv, one_minus_epsilon) * s;
            }
            default: return (uv < 0.f || uv >= 1.f) ? 65536.f : uv * s;
        }
    }
    const BindlessObject::Slot &tex_slot(uint64_t array, uint32_t slot) const {
        auto a = object<BindlessObject>(array);
        if (slot >= a->slots.size() || a->slots[slot].tex2d == 0u) { throw std::runtime_error("empty bindless texture slot"); }
        return a->slots[slot];
    }
    void bindless_tex2d_sample(uint64_t array, uint32_t slot, float u, float v, float out[4]) override {
        auto &s = tex_slot(array, slot);
        auto t = object<TextureObject>(s.tex2d);
        auto sx = static_cast<float>(t->width), sy = static_cast<float>(t->height);
        auto address = s.sampler.address();
        if (s.sampler.filter() == Sampler::Filter::POINT) {
            read_texel(t, static_cast<uint32_t>(coord(address, u, sx)), static_cast<uint32_t>(coord(address, v, sy)), out);
            return;
        }
        auto inv_sx = 1.f / sx, inv_sy = 1.f / sy;
        auto ax = coord(address, u - .5f * inv_sx, sx), bx = coord(address, u + .5f * inv_sx, sx);
        auto ay = coord(address, v - .5f * inv_sy, sy), by = coord(address, v + .5f * inv_sy, sy);
        auto x_min = std::fmin(ax, bx), x_max = std::fmax(ax, bx), y_min = std::fmin(ay, by), y_max = std::fmax(ay, by);
        auto tx = x_max - std::floor(x_max), ty = y_max - std::floor(y_max);
        auto x0 = static_cast<uint32_t>(x_min), y0 = static_cast<uint32_t>(y_min), x1 = static_cast<uint32_t>(x_max), y1 = static_cast<uint32_t>(y_max);
        float v00[4], v01[4], v10[4], v11[4];
        read_texel(t, x0, y0, v00);
        read_texel(t, x1, y0, v01);
        read_texel(t, x0, y1, v10);
        read_texel(t, x1, y1, v11);
        for (auto c = 0; c < 4; c++) {
            auto a = tx * (v01[c] - v00[c]) + v00[c];
            auto b = tx * (v11[c] - v10[c]) + v10[c];
            out[c] = ty * (b - a) + a;
        }
    }
    void bindless_tex2d_size(uint64_t array, uint32_t slot, uint32_t out[2]) override {
        auto t = object<TextureObject>(tex_slot(array, slot).tex2d);
        out[0] = t->width;
        out[1] = t->height;
    }
    void bindless_tex2d_read(uint64_t array, uint32_t slot, uint32_t x, uint32_t y, float out[4]) override {
        read_texel(object<TextureObject>(tex_slot(array, slot).tex2d), x, y, out);
    }
    void bindless_tex3d_read(uint64_t array, uint32_t slot, uint32_t x, uint32_t y, uint32_t z, float out[4]) override {
        auto a = object<BindlessObject>(array);
        if (slot >= a->slots.size() || a->slots[slot].tex3d == 0u) { throw std::runtime_error("empty bindless 3D texture slot"); }
        read_texel(object<TextureObject>(a->slots[slot].tex3d), x, y, out, z);
    }

    std::vector<Candidate> candidates(uint64_t accel, const refinterp::RayData &ray, uint32_t mask) override {
        std::vector<Candidate> out;
        trace<false>(accel, ray, mask, &out);
        return out;
    }

    template<bool any_hit>
    refinterp::HitData trace(uint64_t accel, const refinterp::RayData &ray, uint32_t mask, std::vector<Candidate> *all = nullptr) const {
        // oracle/oracle.cpp: trace_brute — the same expressions, explicit fma where the oracle / CUDA kernels spell it out
        auto a = object<AccelObject>(accel);
        refinterp::HitData best{~0u, ~0u, {0.f, 0.f}, ray.t_max, 0u};
        auto tbest = ray.t_max;
        auto fdot = [](const float x[3], const float y[3]) { return std::fmaf(x[0], y[0], std::fmaf(x[1], y[1], x[2] * y[2])); };
        auto fcross = [](const float x[3], const float y[3], float out[3]) {
            out[0] = std::fmaf(x[1], y[2], -(x[2] * y[1]));
            out[1] = std::fmaf(x[2], y[0], -(x[0] * y[2]));
            out[2] = std::fmaf(x[0], y[1], -(x[1] * y[0]));
        };
        for (auto i = 0u; i < a->instances.size(); i++) {
            auto &inst = a->instances[i];
            if ((inst.visibility & mask) == 0u || inst.mesh == 0u) { continue; }
            auto w = inst.inverse;
            auto o = ray.origin, d = ray.direction;
            float oo[3] = {std::fmaf(w[0], o[0], std::fmaf(w[1], o[1], std::fmaf(w[2], o[2], w[3]))),
                           std::fmaf(w[4], o[0], std::fmaf(w[5], o[1], std::fmaf(w[6], o[2], w[7]))),
                           std::fmaf(w[8], o[0], std::fmaf(w[9], o[1], std::fmaf(w[10], o[2], w[11])))};
            float dd[3] = {std::fmaf(w[0], d[0], std::fmaf(w[1], d[1], w[2] * d[2])),
                           std::fmaf(w[4], d[0], std::fmaf(w[5], d[1], w[6] * d[2])),
                           std::fmaf(w[8], d[0], std::fmaf(w[9], d[1], w[10] * d[2]))};
            auto mesh = object<MeshObject>(inst.mesh);
            auto vbuf = object<BufferObject>(mesh->vertex_buffer)->data.data() + mesh->vertex_offset;
            auto tbuf = reinterpret_cast<const uint32_t *>(object<BufferObject>(mesh->triangle_buffer)->data.data() + mesh->triangle_offset);
            auto tri_count = mesh->triangle_size / 12u;
            auto position = [&](uint32_t v) { return reinterpret_cast<const float *>(vbuf + v * mesh->vertex_stride); };
            // object-space slab test (conservative: padded boxes, NaN-dropping min / max)
            float inv[3] = {1.f / dd[0], 1.f / dd[1], 1.f / dd[2]};
            auto enters = [&](const MeshObject::Node &n, float t_far) {
                auto t0 = ray.t_min, t1 = t_far;
                for (auto ax = 0; ax < 3; ax++) {
                    auto a0 = (n.lo[ax] - oo[ax]) * inv[ax], a1 = (n.hi[ax] - oo[ax]) * inv[ax];
                    if (dd[ax] == 0.f) {// parallel: inside the slab or not at all
                        if (oo[ax] < n.lo[ax] || oo[ax] > n.hi[ax]) { return false; }
                        continue;
                    }
                    t0 = std::fmax(t0, std::fmin(a0, a1));
                    t1 = std::fmin(t1, std::fmax(a0, a1));
                }
                return t0 <= t1 * 1.0001f + 1e-6f;
            };
            auto test = [&](uint32_t k) -> bool {// returns true when the traversal can stop (any-hit)
                auto p0 = position(tbuf[3u * k]), p1 = position(tbuf[3u * k + 1u]), p2 = position(tbuf[3u * k + 2u]);
                float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
                float e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
                float pvec[3], qvec[3];
                fcross(dd, e2, pvec);
                auto det = fdot(e1, pvec);
                if (!(det != 0.0f)) { return false; }
                auto inv_det = 1.0f / det;
                float tvec[3] = {oo[0] - p0[0], oo[1] - p0[1], oo[2] - p0[2]};
                auto u = fdot(tvec, pvec) * inv_det;
                if (!(u >= 0.0f && u <= 1.0f)) { return false; }
                fcross(tvec, e1, qvec);
                auto v = fdot(dd, qvec) * inv_det;
                if (!(v >= 0.0f && u + v <= 1.0f)) { return false; }
                auto t = fdot(e2, qvec) * inv_det;
                // exact ties in t: the lower (instance, primitive) wins, as in an index-ordered brute-force loop
                auto tie = t == tbest && best.inst != ~0u && (i < best.inst || (i == best.inst && k < best.prim));
                if (!(t > ray.t_min && (t < tbest || tie))) { return false; }
                if (all != nullptr) {// ray query: every crossing is a candidate, the query decides what is committed
                    all->push_back({refinterp::HitData{i, k, {u, v}, t, 0u}, inst.opaque});
                    return false;
                }
                tbest = t;
                best = {i, k, {u, v}, t, 0u};
                return any_hit;
            };
            (void)tri_count;
            uint32_t todo[128];
            auto top = 0u;
            todo[top++] = 0u;
            auto done = false;
            while (top != 0u && !done) {
                auto &n = mesh->nodes[todo[--top]];
                if (!enters(n, tbest)) { continue; }
                if (n.right & 0x80000000u) {
                    for (auto j = n.left; j < n.left + (n.right & 0x7fffffffu) && !done; j++) { done = test(mesh->order[j]); }
                } else {
                    if (top + 2u > 128u) { throw std::runtime_error("BVH stack overflow"); }
                    todo[top++] = n.left;
                    todo[top++] = n.right;
                }
            }
            if (done) { return best; }
        }
        return best;
    }
    refinterp::HitData trace_closest(uint64_t accel, const refinterp::RayData &ray, uint32_t mask) override { return trace<false>(accel, ray, mask); }
    bool trace_any(uint64_t accel, const refinterp::RayData &ray, uint32_t mask) override { return trace<true>(accel, ray, mask).inst != ~0u; }
    void instance_transform(uint64_t accel, uint32_t index, float out[16]) override {
        auto &m = object<AccelObject>(accel)->instances.at(index).affine;
        for (auto c = 0; c < 4; c++) {
            for (auto r = 0; r < 3; r++) { out[c * 4 + r] = m[r * 4 + c]; }
            out[c * 4 + 3] = c == 3 ? 1.f : 0.f;
        }
    }
};

}// namespace luisa::compute::interp

LUISA_EXPORT_API luisa::compute::DeviceInterface *create(luisa::compute::Context &&ctx, const luisa::compute::DeviceConfig *) noexcept {
    return luisa::new_with_allocator<luisa::compute::interp::InterpDevice>(std::move(ctx));
}

LUISA_EXPORT_API void destroy(luisa::compute::DeviceInterface *device) noexcept {
    luisa::delete_with_allocator(device);
}

LUISA_EXPORT_API void backend_device_names(luisa::vector<luisa::string> &names) noexcept {
    names.clear();
    names.emplace_back("host AST interpreter");
}

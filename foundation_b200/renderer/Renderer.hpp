// Renderer.hpp — host side of the drop-in boundary, shaped like mos9527/Foundation's renderer:
//     Foundation::Renderer::Renderer(device, allocator); Draw(); ~Renderer()
// (reference: src/Renderer/Renderer.hpp:10-44, public API :40-43).  Where the reference's Draw() records one indexed draw
// (src/Renderer/Renderer.cpp:332-351) this one advances a progressive path-traced render by one sample batch.  It calls ONLY
// the C ABI in include/foundation_pt.h — no CUDA, no torch — so a Foundation maintainer can compile it inside src/Renderer as is.
//
// The reference's RHI / Core headers are not available here (Vulkan, mimalloc, glm are network FetchContent dependencies,
// SURVEY.md §8c), so the three things the class shape needs are declared in minimal form below with the reference's names and
// argument meaning: Core::Allocator (src/Core/Allocator/Allocator.hpp:17-41), a non-owning device handle
// (RHIApplicationObjectHandle<RHIDevice>, src/Renderer/Renderer.hpp:13 — here it only carries the CUDA ordinal), and CHECK
// (src/Core/Core.hpp:17).  In-tree, delete these stand-ins and include the real headers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/foundation_pt.h"
#include "Present.hpp"

namespace Foundation {
namespace Core {
class Allocator {  // same virtual surface as the reference's Core::Allocator
public:
    using size_type = std::size_t;
    using pointer = void*;
    virtual ~Allocator() = default;
    virtual pointer Allocate(size_type size) = 0;
    virtual pointer Allocate(size_type size, size_t alignment) = 0;
    virtual void Deallocate(pointer ptr, size_type size) = 0;
    virtual void Deallocate(pointer ptr) = 0;
    virtual pointer Reallocate(pointer ptr, size_type new_size, size_t alignment) = 0;
    virtual size_type GetUsedMemory() const noexcept = 0;
    inline Allocator* Ptr() { return this; }
};
[[noreturn]] void BugCheck(const char* what);  // print + terminate, like Core::BugCheck (src/Core/Core.cpp:7-10)
}  // namespace Core

#define FOUNDATION_CHECK(expr) do { if (!(expr)) ::Foundation::Core::BugCheck(#expr); } while (0)

namespace Renderer {

struct DeviceHandle { int32_t cuda_ordinal = 0; };  // stands in for RHIApplicationObjectHandle<RHIDevice> (non-owning)

// The scene description the north_star presumes and the reference lacks ("no scene type, no pass type").
// Data arrives in the reference's formats: float3 positions with a byte stride, R16/R32_UINT indices
// (src/Platform/RHI/Common.hpp:18-27), column-major matrices (src/Renderer/Renderer.cpp:28-33).
struct MeshDesc {
    std::vector<float> positions;        // tightly packed xyz
    std::vector<uint32_t> indices;       // 3 per triangle
    std::vector<uint32_t> material_ids;  // 1 per triangle
};
struct SceneDesc {
    std::vector<MeshDesc> meshes;
    std::vector<foundation_pt_material> materials;
    std::vector<foundation_pt_instance> instances;  // empty = every mesh once, identity
    float view[16], proj[16];                        // uniform_buffer.view / .proj
    uint32_t width = 1920, height = 1080;            // the reference's swapchain extent (Renderer.cpp:41)
    float background[3] = {0, 0, 0};
    bool Load(const char* path, std::string* error);  // "FPTS" file written by foundation_b200.scenes.save_scene
};

class Renderer : public FrameSource {   // FrameSource: what PresentUploader (Present.hpp, the RHI-side hand-off) pulls the RGBA8 frame through
    Core::Allocator* m_allocator{nullptr};
    DeviceHandle m_device;
    foundation_pt_group* m_group{nullptr};   // one member per device; the frame is gathered into member 0 inside foundation_pt_group_render
    foundation_pt_context* m_ctx{nullptr};   // member 0 (borrowed from the group)
    template <class F> void ForEachMember(F f);
    uint32_t m_width{0}, m_height{0};
    uint32_t m_samples_done{0};
    uint32_t m_samples_per_draw{1}, m_max_bounces{4};
    uint8_t* m_present_image{nullptr};   // R8G8B8A8_UNORM "swapchain image" (host), allocated through m_allocator
    void Check(int32_t status, const char* what) const;

public:
    Renderer(DeviceHandle device, Core::Allocator* allocator, const SceneDesc& scene, uint64_t seed = 1);
    // Several devices of one box (the reference takes EnumerateDevices()[0] only, src/Editor/Editor.cpp:18): the scene is replicated, the
    // frame is split into interleaved tiles and gathered over NVLink inside Draw().  direct_gather: the accumulate kernels store their
    // tiles straight into device 0's frame over peer memory (FOUNDATION_PT_COMM_DIRECT) instead of packed ncclSend / ncclRecv.
    Renderer(const std::vector<DeviceHandle>& devices, Core::Allocator* allocator, const SceneDesc& scene, uint64_t seed = 1, bool direct_gather = false);
    uint32_t DeviceCount() const { return foundation_pt_group_size(m_group); }
    ~Renderer();
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;
    void SetCamera(const float view[16], const float proj[16]);   // the reference rewrites the UBO every Draw (Renderer.cpp:372-380); resets accumulation
    void SetQuality(uint32_t samples_per_draw, uint32_t max_bounces);
    // Per-frame model matrices, the analogue of `ubo.model = glm::rotate(..., time * 90deg, +Z)` (Renderer.cpp:373): replaces the instance list,
    // rebuilds only the TLAS (every BLAS is kept) and restarts the accumulation.
    void SetInstances(const foundation_pt_instance* instances, uint32_t count);
    void Draw();                                                   // one sample batch + resolve to the present image (blocking, like Renderer.cpp:394)
    const uint8_t* PresentImage() const { return m_present_image; }
    // FrameSource
    uint32_t FrameWidth() const override { return m_width; }
    uint32_t FrameHeight() const override { return m_height; }
    void ResolveRGBA8(uint8_t* dst, size_t size_bytes) override;   // accum / spp -> R8G8B8A8_UNORM straight into the caller's (mapped staging) memory
    uint32_t Width() const { return m_width; }
    uint32_t Height() const { return m_height; }
    uint32_t SamplesDone() const { return m_samples_done; }
    void ReadAccum(std::vector<float>* rgba) const;
    foundation_pt_stats Stats() const;
    foundation_pt_build_stats BuildStats() const { return m_build; }

private:
    foundation_pt_build_stats m_build{};
};

}  // namespace Renderer
}  // namespace Foundation

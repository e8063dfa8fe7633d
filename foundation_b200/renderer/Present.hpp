// Present.hpp — SURVEY.md section 8f rank 1: hand the path-traced frame to Foundation's RHI as an R8G8B8A8_UNORM image, through the
// reference's own staging-upload idiom (mos9527/Foundation src/Renderer/Renderer.cpp:200-270: host-visible coherent buffer -> Map ->
// transition Undefined -> TransferDst -> CopyBufferToImage -> transition to ShaderReadOnly -> Submit -> WaitIdle).  The target is a sampled
// image: the reference's textured-quad pipeline (Renderer.cpp:332-351, Triangle.slang:34-37) then shows it — the swapchain images themselves
// are created with colour-attachment usage only (src/Platform/RHI/Vulkan/Swapchain.cpp:16), so they cannot be copied into directly.
// There is no CUDA <-> Vulkan external-memory route in the reference's RHI, so the frame makes one host round trip (resolve_rgba8).
#pragma once
#include <cstddef>
#include <cstdint>

#if __has_include(<Platform/RHI/Device.hpp>)
#include <Platform/RHI/Device.hpp>          // in-tree: the real abstract RHI
#else
#include "rhi_stub/RHI.hpp"                 // here: the subset Present.cpp uses, declared for the compile check
#endif

namespace Foundation {
namespace Renderer {

// what the uploader needs from the path tracer: implemented by Renderer (foundation_pt_resolve_rgba8) and, in the self-test, by a pattern generator
class FrameSource {
public:
    virtual ~FrameSource() = default;
    virtual uint32_t FrameWidth() const = 0;
    virtual uint32_t FrameHeight() const = 0;
    virtual void ResolveRGBA8(uint8_t* dst, size_t size_bytes) = 0;
};

class PresentUploader {
    Platform::RHI::RHIDevice* m_device;
    Platform::RHI::RHIDeviceQueue* m_queue;
    Platform::RHI::RHICommandPool* m_cmd_pool;
    Platform::RHI::RHIBuffer* m_staging{nullptr};
    Platform::RHI::RHIImage* m_image{nullptr};
    uint32_t m_width, m_height;
    bool m_first{true};

public:
    PresentUploader(Platform::RHI::RHIDevice* device, Platform::RHI::RHIDeviceQueue* queue, Platform::RHI::RHICommandPool* cmd_pool, uint32_t width, uint32_t height);
    Platform::RHI::RHIImage* Image() const { return m_image; }     // bind this where the reference binds m_tex_view (Renderer.cpp:296-306)
    void Upload(FrameSource& frame);                               // blocking, like every upload in the reference's constructor
};

}  // namespace Renderer
}  // namespace Foundation

// rhi_stub/RHI.hpp — the SUBSET of Foundation's abstract RHI that the present path touches, declared here only so that Present.cpp can be
// compiled and exercised without Vulkan / GLFW / mimalloc (none of which exist in this image).  Names, argument
// meaning and call protocol follow mos9527/Foundation src/Platform/RHI: Resource.hpp:9-26 (RHIResourceDesc, RHIBufferDesc), :42-53
// (Map / Unmap), :70-80 (RHIImageDesc); Command.hpp:39-51 (TransitionDesc, Begin/Set/EndTransition), :82-94 (CopyImageRegion,
// CopyBufferToImage), :112-113 (Begin / End); Device.hpp:21-29 (WaitIdle, SubmitDesc, Submit).  Deliberately simplified: plain pointers
// where the reference uses its Handle / ScopedHandle templates (Details.hpp), std::vector where it uses Core::StlSpan.  In-tree, delete this
// file and include <Platform/RHI/Device.hpp> — Present.cpp uses nothing beyond what is declared here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace Foundation {
namespace Platform {
namespace RHI {

enum class RHIResourceHostAccess { Invisible, ReadWrite };
enum class RHIBufferUsage : uint32_t { TransferSource = 1u << 0, TransferDestination = 1u << 1 };
enum class RHIImageUsage : uint32_t { SampledImage = 1u << 0, TransferDestination = 1u << 1 };
enum class RHIResourceFormat { R8G8B8A8_UNORM };
enum class RHIImageLayout { Undefined, TransferDst, ShaderReadOnly };
enum class RHIResourceAccess { Undefined, TransferWrite, ShaderRead };
enum class RHIPipelineStage { TopOfPipe, Transfer, FragmentShader };
struct RHIExtent3D { uint32_t x, y, z; };

struct RHIResourceDesc { std::string name; RHIResourceHostAccess host_access{RHIResourceHostAccess::Invisible}; bool shared{false}; bool coherent{false}; };
struct RHIBufferDesc { RHIResourceDesc resource; RHIBufferUsage usage; size_t size; };
struct RHIImageDesc { RHIResourceDesc resource; RHIImageUsage usage; RHIExtent3D extent{1, 1, 1}; RHIResourceFormat format; RHIImageLayout initial_layout{RHIImageLayout::Undefined}; };

class RHIBuffer { public: virtual ~RHIBuffer() = default; virtual void* Map() = 0; virtual void Unmap() = 0; };
class RHIImage { public: virtual ~RHIImage() = default; };

class RHICommandList {
public:
    struct TransitionDesc { RHIResourceAccess src_access, dst_access; RHIPipelineStage src_stage, dst_stage; RHIImageLayout src_img_layout, dst_img_layout; };
    struct CopyImageRegion { uint32_t src_buffer_offset = 0; RHIExtent3D extent{1, 1, 1}; };
    virtual ~RHICommandList() = default;
    virtual RHICommandList& Begin() = 0;
    virtual void End() = 0;
    virtual RHICommandList& BeginTransition() = 0;
    virtual RHICommandList& SetImageTransition(RHIImage* image, TransitionDesc const& desc) = 0;
    virtual RHICommandList& EndTransition() = 0;
    virtual RHICommandList& CopyBufferToImage(RHIBuffer* src_buffer, RHIImage* dst_image, RHIImageLayout dst_layout, std::vector<CopyImageRegion> const& regions) = 0;
};
class RHICommandPool { public: virtual ~RHICommandPool() = default; virtual RHICommandList* CreateCommandList() = 0; };
class RHIDeviceQueue {
public:
    struct SubmitDesc { std::vector<RHICommandList*> cmd_lists; };
    virtual ~RHIDeviceQueue() = default;
    virtual void Submit(SubmitDesc const& desc) const = 0;
    virtual void WaitIdle() const = 0;
};
class RHIDevice {
public:
    virtual ~RHIDevice() = default;
    virtual RHIBuffer* CreateBuffer(RHIBufferDesc const& desc) = 0;
    virtual RHIImage* CreateImage(RHIImageDesc const& desc) = 0;
};

}  // namespace RHI
}  // namespace Platform
}  // namespace Foundation

// Present.cpp — see Present.hpp.  Only abstract-RHI calls; compiles against the stub subset here and against the real headers in-tree.
#include "Present.hpp"

namespace Foundation {
namespace Renderer {
using namespace Platform::RHI;

PresentUploader::PresentUploader(RHIDevice* device, RHIDeviceQueue* queue, RHICommandPool* cmd_pool, uint32_t width, uint32_t height)
    : m_device(device), m_queue(queue), m_cmd_pool(cmd_pool), m_width(width), m_height(height) {
    RHIImageDesc img{};
    img.resource.name = "Path Traced Frame";
    img.resource.host_access = RHIResourceHostAccess::Invisible;
    img.usage = (RHIImageUsage)((uint32_t)RHIImageUsage::SampledImage | (uint32_t)RHIImageUsage::TransferDestination);
    img.extent = {width, height, 1};
    img.format = RHIResourceFormat::R8G8B8A8_UNORM;           // the reference's colour format (Renderer.cpp:40, :205)
    img.initial_layout = RHIImageLayout::Undefined;
    m_image = m_device->CreateImage(img);
    RHIBufferDesc buf{};                                      // one persistent staging buffer (the reference creates one per upload)
    buf.resource.host_access = RHIResourceHostAccess::ReadWrite;
    buf.resource.coherent = true;
    buf.usage = RHIBufferUsage::TransferSource;
    buf.size = (size_t)width * height * 4;
    m_staging = m_device->CreateBuffer(buf);
}

void PresentUploader::Upload(FrameSource& frame) {
    const size_t bytes = (size_t)m_width * m_height * 4;
    frame.ResolveRGBA8(static_cast<uint8_t*>(m_staging->Map()), bytes);      // foundation_pt_resolve_rgba8 writes straight into the mapped buffer
    m_staging->Unmap();
    RHICommandList* cmd = m_cmd_pool->CreateCommandList();
    cmd->Begin();
    cmd->BeginTransition();
    RHICommandList::TransitionDesc to_dst{};
    to_dst.src_access = m_first ? RHIResourceAccess::Undefined : RHIResourceAccess::ShaderRead;
    to_dst.dst_access = RHIResourceAccess::TransferWrite;
    to_dst.src_stage = m_first ? RHIPipelineStage::TopOfPipe : RHIPipelineStage::FragmentShader;
    to_dst.dst_stage = RHIPipelineStage::Transfer;
    to_dst.src_img_layout = m_first ? RHIImageLayout::Undefined : RHIImageLayout::ShaderReadOnly;
    to_dst.dst_img_layout = RHIImageLayout::TransferDst;
    cmd->SetImageTransition(m_image, to_dst);
    cmd->EndTransition();
    RHICommandList::CopyImageRegion region{};
    region.extent = {m_width, m_height, 1};
    cmd->CopyBufferToImage(m_staging, m_image, RHIImageLayout::TransferDst, {region});
    cmd->BeginTransition();
    RHICommandList::TransitionDesc to_read{};
    to_read.src_access = RHIResourceAccess::TransferWrite; to_read.dst_access = RHIResourceAccess::ShaderRead;
    to_read.src_stage = RHIPipelineStage::Transfer; to_read.dst_stage = RHIPipelineStage::FragmentShader;
    to_read.src_img_layout = RHIImageLayout::TransferDst; to_read.dst_img_layout = RHIImageLayout::ShaderReadOnly;
    cmd->SetImageTransition(m_image, to_read);
    cmd->EndTransition();
    cmd->End();
    RHIDeviceQueue::SubmitDesc submit{};
    submit.cmd_lists = {cmd};
    m_queue->Submit(submit);
    m_queue->WaitIdle();
    m_first = false;
}

}  // namespace Renderer
}  // namespace Foundation

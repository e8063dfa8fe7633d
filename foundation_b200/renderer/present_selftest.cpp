// present_selftest.cpp — runs PresentUploader against a recording mock of the abstract RHI (no GPU, no Vulkan): the call protocol must be
// the reference's staging-upload idiom (src/Renderer/Renderer.cpp:219-251) and the image must receive exactly the resolved bytes.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "Present.hpp"

using namespace Foundation::Platform::RHI;
using namespace Foundation::Renderer;

namespace {
std::string g_log;
struct MockBuffer final : RHIBuffer { std::vector<uint8_t> bytes; bool mapped = false; void* Map() override { mapped = true; g_log += "Map "; return bytes.data(); } void Unmap() override { mapped = false; g_log += "Unmap "; } };
struct MockImage final : RHIImage { std::vector<uint8_t> texels; RHIImageLayout layout = RHIImageLayout::Undefined; uint32_t w = 0, h = 0; };
struct MockCmd final : RHICommandList {
    bool recording = false, in_transition = false;
    std::vector<std::pair<MockBuffer*, MockImage*>> copies;
    RHICommandList& Begin() override { recording = true; g_log += "Begin "; return *this; }
    void End() override { recording = false; g_log += "End "; }
    RHICommandList& BeginTransition() override { in_transition = true; g_log += "BeginTransition "; return *this; }
    RHICommandList& SetImageTransition(RHIImage* image, TransitionDesc const& d) override {
        MockImage* im = static_cast<MockImage*>(image);
        if (!in_transition || im->layout != d.src_img_layout) g_log += "BAD-TRANSITION ";
        im->layout = d.dst_img_layout; g_log += "SetImageTransition "; return *this;
    }
    RHICommandList& EndTransition() override { in_transition = false; g_log += "EndTransition "; return *this; }
    RHICommandList& CopyBufferToImage(RHIBuffer* b, RHIImage* i, RHIImageLayout l, std::vector<CopyImageRegion> const& r) override {
        MockImage* im = static_cast<MockImage*>(i); MockBuffer* bf = static_cast<MockBuffer*>(b);
        if (l != RHIImageLayout::TransferDst || im->layout != RHIImageLayout::TransferDst || bf->mapped || r.size() != 1 || r[0].extent.x != im->w || r[0].extent.y != im->h) g_log += "BAD-COPY ";
        copies.push_back({bf, im}); g_log += "CopyBufferToImage "; return *this;
    }
};
struct MockPool final : RHICommandPool { std::vector<MockCmd*> lists; RHICommandList* CreateCommandList() override { lists.push_back(new MockCmd()); return lists.back(); } ~MockPool() { for (auto* l : lists) delete l; } };
struct MockQueue final : RHIDeviceQueue {
    mutable int submitted = 0, waited = 0;
    void Submit(SubmitDesc const& d) const override {
        for (RHICommandList* c : d.cmd_lists) { MockCmd* m = static_cast<MockCmd*>(c); if (m->recording) g_log += "BAD-SUBMIT "; for (auto& cp : m->copies) cp.second->texels = cp.first->bytes; }
        ++submitted; g_log += "Submit ";
    }
    void WaitIdle() const override { ++waited; g_log += "WaitIdle "; }
};
struct MockDevice final : RHIDevice {
    std::vector<MockBuffer*> buffers; std::vector<MockImage*> images; std::string problems;
    RHIBuffer* CreateBuffer(RHIBufferDesc const& d) override {
        if (d.resource.host_access != RHIResourceHostAccess::ReadWrite || !d.resource.coherent || d.usage != RHIBufferUsage::TransferSource) problems += "staging-desc ";
        buffers.push_back(new MockBuffer()); buffers.back()->bytes.resize(d.size); return buffers.back();
    }
    RHIImage* CreateImage(RHIImageDesc const& d) override {
        if (d.format != RHIResourceFormat::R8G8B8A8_UNORM || !((uint32_t)d.usage & (uint32_t)RHIImageUsage::TransferDestination) || !((uint32_t)d.usage & (uint32_t)RHIImageUsage::SampledImage)) problems += "image-desc ";
        images.push_back(new MockImage()); images.back()->w = d.extent.x; images.back()->h = d.extent.y; return images.back();
    }
    ~MockDevice() { for (auto* b : buffers) delete b; for (auto* i : images) delete i; }
};
struct PatternFrame final : FrameSource {
    uint32_t w, h, frame = 0;
    uint32_t FrameWidth() const override { return w; }
    uint32_t FrameHeight() const override { return h; }
    void ResolveRGBA8(uint8_t* dst, size_t size) override { for (size_t i = 0; i < size; ++i) dst[i] = (uint8_t)(i * 7 + frame * 13); ++frame; }
};
}  // namespace

int main() {
    MockDevice dev; MockQueue queue; MockPool pool;
    PatternFrame frame; frame.w = 64; frame.h = 36;
    PresentUploader up(&dev, &queue, &pool, frame.w, frame.h);
    const std::string want = "Map Unmap Begin BeginTransition SetImageTransition EndTransition CopyBufferToImage BeginTransition SetImageTransition EndTransition End Submit WaitIdle ";
    for (int f = 0; f < 2; ++f) {
        g_log.clear();
        up.Upload(frame);
        if (g_log != want) { std::printf("FAIL protocol (frame %d): %s\n", f, g_log.c_str()); return 1; }
        MockImage* im = static_cast<MockImage*>(up.Image());
        if (im->layout != RHIImageLayout::ShaderReadOnly || im->texels.size() != (size_t)64 * 36 * 4) { std::printf("FAIL image state\n"); return 1; }
        for (size_t i = 0; i < im->texels.size(); ++i) if (im->texels[i] != (uint8_t)(i * 7 + f * 13)) { std::printf("FAIL bytes at %zu\n", i); return 1; }
    }
    if (!dev.problems.empty() || queue.submitted != 2 || queue.waited != 2) { std::printf("FAIL descs: %s\n", dev.problems.c_str()); return 1; }
    std::printf("present path OK: 2 frames, %zu bytes each, protocol = Renderer.cpp:219-251\n", (size_t)64 * 36 * 4);
    return 0;
}

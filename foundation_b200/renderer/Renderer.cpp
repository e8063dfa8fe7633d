// Renderer.cpp — see Renderer.hpp.  Only foundation_pt_* calls; every non-zero status becomes a CHECK failure, so
// Foundation-side behaviour (fail fast with a message, src/Core/Core.hpp:17) is unchanged.
#include "Renderer.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace Foundation {
namespace Core {
void BugCheck(const char* what) {
    std::fprintf(stderr, "[Foundation] BugCheck: %s\n", what);
    std::abort();
}
}  // namespace Core

namespace Renderer {

namespace {
// Host-scratch callbacks over Core::Allocator — the same bridge as the reference's vkCustomCpuAllocation / vkCustomCpuFree
// (src/Platform/RHI/Vulkan/Application.cpp:93-110).
void* AllocCb(void* user, size_t size, size_t alignment) { return static_cast<Core::Allocator*>(user)->Allocate(size, alignment); }
void FreeCb(void* user, void* ptr) { static_cast<Core::Allocator*>(user)->Deallocate(ptr); }
}  // namespace

void Renderer::Check(int32_t status, const char* what) const {
    if (status == FOUNDATION_PT_OK) return;
    std::fprintf(stderr, "[Foundation] %s failed (%d): %s\n", what, status, foundation_pt_last_error(m_ctx));
    FOUNDATION_CHECK(status == FOUNDATION_PT_OK && "foundation_pt call failed");
}

Renderer::Renderer(DeviceHandle device, Core::Allocator* allocator, const SceneDesc& scene, uint64_t seed)
    : Renderer(std::vector<DeviceHandle>{device}, allocator, scene, seed, false) {}

template <class F> void Renderer::ForEachMember(F f) {
    for (uint32_t i = 0; i < foundation_pt_group_size(m_group); ++i) f(foundation_pt_group_context(m_group, i));
}

Renderer::Renderer(const std::vector<DeviceHandle>& devices, Core::Allocator* allocator, const SceneDesc& scene, uint64_t seed, bool direct_gather)
    : m_allocator(allocator), m_device(devices.empty() ? DeviceHandle{} : devices[0]), m_width(scene.width), m_height(scene.height) {
    FOUNDATION_CHECK(allocator != nullptr);
    FOUNDATION_CHECK(!devices.empty());
    foundation_pt_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg; cfg.width = scene.width; cfg.height = scene.height; cfg.seed = seed;
    std::memcpy(cfg.background, scene.background, sizeof cfg.background);
    foundation_pt_allocator cb{allocator, &AllocCb, &FreeCb};
    std::vector<int32_t> ordinals;
    for (const DeviceHandle& d : devices) ordinals.push_back(d.cuda_ordinal);
    int32_t st = foundation_pt_group_create(&cfg, ordinals.data(), (uint32_t)ordinals.size(), 32, direct_gather ? (uint32_t)FOUNDATION_PT_COMM_DIRECT : 0u, &cb, &m_group);
    if (st != FOUNDATION_PT_OK) std::fprintf(stderr, "[Foundation] foundation_pt_group_create failed (%d): %s\n", st, foundation_pt_last_error(nullptr));
    FOUNDATION_CHECK(st == FOUNDATION_PT_OK && "foundation_pt_group_create failed");
    m_ctx = foundation_pt_group_context(m_group, 0);
    // one-time blocking uploads on every member (the scene is replicated per device): the analogue of the staging copies at
    // src/Renderer/Renderer.cpp:133-197
    ForEachMember([&](foundation_pt_context* c) {
        foundation_pt_context* saved = m_ctx; m_ctx = c;     // Check() reports the failing member's message
        if (!scene.materials.empty()) Check(foundation_pt_materials_set(c, scene.materials.data(), (uint32_t)scene.materials.size()), "materials_set");
        for (const MeshDesc& m : scene.meshes) {
            uint32_t id = 0;
            Check(foundation_pt_mesh_create(c, m.positions.data(), 3 * sizeof(float), (uint32_t)(m.positions.size() / 3), m.indices.data(), FOUNDATION_PT_INDEX_U32,
                                            (uint32_t)(m.indices.size() / 3), m.material_ids.empty() ? nullptr : m.material_ids.data(), &id),
                  "mesh_create");
        }
        if (!scene.instances.empty()) Check(foundation_pt_instances_set(c, scene.instances.data(), (uint32_t)scene.instances.size()), "instances_set");
        m_build.struct_size = sizeof m_build;
        Check(foundation_pt_scene_commit(c, &m_build), "scene_commit");
        Check(foundation_pt_camera_set(c, scene.view, scene.proj), "camera_set");
        m_ctx = saved;
    });
    m_present_image = static_cast<uint8_t*>(m_allocator->Allocate((size_t)m_width * m_height * 4, 16));
    FOUNDATION_CHECK(m_present_image != nullptr);
}

Renderer::~Renderer() {
    // like the reference's destructor (Renderer.cpp:402-406) this waits for the device: destroy() synchronises every member's streams
    if (m_present_image) m_allocator->Deallocate(m_present_image);
    if (m_group) foundation_pt_group_destroy(m_group);
}

void Renderer::SetCamera(const float view[16], const float proj[16]) {
    ForEachMember([&](foundation_pt_context* c) { Check(foundation_pt_camera_set(c, view, proj), "camera_set"); });
    m_samples_done = 0;   // a moved camera restarts the progressive accumulation
}
void Renderer::SetQuality(uint32_t samples_per_draw, uint32_t max_bounces) {
    m_samples_per_draw = samples_per_draw ? samples_per_draw : 1; m_max_bounces = max_bounces;
}

void Renderer::SetInstances(const foundation_pt_instance* instances, uint32_t count) {
    ForEachMember([&](foundation_pt_context* c) {
        Check(foundation_pt_instances_set(c, instances, count), "instances_set");
        Check(foundation_pt_scene_commit(c, &m_build), "scene_commit");   // TLAS-only: the meshes did not change
    });
    m_samples_done = 0;
}

void Renderer::Draw() {
    int32_t st = foundation_pt_group_render(m_group, m_samples_done, m_samples_per_draw, m_max_bounces);   // all devices render their tiles; frame gathered into device 0
    if (st != FOUNDATION_PT_OK) std::fprintf(stderr, "[Foundation] group_render failed (%d): %s\n", st, foundation_pt_group_last_error(m_group));
    FOUNDATION_CHECK(st == FOUNDATION_PT_OK && "foundation_pt_group_render failed");
    m_samples_done += m_samples_per_draw;
    Check(foundation_pt_resolve_rgba8(m_ctx, m_present_image, (size_t)m_width * m_height * 4), "resolve_rgba8");
}

void Renderer::ResolveRGBA8(uint8_t* dst, size_t size_bytes) { Check(foundation_pt_resolve_rgba8(m_ctx, dst, size_bytes), "resolve_rgba8"); }

void Renderer::ReadAccum(std::vector<float>* rgba) const {
    rgba->resize((size_t)m_width * m_height * 4);
    Check(foundation_pt_read_accum(m_ctx, rgba->data(), rgba->size() * sizeof(float)), "read_accum");
}
foundation_pt_stats Renderer::Stats() const {
    foundation_pt_stats s; std::memset(&s, 0, sizeof s); s.struct_size = sizeof s;
    Check(foundation_pt_stats_get(m_ctx, &s), "stats_get");
    return s;
}

// ---- scene ingestion: "FPTS" container written by foundation_b200.scenes.save_scene ---------------------------------
bool SceneDesc::Load(const char* path, std::string* error) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { if (error) *error = std::string("cannot open ") + path; return false; }
    auto rd = [&](void* p, size_t n) { return std::fread(p, 1, n, f) == n; };
    uint32_t hdr[8];
    bool ok = rd(hdr, sizeof hdr) && hdr[0] == 0x53545046u /* 'FPTS' */ && hdr[1] == 1;
    if (ok) {
        width = hdr[2]; height = hdr[3];
        uint32_t nmesh = hdr[4], nmat = hdr[5], ninst = hdr[6];
        ok = rd(view, 64) && rd(proj, 64) && rd(background, 12);
        materials.resize(nmat); ok = ok && rd(materials.data(), (size_t)nmat * sizeof(foundation_pt_material));
        instances.resize(ninst); ok = ok && (ninst == 0 || rd(instances.data(), (size_t)ninst * sizeof(foundation_pt_instance)));
        meshes.resize(nmesh);
        for (uint32_t m = 0; ok && m < nmesh; ++m) {
            uint32_t c[2];
            ok = rd(c, 8);
            if (!ok) break;
            meshes[m].positions.resize((size_t)c[0] * 3); meshes[m].indices.resize((size_t)c[1] * 3); meshes[m].material_ids.resize(c[1]);
            ok = rd(meshes[m].positions.data(), (size_t)c[0] * 12) && rd(meshes[m].indices.data(), (size_t)c[1] * 12) && rd(meshes[m].material_ids.data(), (size_t)c[1] * 4);
        }
    }
    std::fclose(f);
    if (!ok && error) *error = std::string("malformed scene file ") + path;
    return ok;
}

}  // namespace Renderer
}  // namespace Foundation

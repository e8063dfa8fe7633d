// Editor.cpp — headless stand-in for the reference's only caller of the renderer (src/Editor/Editor.cpp:13-44):
// construct the allocator, the device handle and the Renderer, call Draw() in a loop, report allocator bytes at exit.
// There is no window on the GPU box, so the loop runs a fixed number of frames and writes the result to disk instead of presenting.
//   usage: foundation_editor [--gpus N] [--direct] <scene.fpts> <frames> <samples_per_draw> <max_bounces> [accum.raw] [out.ppm] [spin_degrees_per_frame] [out.pfm]
// --gpus N renders on CUDA devices 0..N-1 of the box (scene replicated, frame tile-partitioned, gathered over NVLink into device 0 inside
// Draw(); --direct selects the fused peer-store gather), where the reference takes EnumerateDevices()[0] only (src/Editor/Editor.cpp:18).
// With a spin angle every frame rotates all instances about +Z by frame * angle before drawing, like the reference's per-frame model
// matrix (src/Renderer/Renderer.cpp:373): TLAS-only rebuild + accumulation restart each Draw().
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "Renderer.hpp"

using namespace Foundation;

namespace {
// counting heap allocator, the role of HeapAllocatorMultiThreaded (src/Core/Allocator/HeapAllocator.hpp:9-70)
class CountingAllocator final : public Core::Allocator {
    std::atomic<size_t> m_used{0};
    struct Header { size_t size; void* base; };
public:
    pointer Allocate(size_type size) override { return Allocate(size, alignof(std::max_align_t)); }
    pointer Allocate(size_type size, size_t alignment) override {
        if (alignment < alignof(Header)) alignment = alignof(Header);
        void* base = std::malloc(size + alignment + sizeof(Header));
        if (!base) return nullptr;
        uintptr_t p = (reinterpret_cast<uintptr_t>(base) + sizeof(Header) + alignment - 1) & ~(uintptr_t)(alignment - 1);
        Header* h = reinterpret_cast<Header*>(p) - 1; h->size = size; h->base = base;
        m_used += size;
        return reinterpret_cast<void*>(p);
    }
    void Deallocate(pointer ptr, size_type) override { Deallocate(ptr); }
    void Deallocate(pointer ptr) override {
        if (!ptr) return;
        Header* h = reinterpret_cast<Header*>(ptr) - 1;
        m_used -= h->size; std::free(h->base);
    }
    pointer Reallocate(pointer ptr, size_type new_size, size_t alignment) override {
        pointer n = Allocate(new_size, alignment);
        if (ptr && n) { Header* h = reinterpret_cast<Header*>(ptr) - 1; std::memcpy(n, ptr, h->size < new_size ? h->size : new_size); Deallocate(ptr); }
        return n;
    }
    size_type GetUsedMemory() const noexcept override { return m_used.load(); }
};
CountingAllocator g_Allocator;
}  // namespace

int main(int argc, char** argv) {
    int gpus = 1; bool direct = false;
    {   // strip the options, keep the positional arguments
        int w = 1;
        for (int i = 1; i < argc; ++i) {
            if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
            else if (!std::strcmp(argv[i], "--direct")) direct = true;
            else argv[w++] = argv[i];
        }
        argc = w;
        if (gpus < 1) gpus = 1;
    }
    if (argc < 5) { std::fprintf(stderr, "usage: %s scene.fpts frames samples_per_draw max_bounces [accum.raw] [out.ppm] [spin] [out.pfm]\n", argv[0]); return 2; }
    Renderer::SceneDesc scene; std::string err;
    if (!scene.Load(argv[1], &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
    int frames = std::atoi(argv[2]); uint32_t spd = (uint32_t)std::atoi(argv[3]), bounces = (uint32_t)std::atoi(argv[4]);
    {
        std::vector<Renderer::DeviceHandle> devices;            // the reference takes EnumerateDevices()[0] (Editor.cpp:18); --gpus N takes the first N
        for (int d = 0; d < gpus; ++d) devices.push_back(Renderer::DeviceHandle{d});
        Renderer::Renderer renderer(devices, g_Allocator.Ptr(), scene, /*seed*/ 7, direct);
        renderer.SetQuality(spd, bounces);
        auto t0 = std::chrono::steady_clock::now();
        const double spin = argc > 7 ? std::atof(argv[7]) : 0.0;
        std::vector<foundation_pt_instance> moved(scene.instances);
        for (int i = 0; i < frames; ++i) {                      // Main Loop (Editor.cpp:20-23)
            if (spin != 0.0 && !moved.empty()) {
                const double a = (i + 1) * spin * 3.14159265358979323846 / 180.0;
                const float c = (float)std::cos(a), s = (float)std::sin(a);
                for (size_t k = 0; k < moved.size(); ++k) {     // model = rotate(angle, +Z) * original (rows of the 3x4)
                    const float* t = scene.instances[k].transform; float* o = moved[k].transform;
                    for (int col = 0; col < 4; ++col) { o[col] = c * t[col] - s * t[4 + col]; o[4 + col] = s * t[col] + c * t[4 + col]; o[8 + col] = t[8 + col]; }
                }
                renderer.SetInstances(moved.data(), (uint32_t)moved.size());
            }
            renderer.Draw();
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        foundation_pt_build_stats bs = renderer.BuildStats();
        std::printf("gpus=%u ", renderer.DeviceCount());
        std::printf("frames=%d spp=%u ms=%.3f spp_per_s=%.2f tris=%llu nodes8=%llu build_ms=%.2f\n", frames, renderer.SamplesDone(), ms,
                    renderer.SamplesDone() / (ms * 1e-3), (unsigned long long)bs.num_triangles, (unsigned long long)bs.num_nodes8, bs.build_ms);
        if (argc > 5) {   // the float4 accumulation buffer as is: radiance sum in rgb, sample count in alpha
            std::vector<float> acc; renderer.ReadAccum(&acc);
            FILE* f = std::fopen(argv[5], "wb");
            if (f) { std::fwrite(acc.data(), sizeof(float), acc.size(), f); std::fclose(f); }   // raw float4 dump (row-major from the top-left)
        }
        if (argc > 8) {   // mean linear radiance as a colour PFM: little-endian (scale -1.0), rows bottom-up
            std::vector<float> acc; renderer.ReadAccum(&acc);
            FILE* f = std::fopen(argv[8], "wb");
            if (f) {
                const uint32_t w = renderer.Width(), h = renderer.Height();
                std::fprintf(f, "PF\n%u %u\n-1.0\n", w, h);
                std::vector<float> row(3 * (size_t)w);
                for (uint32_t y = h; y-- > 0;) {
                    for (uint32_t x = 0; x < w; ++x) {
                        const float* a = &acc[4 * ((size_t)y * w + x)];
                        for (int c = 0; c < 3; ++c) row[3 * x + c] = a[3] > 0.0f ? a[c] / a[3] : 0.0f;
                    }
                    std::fwrite(row.data(), sizeof(float), row.size(), f);
                }
                std::fclose(f);
            }
        }
        if (argc > 6) {
            FILE* f = std::fopen(argv[6], "wb");
            if (f) {
                std::fprintf(f, "P6\n%u %u\n255\n", renderer.Width(), renderer.Height());
                const uint8_t* img = renderer.PresentImage();
                for (size_t p = 0; p < (size_t)renderer.Width() * renderer.Height(); ++p) std::fwrite(img + 4 * p, 1, 3, f);
                std::fclose(f);
            }
        }
    }
    std::printf("Quitting. Memory Used: %zu bytes\n", g_Allocator.GetUsedMemory());   // Editor.cpp:25
    return 0;
}

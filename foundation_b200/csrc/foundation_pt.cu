// foundation_pt.cu — the C ABI (include/foundation_pt.h) over the sm_100a kernels in pt_kernels.cuh.
// Host side of the drop-in boundary: context + scene upload (the slot of Renderer::Renderer's blocking
// staging uploads, mos9527/Foundation src/Renderer/Renderer.cpp:133-197), acceleration-structure build,
// the wavefront render loop (the slot of Renderer::Record's pass body, Renderer.cpp:332-351) and the
// explicit-ray-set interface used for parity and the Mrays/s metric.
// There is NO CPU fallback: without a CUDA device foundation_pt_create fails with FOUNDATION_PT_ERR_NO_DEVICE.
#include <dlfcn.h>
#include <nccl.h>   // types only: the NCCL entry points are resolved at run time (dlopen), the library has no link-time NCCL dependency

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/foundation_pt.h"
#ifndef PT_COLLAPSE_PERSISTENT
#define PT_COLLAPSE_PERSISTENT 1   // 1: all collapse levels beyond the top in one cooperative kernel (k_collapse_levels); 0: round 1's per-level kernels + host round trips
#endif
#ifndef PT_ONESWEEP
#define PT_ONESWEEP 1        // 1 (default since the ranking moved to shared-memory atomicOr, 3072-key tiles): one up-front histogram kernel + ONE kernel per pass with
                             // per-digit decoupled look-back; 0: histogram + look-back scan + scatter kernel per pass.  10 M-pair sort, same box
                             // (profiles/r02_ab_sort_ranking.log): 0.94 ms vs 1.05 ms (and 1.13 ms for round 1's match_any three-kernel form).  With match_any
                             // ranking and 2304-key tiles the one-sweep form had been the slower one (1.20 vs 1.13 ms, profiles/r02_ab_traversal_build.log).
#endif
#ifndef PT_AGGLOMERATIVE
#define PT_AGGLOMERATIVE 1   // 1: the radix tree is built bottom-up inside the refit (k_refit_agg, k_refit_agg_up); 0: k_karras + k_refit + k_refit_up
#endif
#include "pt_kernels.cuh"

static_assert(sizeof(foundation_pt_ray) == 32 && sizeof(foundation_pt_hit) == 16, "ABI record sizes");
static_assert(sizeof(foundation_pt_material) == sizeof(PtMaterial) && sizeof(PtMaterial) == 32, "material layout");
static_assert(sizeof(PtInstance) == 112 && sizeof(PtTri) == 48 && sizeof(PtLight) == 64, "device record sizes");
static_assert(sizeof(foundation_pt_instance) == 64, "instance ABI size");

namespace {

thread_local std::string g_create_error = "no error";

// Device allocations.  While g_pool_stream is set (scene_commit) they are stream-ordered (cudaMallocAsync on the context's
// stream, pool release threshold raised at create) so the ~30 scratch buffers of a build cost microseconds instead of the
// milliseconds cudaMalloc / cudaFree take each (and cudaFree synchronises the device).
thread_local cudaStream_t g_pool_stream = nullptr;
struct DevBuf {
    void* p = nullptr; size_t bytes = 0; cudaStream_t owner = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), owner(o.owner) { o.p = nullptr; o.bytes = 0; o.owner = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; owner = o.owner; o.p = nullptr; o.bytes = 0; o.owner = nullptr; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) { if (owner) cudaFreeAsync(p, owner); else cudaFree(p); }
        p = nullptr; bytes = 0; owner = nullptr;
    }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        cudaError_t e;
        if (g_pool_stream) { e = cudaMallocAsync(&p, n, g_pool_stream); owner = g_pool_stream; }
        else { e = cudaMalloc(&p, n); owner = nullptr; }
        if (e == cudaSuccess) bytes = n; else { p = nullptr; owner = nullptr; }
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PoolScope {   // RAII: stream-ordered allocation inside one API call
    explicit PoolScope(cudaStream_t s) { g_pool_stream = s; }
    ~PoolScope() { g_pool_stream = nullptr; }
};

// std allocator over the caller's host-allocation callbacks (foundation_pt_allocator; NULL callbacks = malloc/free):
// every host-side copy of scene data the context keeps goes through it, like the reference threads Core::Allocator
// through its containers (src/Core/Allocator/StlContainers.hpp:11-66).
template <class T>
struct CbAlloc {
    using value_type = T;
    const foundation_pt_allocator* cb = nullptr;
    CbAlloc() = default;
    explicit CbAlloc(const foundation_pt_allocator* c) : cb(c) {}
    template <class U> CbAlloc(const CbAlloc<U>& o) : cb(o.cb) {}
    T* allocate(size_t n) {
        void* p = (cb && cb->alloc) ? cb->alloc(cb->user, n * sizeof(T), alignof(T) < 16 ? 16 : alignof(T)) : std::malloc(n * sizeof(T));
        if (!p) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { if (cb && cb->free) cb->free(cb->user, p); else std::free(p); }
    template <class U> bool operator==(const CbAlloc<U>& o) const { return cb == o.cb; }
    template <class U> bool operator!=(const CbAlloc<U>& o) const { return cb != o.cb; }
};
template <class T> using HostVec = std::vector<T, CbAlloc<T>>;

struct Mesh {
    // host copies (light extraction, validation), allocated through the caller's allocator
    HostVec<uint8_t> h_pos; HostVec<uint8_t> h_idx; HostVec<uint32_t> h_mat;
    explicit Mesh(const foundation_pt_allocator* cb) : h_pos(CbAlloc<uint8_t>(cb)), h_idx(CbAlloc<uint8_t>(cb)), h_mat(CbAlloc<uint32_t>(cb)) {}
    Mesh(Mesh&&) = default; Mesh& operator=(Mesh&&) = default;
    uint32_t stride = 0, idx_fmt = 0, nverts = 0, ntris = 0;
    DevBuf d_pos, d_idx, d_mat;
    DevBuf d_uv, d_col;      // optional per-vertex attributes, tightly packed (float2 / float3)
    // build products
    DevBuf d_nodes, d_tris, d_order;
    uint32_t num_nodes = 0;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, pad = 0;
    // emissive triangles of the mesh (ascending), valid for the emissive set of the materials they were extracted with: a commit that only moves instances
    // or rebuilds a deformed mesh reuses them (one kernel and two host round trips less per mesh and commit)
    std::vector<uint32_t> em_tris; std::vector<uint8_t> em_for; bool em_valid = false;
    PtMeshRaw raw() const { PtMeshRaw r; r.pos = d_pos.as<uint8_t>(); r.stride = stride; r.idx = d_idx.p; r.idx_fmt = idx_fmt; r.ntris = ntris; r.mat = d_mat.as<uint32_t>(); return r; }
    void host_tri(uint32_t i, float* v9) const {
        uint32_t id[3];
        for (int k = 0; k < 3; ++k) {
            if (idx_fmt == 32) id[k] = reinterpret_cast<const uint32_t*>(h_idx.data())[3 * (size_t)i + k];
            else if (idx_fmt == 16) id[k] = reinterpret_cast<const uint16_t*>(h_idx.data())[3 * (size_t)i + k];
            else id[k] = 3 * i + k;
            memcpy(v9 + 3 * k, h_pos.data() + (size_t)id[k] * stride, 12);
        }
    }
};

}  // namespace

struct foundation_pt_context {
    foundation_pt_config cfg{};
    foundation_pt_allocator host_alloc{};
    int device = 0, num_sms = 0, trace_blocks_per_sm = 0 /* 0 = kernel's own: 8 flat, 6 two-level */, fetch_thresh = 24;
    int e2e_chunk_log2 = 22, e2e_tail_log2 = 20;   // host-buffer trace calls: rays per pipeline chunk, and per chunk over the last 2^(e2e_chunk_log2 + 1) rays
    int collapse_blocks = 0; // grid of the persistent collapse kernel: one resident wave on this device
    int refit_blocks = 0;   // occupancy of the tiled refit kernel on this context's device (queried at the first build)
    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;   // compute, H2D, D2H
    // two waves of a render in flight at once: wave i runs on stream (i even) / stream_b (i odd) with its own copy of the wavefront state, so the tail of every
    // traversal kernel — a persistent grid whose last warps finish long rays alone — and the latency-bound late bounces overlap the other wave's kernels
    cudaStream_t stream_b = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_acc[2] = {nullptr, nullptr};
    int wave_sets = 1;             // copies of the wavefront state (1 or 2)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    mutable std::string err = "no error";

    std::vector<Mesh> meshes;
    std::vector<PtMaterial> mats;
    std::vector<foundation_pt_instance> insts; bool has_insts = false;

    // committed scene
    bool committed = false, two_level = false, render_pending = false;
    std::vector<cudaEvent_t> stage_events;            // FOUNDATION_PT_FLAG_STAGE_TIMING: pool of events, one per stage boundary
    std::vector<int> stage_marks;                     // stage that ENDS at event i (-1 = start of the render)
    DevBuf d_nodes_all, d_tris_all, d_instances, d_inst_in, d_mesh_info, d_mats, d_lights;
    std::vector<PtMeshInfo> mesh_info;
    DevBuf d_tlas_order; uint32_t tlas_nodes = 0, num_inst = 0;
    bool flat_valid = false, tlas_only_commit = false; size_t flat_meshes = 0; uint32_t flat_tlas_cap = 0;
    PtSceneView view{};
    uint32_t num_lights = 0; float light_area = 0, ray_eps = 0;
    size_t l2_window_bytes = 0, l2_carve_bytes = 0;
    bool use_pool = false;
    float wlo[3] = {0, 0, 0}, whi[3] = {0, 0, 0};
    foundation_pt_build_stats bstats{};

    PtCamera cam{}; bool cam_set = false;

    // vertex attributes + albedo textures (the reference's material inputs: vertex colour, uv, one RGBA8 texture)
    struct Texture { DevBuf texels; uint32_t width = 0, height = 0; };
    std::vector<Texture> textures;
    std::vector<uint32_t> mat_tex;                       // per material: texture id or PT_NONE
    DevBuf d_mesh_attr, d_textures, d_mat_tex, w_hit_uv;
    bool attr_enabled = false;

    // wavefront state
    uint32_t part_rank = 0, part_count = 1, part_tile = 32;
    bool wave_ready = false;
    DevBuf w_ray_o, w_ray_d, w_beta, w_L, w_rng, w_hit, w_active, w_next, w_sorted, w_sh_o, w_sh_d, w_sh_c, w_slot_pixel, w_ctr, w_keyhist, d_accum;
    uint32_t num_slots = 0;      // owned pixels
    uint32_t wave_samples = 1;   // samples of every owned pixel traced together in one wave
    DevBuf d_status, d_counters;

    // explicit ray set
    DevBuf d_rays, d_hits, d_hit_inst, d_occ; uint64_t num_rays = 0;

    // multi-GPU frame (stage C1): communicator over the ranks that share one frame
    ncclComm_t comm = nullptr; uint32_t comm_rank = 0, comm_count = 1, comm_flags = 0; bool comm_owned = false;
    float4* remote_accum = nullptr; bool remote_is_ipc = false;     // direct mode: the root's accumulation buffer as seen from this device
    DevBuf d_pack, d_stage, d_stage_off, d_row_base, d_comm_scratch;
    std::vector<uint32_t> rank_pixels, stage_off;
    float gather_ms = 0;

    foundation_pt_stats stats{};
    uint64_t total_launches = 0; uint32_t call_launches = 0;

    int32_t fail(int32_t code, const std::string& msg) const { err = msg; return code; }
};

namespace {

typedef foundation_pt_context Ctx;

#define PT_CK(expr)                                                                                            \
    do {                                                                                                       \
        cudaError_t e_ = (expr);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return ctx->fail(e_ == cudaErrorMemoryAllocation ? FOUNDATION_PT_ERR_OOM : FOUNDATION_PT_ERR_CUDA, \
                             std::string(#expr) + ": " + cudaGetErrorString(e_));                              \
    } while (0)

#define PT_LAUNCH(ctx, kernel, grid, block, ...)                        \
    do {                                                                \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);     \
        (ctx)->call_launches++; (ctx)->total_launches++;                \
    } while (0)

#define PT_LAUNCH_ON(ctx, st, kernel, grid, block, ...)                 \
    do {                                                                \
        kernel<<<(grid), (block), 0, (st)>>>(__VA_ARGS__);              \
        (ctx)->call_launches++; (ctx)->total_launches++;                \
    } while (0)

inline uint32_t grid_for(const Ctx* ctx, uint64_t n, uint32_t block, uint32_t blocks_per_sm) {
    uint64_t need = (n + block - 1) / block;
    uint64_t cap = (uint64_t)ctx->num_sms * blocks_per_sm;
    if (need < 1) need = 1;
    return (uint32_t)(need < cap ? need : cap);
}

inline uint32_t trace_blocks(const Ctx* ctx) {   // persistent grid of the traversal kernels: resident CTAs per SM x SMs
    return ctx->trace_blocks_per_sm ? (uint32_t)ctx->trace_blocks_per_sm : (uint32_t)PT_TRACE_MIN_BLOCKS(ctx->two_level);
}

// ------------------------------------------------------------------------------------------------
// device exclusive scan (in place allowed); total written to d_total (device uint32) if non-null
// ------------------------------------------------------------------------------------------------
[[maybe_unused]] int32_t scan_u32(Ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n, DevBuf& chunk_sums, uint32_t* d_total) {   // used by the non-default sort / collapse paths
    uint32_t chunks = (n + PT_SCAN_CHUNK - 1) / PT_SCAN_CHUNK;
    if (chunks == 0) chunks = 1;
    const size_t state_bytes = ((size_t)chunks + 1) * 8;          // one status word per tile + the ticket counter
    if (chunk_sums.bytes < state_bytes) PT_CK(chunk_sums.alloc(state_bytes + 4096));
    PT_CK(cudaMemsetAsync(chunk_sums.p, 0, state_bytes, ctx->stream));
    PT_LAUNCH(ctx, k_scan_lookback, chunks, PT_SCAN_THREADS, in, out, n, chunk_sums.as<unsigned long long>(), chunks, d_total);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// A2: radix sort driver. keys/vals end up in (keys_a, vals_a) (8 passes = even number of swaps).
// ------------------------------------------------------------------------------------------------
int32_t radix_sort(Ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n, DevBuf& hist, DevBuf& chunk_sums) {
#if PT_ONESWEEP
    (void)chunk_sums;
    const uint32_t tiles = (n + PT_OS_TILE - 1) / PT_OS_TILE;
    if (tiles == 0) return 0;
    // [8 x 256 digit histograms | 8 tickets (padded to 64 words) | 8 x tiles x 256 status words], zeroed by one memset
    const size_t words = 8 * 256 + 64 + (size_t)8 * tiles * 256;
    if (hist.bytes < words * 4) PT_CK(hist.alloc(words * 4));
    PT_CK(cudaMemsetAsync(hist.p, 0, words * 4, ctx->stream));
    uint32_t* d_hist = hist.as<uint32_t>(); uint32_t* d_ticket = d_hist + 8 * 256; uint32_t* d_status = d_ticket + 64;
    PT_LAUNCH(ctx, k_rs_hist_all, grid_for(ctx, n, 256, 8), 256, keys_a, n, d_hist);
    PT_LAUNCH(ctx, k_rs_hist_scan, 8, 256, d_hist);
    for (int pass = 0; pass < 8; ++pass) {
        PT_LAUNCH(ctx, k_rs_onesweep, tiles, PT_OS_THREADS, keys_a, vals_a, keys_b, vals_b, n, 8 * pass, d_hist + 256 * pass, d_status + (size_t)pass * tiles * 256, d_ticket + pass);
        std::swap(keys_a, keys_b); std::swap(vals_a, vals_b);
    }
    return 0;
#else
    uint32_t tiles = (n + PT_RS_TILE - 1) / PT_RS_TILE;
    if (tiles == 0) return 0;
    if (hist.bytes < (size_t)tiles * 256 * 4) PT_CK(hist.alloc((size_t)tiles * 256 * 4));
    for (int pass = 0; pass < 8; ++pass) {
        int shift = 8 * pass;
        PT_LAUNCH(ctx, k_rs_hist, tiles, PT_RS_THREADS, keys_a, n, shift, hist.as<uint32_t>(), tiles);
        int32_t rc = scan_u32(ctx, hist.as<uint32_t>(), hist.as<uint32_t>(), tiles * 256, chunk_sums, nullptr);
        if (rc) return rc;
        PT_LAUNCH(ctx, k_rs_scatter, tiles, PT_RS_THREADS, keys_a, vals_a, keys_b, vals_b, n, shift, hist.as<uint32_t>(), tiles);
        std::swap(keys_a, keys_b); std::swap(vals_a, vals_b);
    }
    return 0;
#endif
}

// ------------------------------------------------------------------------------------------------
// Generic LBVH -> BVH8 build over n primitives whose boxes are in d_prim_box and whose Morton keys are
// already in (keys, vals).  Outputs: nodes (exact size), leaf_seq (n), order (n) = sorted vals.
// ------------------------------------------------------------------------------------------------
struct BuildOut { DevBuf nodes, leaf_seq, order; uint32_t num_nodes = 0; float sort_ms = 0; };

int32_t build_bvh8(Ctx* ctx, uint32_t n, const PtBox* d_prim_box, DevBuf& keys, DevBuf& vals, const PtBuildParams* d_bp, uint32_t max_leaf, BuildOut* out) {
    DevBuf keys_b, vals_b, hist, chunk_sums;
    PT_CK(keys_b.alloc((size_t)n * 8)); PT_CK(vals_b.alloc((size_t)n * 4));
    PT_CK(cudaEventRecord(ctx->ev2, ctx->stream));
    int32_t rc = radix_sort(ctx, keys.as<uint64_t>(), vals.as<uint32_t>(), keys_b.as<uint64_t>(), vals_b.as<uint32_t>(), n, hist, chunk_sums);
    if (rc) return rc;
    PT_CK(cudaEventRecord(ctx->ev3, ctx->stream));
    keys_b.release(); vals_b.release(); hist.release();
    // BVH2
    DevBuf left, right, first, last, parent, box, flags, cost, plan;
    size_t ni = n > 1 ? n - 1 : 1;
    PT_CK(cost.alloc(ni * 32)); PT_CK(plan.alloc(ni * 8));
    PT_CK(left.alloc(ni * 4)); PT_CK(right.alloc(ni * 4)); PT_CK(first.alloc(ni * 4)); PT_CK(last.alloc(ni * 4));
#if !PT_AGGLOMERATIVE
    PT_CK(parent.alloc((2 * (size_t)n) * 4));
#endif
    PT_CK(box.alloc((2 * (size_t)n) * sizeof(PtBox))); PT_CK(flags.alloc(ni * 4));
    PT_CK(cudaMemsetAsync(flags.p, 0, ni * 4, ctx->stream));
    PtBvh2 b; b.n = n; b.left = left.as<uint32_t>(); b.right = right.as<uint32_t>(); b.first = first.as<uint32_t>(); b.last = last.as<uint32_t>();
    b.parent = parent.as<uint32_t>(); b.box = box.as<PtBox>(); b.cost = cost.as<float>(); b.plan = plan.as<uint64_t>();
#if !PT_AGGLOMERATIVE
    if (n > 1) PT_LAUNCH(ctx, k_karras, grid_for(ctx, n - 1, 256, 8), 256, keys.as<uint64_t>(), b);
#endif
    DevBuf root_ref;                 // BVH2 ref of the root: 0 with Karras' numbering, the top split with the agglomerative build
    PT_CK(root_ref.alloc(16));
    PT_CK(cudaMemsetAsync(root_ref.p, 0, 16, ctx->stream));
    {   // A4 (+ A3 when agglomerative): tile-local part (shared memory, round-synchronous), then the few subtree roots per tile climb the upper levels
        DevBuf up_list, up_count;   // freed stream-ordered when the scope ends
        PT_CK(up_list.alloc((size_t)n * 4)); PT_CK(up_count.alloc(16));
        PT_CK(cudaMemsetAsync(up_count.p, 0, 4, ctx->stream));
        int& refit_blocks = ctx->refit_blocks;     // resident blocks per SM (shared-memory bound), per context / device: the grid is exactly one wave, tiles are strided over it
#if PT_AGGLOMERATIVE
        if (!refit_blocks && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&refit_blocks, k_refit_agg, PT_REFIT_THREADS, 0) != cudaSuccess || refit_blocks < 1)) { cudaGetLastError(); refit_blocks = 4; }
        PT_LAUNCH(ctx, k_refit_agg, grid_for(ctx, n, PT_REFIT_TILE, (uint32_t)refit_blocks), PT_REFIT_THREADS, b, keys.as<uint64_t>(), d_prim_box, vals.as<uint32_t>(),
                  up_list.as<uint32_t>(), up_count.as<uint32_t>(), root_ref.as<uint32_t>(), max_leaf);
        if (n > 1) PT_LAUNCH(ctx, k_refit_agg_up, grid_for(ctx, n / 16 + 1, 128, 8), 128, b, keys.as<uint64_t>(), up_list.as<uint32_t>(), up_count.as<uint32_t>(),
                             flags.as<uint32_t>(), root_ref.as<uint32_t>(), max_leaf);
#else
        if (!refit_blocks && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&refit_blocks, k_refit, PT_REFIT_TILE, 0) != cudaSuccess || refit_blocks < 1)) { cudaGetLastError(); refit_blocks = 4; }
        PT_LAUNCH(ctx, k_refit, grid_for(ctx, n, PT_REFIT_TILE, (uint32_t)refit_blocks), PT_REFIT_TILE, b, d_prim_box, vals.as<uint32_t>(), up_list.as<uint32_t>(), up_count.as<uint32_t>(), max_leaf);
        if (n > 1) PT_LAUNCH(ctx, k_refit_up, grid_for(ctx, n / 16 + 1, 128, 8), 128, b, up_list.as<uint32_t>(), up_count.as<uint32_t>(), flags.as<uint32_t>(), max_leaf);
#endif
    }
    // collapse, level by level
    DevBuf nodes_tmp, refs_a, refs_b, slots, n_int, n_prim, totals;
    PT_CK(nodes_tmp.alloc((size_t)n * sizeof(PtNode8)));
    PT_CK(refs_a.alloc((size_t)n * 4)); PT_CK(refs_b.alloc((size_t)n * 4));
    PT_CK(out->leaf_seq.alloc((size_t)n * 4));
    PT_CK(totals.alloc(16));
    PT_CK(cudaMemcpyAsync(refs_a.p, root_ref.p, 4, cudaMemcpyDeviceToDevice, ctx->stream));
#if !PT_COLLAPSE_PERSISTENT
    uint32_t m = 1;
#endif
    uint32_t level_start = 0, prim_total = 0;
    {   // top of the tree: all levels of at most PT_TOP_NODES wide nodes in one single-block launch
        PT_LAUNCH(ctx, k_collapse_top, 1, PT_TOP_NODES, b, refs_a.as<uint32_t>(), refs_b.as<uint32_t>(), max_leaf, d_bp, nodes_tmp.as<PtNode8>(), out->leaf_seq.as<uint32_t>(),
                  totals.as<uint32_t>());
#if !PT_COLLAPSE_PERSISTENT
        uint32_t st[4];
        PT_CK(cudaMemcpyAsync(st, totals.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
        PT_CK(cudaStreamSynchronize(ctx->stream));
        m = st[0]; level_start = st[1]; prim_total = st[2];
        if (st[3]) std::swap(refs_a, refs_b);
#endif
    }
#if PT_COLLAPSE_PERSISTENT
    {
        // every remaining level in one cooperative launch: the grid is one resident wave, levels are separated by a grid barrier and the
        // level totals never leave the device
        if (!ctx->collapse_blocks) {
            int per_sm = 0, coop = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_collapse_levels, PT_CL_THREADS, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
            if (!coop) return ctx->fail(FOUNDATION_PT_ERR_UNSUPPORTED, "device cannot launch cooperative kernels");
            ctx->collapse_blocks = per_sm * ctx->num_sms;
        }
        const uint32_t cap_level = n / 2 + 256;          // a wide node below the root stands for >= 2 primitives and the nodes of a level are disjoint
        DevBuf block_sums, barrier;
        PT_CK(slots.alloc((size_t)cap_level * 32)); PT_CK(n_int.alloc((size_t)cap_level * 4)); PT_CK(n_prim.alloc((size_t)cap_level * 4));
        PT_CK(block_sums.alloc((size_t)ctx->collapse_blocks * 8)); PT_CK(barrier.alloc(16));
        PT_CK(cudaMemsetAsync(barrier.p, 0, 16, ctx->stream));
        PtCollapseArgs ca;
        ca.b = b; ca.refs_a = refs_a.as<uint32_t>(); ca.refs_b = refs_b.as<uint32_t>(); ca.max_leaf = max_leaf; ca.bp = d_bp; ca.nodes = nodes_tmp.as<PtNode8>();
        ca.leaf_seq = out->leaf_seq.as<uint32_t>(); ca.state = totals.as<uint32_t>(); ca.slots = slots.as<uint32_t>(); ca.n_int = n_int.as<uint32_t>();
        ca.n_prim = n_prim.as<uint32_t>(); ca.cap = cap_level; ca.block_sums = block_sums.as<uint32_t>(); ca.barrier = barrier.as<uint32_t>(); ca.node_cap = n;
        // k_collapse_top left {m, level_start, prim_total, parity} in `totals`: the kernel takes them from there (no host round trip; with m = 0 —
        // a tree that fits the top kernel — it returns at once)
        void* kargs[] = {&ca};
        PT_CK(cudaLaunchCooperativeKernel((const void*)k_collapse_levels, dim3((unsigned)ctx->collapse_blocks), dim3(PT_CL_THREADS), kargs, 0, ctx->stream));
        ctx->call_launches++; ctx->total_launches++;
        uint32_t st[4];
        PT_CK(cudaMemcpyAsync(st, totals.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
        PT_CK(cudaStreamSynchronize(ctx->stream));
        if (st[3] || st[0]) return ctx->fail(FOUNDATION_PT_ERR_STATE, "BVH8 collapse exceeded node capacity");
        level_start = st[1]; prim_total = st[2];
    }
#else
    size_t cap = 0;
    while (m > 0) {
        if (cap < m) {
            cap = (size_t)m + m / 2 + 64;
            PT_CK(slots.alloc(cap * 32)); PT_CK(n_int.alloc(cap * 4)); PT_CK(n_prim.alloc(cap * 4));
        }
        if ((uint64_t)level_start + m > n) return ctx->fail(FOUNDATION_PT_ERR_STATE, "BVH8 collapse exceeded node capacity");
        PT_LAUNCH(ctx, k_collapse_select, grid_for(ctx, m, 128, 16), 128, b, refs_a.as<uint32_t>(), m, max_leaf, slots.as<uint32_t>(), n_int.as<uint32_t>(),
                  n_prim.as<uint32_t>());
        rc = scan_u32(ctx, n_int.as<uint32_t>(), n_int.as<uint32_t>(), m, chunk_sums, totals.as<uint32_t>());
        if (rc) return rc;
        rc = scan_u32(ctx, n_prim.as<uint32_t>(), n_prim.as<uint32_t>(), m, chunk_sums, totals.as<uint32_t>() + 1);
        if (rc) return rc;
        uint32_t next_start = level_start + m;
        PT_LAUNCH(ctx, k_collapse_emit, grid_for(ctx, m, 128, 16), 128, b, refs_a.as<uint32_t>(), m, max_leaf, d_bp, slots.as<uint32_t>(), n_int.as<uint32_t>(),
                  n_prim.as<uint32_t>(), level_start, next_start, prim_total, nodes_tmp.as<PtNode8>(), refs_b.as<uint32_t>(), out->leaf_seq.as<uint32_t>());
        uint32_t tot[2];
        PT_CK(cudaMemcpyAsync(tot, totals.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PT_CK(cudaStreamSynchronize(ctx->stream));
        prim_total += tot[1]; level_start = next_start; m = tot[0];
        std::swap(refs_a, refs_b);
    }
#endif
    if (prim_total != n) return ctx->fail(FOUNDATION_PT_ERR_STATE, "BVH8 collapse lost primitives");
    out->num_nodes = level_start;
    PT_CK(out->nodes.alloc((size_t)level_start * sizeof(PtNode8)));
    PT_CK(cudaMemcpyAsync(out->nodes.p, nodes_tmp.p, (size_t)level_start * sizeof(PtNode8), cudaMemcpyDeviceToDevice, ctx->stream));
    // no synchronisation here: the scratch buffers are released in stream order, and ev3 lies before the collapse's read-back above
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3); out->sort_ms = ms;
    out->order = std::move(vals);
    return 0;
}

int32_t build_blas(Ctx* ctx, Mesh& m, float* sort_ms) {
    uint32_t n = m.ntris;
    DevBuf prim_box, bounds, bp, keys, vals;
    PT_CK(prim_box.alloc((size_t)n * sizeof(PtBox))); PT_CK(bounds.alloc(32)); PT_CK(bp.alloc(sizeof(PtBuildParams)));
    PT_CK(keys.alloc((size_t)n * 8)); PT_CK(vals.alloc((size_t)n * 4));
    uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    PT_CK(cudaMemcpyAsync(bounds.p, init, 24, cudaMemcpyHostToDevice, ctx->stream));
    PtMeshRaw raw = m.raw();
    PT_LAUNCH(ctx, k_tri_boxes, grid_for(ctx, n, 256, 8), 256, raw, prim_box.as<PtBox>(), bounds.as<uint32_t>());
    PT_LAUNCH(ctx, k_build_params, 1, 32, bounds.as<uint32_t>(), bp.as<PtBuildParams>());
    PT_LAUNCH(ctx, k_morton_tris, grid_for(ctx, n, 256, 8), 256, raw, bp.as<PtBuildParams>(), keys.as<uint64_t>(), vals.as<uint32_t>());
    BuildOut out;
    int32_t rc = build_bvh8(ctx, n, prim_box.as<PtBox>(), keys, vals, bp.as<PtBuildParams>(), ctx->cfg.max_leaf_tris, &out);
    if (rc) return rc;
    PT_CK(m.d_tris.alloc((size_t)n * sizeof(PtTri)));
    PT_LAUNCH(ctx, k_write_tris, grid_for(ctx, n, 256, 8), 256, raw, out.order.as<uint32_t>(), out.leaf_seq.as<uint32_t>(), m.d_tris.as<PtTri>());
    PtBuildParams hbp;
    PT_CK(cudaMemcpyAsync(&hbp, bp.p, sizeof hbp, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    memcpy(m.lo, hbp.lo, 12); memcpy(m.hi, hbp.hi, 12); m.pad = hbp.pad;
    m.d_nodes = std::move(out.nodes); m.num_nodes = out.num_nodes; m.d_order = std::move(out.order);
    *sort_ms += out.sort_ms;
    return 0;
}

int32_t ensure_status(Ctx* ctx) {
    if (!ctx->d_status.p) {
        PT_CK(ctx->d_status.alloc(16)); PT_CK(ctx->d_counters.alloc(sizeof(PtDevCounters)));
        PT_CK(cudaMemsetAsync(ctx->d_status.p, 0, 16, ctx->stream));
    }
    return 0;
}
int32_t check_status(Ctx* ctx) {
    uint32_t st = 0;
    PT_CK(cudaMemcpyAsync(&st, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    if (st) { cudaMemsetAsync(ctx->d_status.p, 0, 16, ctx->stream); return ctx->fail(FOUNDATION_PT_ERR_STATE, "traversal stack overflow (PT_STACK_SIZE)"); }
    return 0;
}

// Keep the BVH8 node array resident in L2 (B200: 126 MB L2; 10 M triangles -> 111 MB of nodes): the traversal is latency-bound on
// node fetches, the ray / hit streams are marked evict-first.  Best effort: silently skipped if the device refuses.
void pin_nodes_in_l2(Ctx* ctx, size_t node_bytes) {
    if (!getenv("FOUNDATION_PT_L2_PIN")) return;   // opt-in: measured SLOWER on the 10 M-triangle terrain (4.93 vs 5.57 Grays/s): the carve-out
                                                    // takes L2 away from the 480 MB triangle array, which misses either way
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) { cudaGetLastError(); return; }
    size_t carve = (size_t)prop.persistingL2CacheMaxSize;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
    size_t win = node_bytes < (size_t)prop.accessPolicyMaxWindowSize ? node_bytes : (size_t)prop.accessPolicyMaxWindowSize;
    cudaStreamAttrValue attr; memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = const_cast<PtU4*>(ctx->view.nodes);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = win <= carve ? 1.0f : (float)((double)carve / (double)win);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    ctx->l2_window_bytes = win; ctx->l2_carve_bytes = carve;
}

template <bool ANY>
int32_t launch_trace(Ctx* ctx, const float4* rays, uint64_t n, float4* hits, uint32_t* inst, uint8_t* occ) {
    if (n == 0) return 0;
    uint32_t grid = grid_for(ctx, n, 128, trace_blocks(ctx));
    unsigned long long* wc = reinterpret_cast<unsigned long long*>(ctx->d_status.as<uint32_t>() + 2);
    PT_CK(cudaMemsetAsync(wc, 0, 8, ctx->stream));
    if (ctx->two_level) PT_LAUNCH(ctx, (k_trace_rays<ANY, true, false>), grid, 128, ctx->view, rays, (unsigned long long)n, hits, inst, occ, ctx->d_status.as<uint32_t>(), nullptr, wc, ctx->fetch_thresh);
    else PT_LAUNCH(ctx, (k_trace_rays<ANY, false, false>), grid, 128, ctx->view, rays, (unsigned long long)n, hits, inst, occ, ctx->d_status.as<uint32_t>(), nullptr, wc, ctx->fetch_thresh);
    PT_CK(cudaGetLastError());
    return 0;
}

// An asynchronous render still in flight is settled (foundation_pt_wait) before any other call touches the context: every entry point
// except render_async behaves as if the context were idle.
void settle(Ctx* ctx);
void begin_call(Ctx* ctx) { settle(ctx); ctx->call_launches = 0; cudaEventRecord(ctx->ev0, ctx->stream); }
int32_t end_call(Ctx* ctx) {
    PT_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.last_ms = ms; ctx->stats.kernel_launches = ctx->call_launches; ctx->stats.total_launches = ctx->total_launches;
    return 0;
}

int32_t setup_wave(Ctx* ctx) {
    if (ctx->wave_ready) return 0;
    uint32_t W = ctx->cfg.width, H = ctx->cfg.height;
    HostVec<uint32_t> slot_pixel{CbAlloc<uint32_t>(ctx->host_alloc.alloc ? &ctx->host_alloc : nullptr)};
    bool whole = ctx->part_count <= 1;
    if (!whole) {
        uint32_t T = ctx->part_tile ? ctx->part_tile : 32;
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x)
                if (((x / T) + (y / T)) % ctx->part_count == ctx->part_rank) slot_pixel.push_back(y * W + x);
        ctx->num_slots = (uint32_t)slot_pixel.size();
        PT_CK(ctx->w_slot_pixel.alloc((size_t)ctx->num_slots * 4));
        PT_CK(cudaMemcpyAsync(ctx->w_slot_pixel.p, slot_pixel.data(), (size_t)ctx->num_slots * 4, cudaMemcpyHostToDevice, ctx->stream));
        PT_CK(cudaStreamSynchronize(ctx->stream));
    } else { ctx->num_slots = W * H; ctx->w_slot_pixel.release(); }
    // several samples per wave: more rays in flight per launch and fewer launches per sample (results unchanged: one slot per
    // (pixel, sample), accumulated in sample order).  Capped at 64 samples / 32 M slots (5.2 GB of wavefront state): measured on one box at 1080p, 64 spp,
    // 8 bounces (profiles/r02_ab_wave_samples.log): 8 samples per wave 313 / 438 spp/s (configs 3 / 2), 16 -> 321 / 455, 32 -> 326 / 464 (10 GB).
    ctx->wave_samples = 1;
    if (ctx->num_slots) { uint64_t k = (32ull << 20) / ctx->num_slots; ctx->wave_samples = (uint32_t)(k < 1 ? 1 : (k > 64 ? 64 : k)); }
    if (const char* e = getenv("FOUNDATION_PT_WAVE_SAMPLES")) { int v = atoi(e); if (v >= 1 && v <= 64) ctx->wave_samples = v; }
    // FOUNDATION_PT_DUAL_WAVE=0 keeps one copy of the wavefront state (half the memory: 5.2 instead of 10.4 GB at 1080p) and one wave in flight;
    // per-stage timing needs the stages of a render in one stream
    ctx->wave_sets = 2;
    if (const char* e = getenv("FOUNDATION_PT_DUAL_WAVE")) { if (atoi(e) == 0) ctx->wave_sets = 1; }
    if (ctx->cfg.flags & FOUNDATION_PT_FLAG_STAGE_TIMING) ctx->wave_sets = 1;
    if (ctx->wave_sets == 2 && !ctx->stream_b) {
        if (cudaStreamCreateWithFlags(&ctx->stream_b, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_acc[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_acc[1], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ctx->wave_sets = 1; }
    }
    const size_t S1 = (size_t)(ctx->num_slots ? ctx->num_slots : 1) * ctx->wave_samples;
    auto alloc_state = [&](size_t S) -> cudaError_t {
        struct { DevBuf* b; size_t bytes; } bufs[] = {{&ctx->w_ray_o, 16}, {&ctx->w_ray_d, 16}, {&ctx->w_beta, 16}, {&ctx->w_L, 16}, {&ctx->w_rng, 16}, {&ctx->w_hit, 16},
                                                      {&ctx->w_active, 4}, {&ctx->w_next, 4}, {&ctx->w_sorted, 4}, {&ctx->w_sh_o, 16}, {&ctx->w_sh_d, 16}, {&ctx->w_sh_c, 16}};
        for (auto& x : bufs) { cudaError_t e = x.b->alloc(S * x.bytes); if (e != cudaSuccess) return e; }
        if (ctx->attr_enabled) return ctx->w_hit_uv.alloc(S * 8);
        ctx->w_hit_uv.release();
        return cudaSuccess;
    };
    cudaError_t ea = alloc_state(S1 * (size_t)ctx->wave_sets);
    if (ea == cudaErrorMemoryAllocation && ctx->wave_sets == 2) {      // no room for the second copy: one wave at a time
        cudaGetLastError();
        for (DevBuf* b : {&ctx->w_ray_o, &ctx->w_ray_d, &ctx->w_beta, &ctx->w_L, &ctx->w_rng, &ctx->w_hit, &ctx->w_active, &ctx->w_next, &ctx->w_sorted, &ctx->w_sh_o, &ctx->w_sh_d,
                          &ctx->w_sh_c, &ctx->w_hit_uv}) b->release();
        ctx->wave_sets = 1;
        ea = alloc_state(S1);
    }
    PT_CK(ea);
    PT_CK(ctx->w_ctr.alloc(2 * sizeof(PtWaveCounters))); PT_CK(ctx->w_keyhist.alloc(2 * (PT_KEY_BUCKETS + 1) * 4));
    PT_CK(cudaMemsetAsync(ctx->w_ctr.p, 0, 2 * sizeof(PtWaveCounters), ctx->stream));
    if (!ctx->d_accum.p) {
        PT_CK(ctx->d_accum.alloc((size_t)W * H * 16));
        PT_CK(cudaMemsetAsync(ctx->d_accum.p, 0, (size_t)W * H * 16, ctx->stream));
    }
    ctx->wave_ready = true;
    return 0;
}

template <bool TWO>
int32_t render_impl(Ctx* ctx, uint32_t s0, uint32_t ns, uint32_t max_bounces) {
    const size_t S1 = (size_t)ctx->num_slots * ctx->wave_samples;       // slots of one copy of the wavefront state
    PtWave wset[2];
    for (int k = 0; k < ctx->wave_sets; ++k) {
        PtWave& w = wset[k];
        const size_t o = (size_t)k * S1;
        w.ray_o = ctx->w_ray_o.as<float4>() + o; w.ray_d = ctx->w_ray_d.as<float4>() + o; w.beta = ctx->w_beta.as<float4>() + o; w.L = ctx->w_L.as<float4>() + o;
        w.rng = ctx->w_rng.as<uint4>() + o; w.hit = ctx->w_hit.as<float4>() + o; w.active = ctx->w_active.as<uint32_t>() + o; w.next = ctx->w_next.as<uint32_t>() + o;
        w.sorted = ctx->w_sorted.as<uint32_t>() + o; w.sh_o = ctx->w_sh_o.as<float4>() + o; w.sh_d = ctx->w_sh_d.as<float4>() + o; w.sh_c = ctx->w_sh_c.as<float4>() + o;
        w.slot_pixel = ctx->part_count > 1 ? ctx->w_slot_pixel.as<uint32_t>() : nullptr;
        w.ctr = ctx->w_ctr.as<PtWaveCounters>() + k; w.key_hist = ctx->w_keyhist.as<uint32_t>() + (size_t)k * (PT_KEY_BUCKETS + 1);
        w.num_slots = ctx->num_slots; w.num_pixels = ctx->num_slots;
        w.hit_uv = ctx->attr_enabled ? ctx->w_hit_uv.as<float2>() + o : nullptr;
    }
    PtShadeScene ss;
    ss.sv = ctx->view; ss.mats = ctx->d_mats.as<PtMaterial>(); ss.num_mats = (uint32_t)ctx->mats.size();
    ss.mesh_attr = ctx->attr_enabled ? ctx->d_mesh_attr.as<PtMeshAttr>() : nullptr; ss.textures = ctx->d_textures.as<PtTexture>(); ss.mat_tex = ctx->d_mat_tex.as<uint32_t>();
    ss.sc.lights = ctx->d_lights.as<PtLight>(); ss.sc.num_lights = ctx->num_lights; ss.sc.light_area = ctx->light_area; ss.sc.ray_eps = ctx->ray_eps;
    ss.sc.flags = ctx->cfg.flags; ss.sc.seed = ctx->cfg.seed; ss.sc.max_bounces = max_bounces;
    ss.sc.bg[0] = ctx->cfg.background[0]; ss.sc.bg[1] = ctx->cfg.background[1]; ss.sc.bg[2] = ctx->cfg.background[2];
    const bool sort = (ctx->cfg.flags & FOUNDATION_PT_FLAG_MATERIAL_SORT) && !(ctx->cfg.flags & FOUNDATION_PT_FLAG_NO_MATERIAL_SORT);
    uint32_t* status = ctx->d_status.as<uint32_t>();
    // optional per-stage timing: one event per stage boundary, evaluated in foundation_pt_wait (single wave set, see setup_wave)
    const bool timing = (ctx->cfg.flags & FOUNDATION_PT_FLAG_STAGE_TIMING) != 0;
    ctx->stage_marks.clear();
    auto mark = [&](int stage) {
        if (!timing) return;
        const size_t i = ctx->stage_marks.size();
        if (i >= ctx->stage_events.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; } ctx->stage_events.push_back(e); }
        cudaEventRecord(ctx->stage_events[i], ctx->stream);
        ctx->stage_marks.push_back(stage);
    };
    mark(-1);
    // Waves alternate between the two streams / state copies when the call has more than one wave.  What orders them: (1) a stream runs its own waves in order, so a state
    // copy is never reused early; (2) the accumulation of wave i waits for the accumulation of wave i - 1 (event), so every pixel still adds its samples in ascending sample
    // order — the frame stays bit-identical to the one-wave-at-a-time render; (3) the second stream forks from and joins the context's stream around the call.
    // a call that fits one wave (a rank of a multi-GPU frame owns few pixels, so 64 samples of them are one wave) is cut in two halves when both are still large
    uint32_t wave_samples = ctx->wave_samples;
    if (ctx->wave_sets == 2 && !timing && ns >= 2 && ns <= wave_samples && (uint64_t)ctx->num_slots * ns >= (1ull << 23)) wave_samples = (ns + 1) / 2;
    const uint32_t num_waves = (ns + wave_samples - 1) / wave_samples;
    const bool dual = ctx->wave_sets == 2 && num_waves > 1 && !timing;
    if (dual) { PT_CK(cudaEventRecord(ctx->ev_fork, ctx->stream)); PT_CK(cudaStreamWaitEvent(ctx->stream_b, ctx->ev_fork, 0)); }
    uint32_t wave = 0;
    for (uint32_t smp = s0; smp < s0 + ns; ++wave) {
        const int k = dual ? (int)(wave & 1u) : 0;
        cudaStream_t st = k ? ctx->stream_b : ctx->stream;
        PtWave& w = wset[k];
        const uint32_t batch = (s0 + ns - smp) < wave_samples ? (s0 + ns - smp) : wave_samples;
        const uint32_t S = ctx->num_slots * batch;
        w.num_slots = S;
        const uint32_t g256 = grid_for(ctx, S, 256, 8), g128 = grid_for(ctx, S, 128, 8), gtrace = grid_for(ctx, S, 128, trace_blocks(ctx));
        PtFrame f; f.cam = ctx->cam; f.seed = ctx->cfg.seed; f.width = ctx->cfg.width; f.height = ctx->cfg.height; f.sample = smp; f.flags = ctx->cfg.flags;
        smp += batch;
        w.active = ctx->w_active.as<uint32_t>() + (size_t)k * S1; w.next = ctx->w_next.as<uint32_t>() + (size_t)k * S1;
        PT_LAUNCH_ON(ctx, st, k_raygen, g256, 256, w, f);
        mark(0);
        for (uint32_t b = 0; b <= max_bounces; ++b) {
            if (sort) PT_LAUNCH_ON(ctx, st, k_key_clear, 2, 1024, w.key_hist);
            if (w.hit_uv) PT_LAUNCH_ON(ctx, st, (k_extend<TWO, true>), gtrace, 128, ctx->view, w, status, ctx->fetch_thresh);
            else PT_LAUNCH_ON(ctx, st, (k_extend<TWO, false>), gtrace, 128, ctx->view, w, status, ctx->fetch_thresh);
            mark(1);
            const uint32_t* list = w.active;
            if (sort) {
                PT_LAUNCH_ON(ctx, st, k_key_hist, g256, 256, w);
                PT_LAUNCH_ON(ctx, st, k_key_scan, 1, 1024, w.key_hist);
                PT_LAUNCH_ON(ctx, st, k_key_scatter, g256, 256, w);
                list = w.sorted;
                mark(4);
            }
            PT_LAUNCH_ON(ctx, st, k_shade<TWO>, g128, 128, ss, w, list);
            mark(2);
            PT_LAUNCH_ON(ctx, st, k_connect<TWO>, gtrace, 128, ctx->view, w, status, ctx->fetch_thresh);
            PT_LAUNCH_ON(ctx, st, k_bounce_end, 1, 32, w);
            mark(3);
            std::swap(w.active, w.next);
        }
        if (dual && wave > 0) PT_CK(cudaStreamWaitEvent(st, ctx->ev_acc[(wave - 1u) & 1u], 0));
        PT_LAUNCH_ON(ctx, st, k_accumulate, grid_for(ctx, ctx->num_slots, 256, 8), 256, w, ctx->d_accum.as<float4>(),
                     (ctx->comm_count > 1 && (ctx->comm_flags & FOUNDATION_PT_COMM_DIRECT) && ctx->comm_rank != 0) ? ctx->remote_accum : nullptr);
        if (dual) PT_CK(cudaEventRecord(ctx->ev_acc[wave & 1u], st));
        mark(5);
    }
    if (dual) { PT_CK(cudaEventRecord(ctx->ev_join, ctx->stream_b)); PT_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0)); }
    PT_CK(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------------
// C1: NCCL, resolved at run time.  libfoundation_pt.so has no link-time NCCL dependency: a single-GPU host never loads it, and a
// process that already carries an NCCL (e.g. the one bundled with torch) shares that copy (same SONAME).
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr; std::string err;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr; decltype(&ncclCommInitRank) CommInitRank = nullptr; decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr; decltype(&ncclGroupStart) GroupStart = nullptr; decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr; decltype(&ncclRecv) Recv = nullptr; decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr; decltype(&ncclGetErrorString) GetErrorString = nullptr; decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok() const { return lib != nullptr && err.empty(); }
};
const NcclApi& nccl_api() {
    static const NcclApi api = [] {
        NcclApi a;
        const char* names[] = {getenv("FOUNDATION_PT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { if (n && *n && (a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (!a.lib) { a.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return a; }
#define PT_NCCL_SYM(field, sym) do { a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, #sym)); if (!a.field) a.err = "libnccl lacks " #sym; } while (0)
        PT_NCCL_SYM(GetUniqueId, ncclGetUniqueId); PT_NCCL_SYM(CommInitRank, ncclCommInitRank); PT_NCCL_SYM(CommInitAll, ncclCommInitAll);
        PT_NCCL_SYM(CommDestroy, ncclCommDestroy); PT_NCCL_SYM(GroupStart, ncclGroupStart); PT_NCCL_SYM(GroupEnd, ncclGroupEnd);
        PT_NCCL_SYM(Send, ncclSend); PT_NCCL_SYM(Recv, ncclRecv); PT_NCCL_SYM(AllReduce, ncclAllReduce); PT_NCCL_SYM(Broadcast, ncclBroadcast);
        PT_NCCL_SYM(GetErrorString, ncclGetErrorString); PT_NCCL_SYM(GetVersion, ncclGetVersion);
#undef PT_NCCL_SYM
        return a;
    }();
    return api;
}
#define PT_NCCL(expr)                                                                                                      \
    do {                                                                                                                   \
        ncclResult_t r_ = (expr);                                                                                          \
        if (r_ != ncclSuccess) return ctx->fail(FOUNDATION_PT_ERR_COMM, std::string(#expr) + ": " + nccl_api().GetErrorString(r_)); \
    } while (0)

// Pixels every rank owns, and for the root the tables k_unpack_gathered needs: row_base[r][y] = pixels of rank r in the rows above y,
// stage_off[r] = where rank r's block starts in the root's staging buffer.  Same ownership rule as setup_wave.
int32_t plan_gather(Ctx* ctx) {
    const uint32_t W = ctx->cfg.width, H = ctx->cfg.height, T = ctx->part_tile ? ctx->part_tile : 32, N = ctx->part_count;
    const uint32_t tiles_x = (W + T - 1) / T;
    std::vector<uint32_t> row_base((size_t)N * H);
    ctx->rank_pixels.assign(N, 0);
    for (uint32_t y = 0; y < H; ++y) {
        const uint32_t ty = y / T;
        for (uint32_t r = 0; r < N; ++r) row_base[(size_t)r * H + y] = ctx->rank_pixels[r];
        for (uint32_t tx = 0; tx < tiles_x; ++tx) ctx->rank_pixels[(tx + ty) % N] += std::min(T, W - tx * T);
    }
    ctx->stage_off.assign(N, 0);
    uint32_t off = 0;
    for (uint32_t r = 0; r < N; ++r) { ctx->stage_off[r] = off; off += ctx->rank_pixels[r]; }   // the root's own slot stays unused: offsets are root-independent
    PT_CK(ctx->d_row_base.alloc(row_base.size() * 4)); PT_CK(ctx->d_stage_off.alloc((size_t)N * 4));
    PT_CK(cudaMemcpyAsync(ctx->d_row_base.p, row_base.data(), row_base.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaMemcpyAsync(ctx->d_stage_off.p, ctx->stage_off.data(), (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(ctx->d_comm_scratch.alloc(256));
    PT_CK(cudaMemsetAsync(ctx->d_comm_scratch.p, 0, 256, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int32_t ensure_accum(Ctx* ctx) {
    if (ctx->d_accum.p) return 0;
    const size_t need = (size_t)ctx->cfg.width * ctx->cfg.height * 16;
    PT_CK(ctx->d_accum.alloc(need));
    PT_CK(cudaMemsetAsync(ctx->d_accum.p, 0, need, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Enqueues the gather of one frame on the streams of the `n` local members of a communicator (n = 1: one process per GPU; n = count:
// a single-process group).  Stream-ordered behind whatever render is in flight; the caller synchronises.
//   NCCL form:   pack owned pixels -> ncclSend to the root | root: ncclRecv every block -> scatter into the frame.
//   direct form: the owners' accumulate kernels already stored their pixels into the root's frame over NVLink (k_accumulate `remote`),
//                what is left is a barrier: a 4-byte all-reduce orders every rank's stores before the root's next read.
int32_t gather_enqueue(Ctx** cs, uint32_t n, uint32_t root) {
    const NcclApi& api = nccl_api();
    Ctx* ctx = cs[0];
    if (!api.ok()) return ctx->fail(FOUNDATION_PT_ERR_COMM, api.err);
    for (uint32_t i = 0; i < n; ++i) { Ctx* c = cs[i]; if (!c->comm || root >= c->comm_count) return c->fail(FOUNDATION_PT_ERR_STATE, "gather: no communicator (comm_init / group_create first) or bad root"); }
    const bool direct = (ctx->comm_flags & FOUNDATION_PT_COMM_DIRECT) != 0;
    if (direct && root != 0) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "gather: direct mode gathers to rank 0");
    for (uint32_t i = 0; i < n; ++i) {
        ctx = cs[i];
        PT_CK(cudaSetDevice(ctx->device));
        int32_t rc = setup_wave(ctx);          // the owned-pixel list (slot order) lives with the wavefront state
        if (rc) return rc;
        PT_CK(cudaEventRecord(ctx->ev2, ctx->stream));
        if (direct) continue;
        if (ctx->comm_rank != root) {
            PT_CK(ctx->d_pack.bytes >= (size_t)ctx->num_slots * 16 ? cudaSuccess : ctx->d_pack.alloc((size_t)ctx->num_slots * 16 + 16));
            if (ctx->num_slots) PT_LAUNCH(ctx, k_pack_owned, grid_for(ctx, ctx->num_slots, 256, 8), 256, ctx->d_accum.as<float4>(), ctx->w_slot_pixel.as<uint32_t>(), ctx->num_slots, ctx->d_pack.as<float4>());
        } else {
            const size_t total = (size_t)ctx->stage_off.back() + ctx->rank_pixels.back();
            if (ctx->d_stage.bytes < total * 16) PT_CK(ctx->d_stage.alloc(total * 16 + 16));
        }
    }
    ctx = cs[0];
    PT_NCCL(api.GroupStart());
    for (uint32_t i = 0; i < n; ++i) {
        Ctx* c = cs[i];
        cudaSetDevice(c->device);
        ncclResult_t r = ncclSuccess;
        if (direct) r = api.AllReduce(c->d_comm_scratch.p, c->d_comm_scratch.as<float>() + 1, 1, ncclFloat32, ncclSum, c->comm, c->stream);
        else if (c->comm_rank != root) { if (c->num_slots) r = api.Send(c->d_pack.p, (size_t)c->num_slots * 4, ncclFloat32, (int)root, c->comm, c->stream); }
        else
            for (uint32_t p = 0; p < c->comm_count && r == ncclSuccess; ++p)
                if (p != root && c->rank_pixels[p]) r = api.Recv(c->d_stage.as<float4>() + c->stage_off[p], (size_t)c->rank_pixels[p] * 4, ncclFloat32, (int)p, c->comm, c->stream);
        if (r != ncclSuccess) { api.GroupEnd(); return c->fail(FOUNDATION_PT_ERR_COMM, std::string("NCCL send/recv: ") + api.GetErrorString(r)); }
    }
    PT_NCCL(api.GroupEnd());
    for (uint32_t i = 0; i < n; ++i) {
        ctx = cs[i];
        PT_CK(cudaSetDevice(ctx->device));
        if (!direct && ctx->comm_rank == root) {
            PtGatherPlan g; g.width = ctx->cfg.width; g.height = ctx->cfg.height; g.tile = ctx->part_tile ? ctx->part_tile : 32; g.count = ctx->comm_count; g.root = root; g.only = PT_NONE;
            PT_LAUNCH(ctx, k_unpack_gathered, grid_for(ctx, (uint64_t)g.width * g.height, 256, 8), 256, g, ctx->d_stage.as<float4>(), ctx->d_stage_off.as<uint32_t>(),
                      ctx->d_row_base.as<uint32_t>(), ctx->d_accum.as<float4>());
        }
        PT_CK(cudaEventRecord(ctx->ev3, ctx->stream));
    }
    return 0;
}
int32_t gather_finish(Ctx* ctx) {
    PT_CK(cudaSetDevice(ctx->device));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->gather_ms = ms; else cudaGetLastError();
    ctx->stats.gather_ms = ctx->gather_ms;
    return 0;
}
void comm_release(Ctx* ctx) {
    if (ctx->remote_accum && ctx->remote_is_ipc) cudaIpcCloseMemHandle(ctx->remote_accum);
    ctx->remote_accum = nullptr; ctx->remote_is_ipc = false;
    if (ctx->comm && ctx->comm_owned && nccl_api().ok()) nccl_api().CommDestroy(ctx->comm);
    ctx->comm = nullptr; ctx->comm_owned = false; ctx->comm_count = 1; ctx->comm_rank = 0; ctx->comm_flags = 0;
    cudaGetLastError();
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
#define PT_TRY try {
#define PT_CATCH(ctxexpr)                                                                                               \
    } catch (const std::bad_alloc&) { if (ctxexpr) (ctxexpr)->err = "host allocation failed"; return FOUNDATION_PT_ERR_OOM; } \
    catch (const std::exception& e) { if (ctxexpr) (ctxexpr)->err = e.what(); return FOUNDATION_PT_ERR_STATE; }         \
    catch (...) { if (ctxexpr) (ctxexpr)->err = "unknown exception"; return FOUNDATION_PT_ERR_STATE; }

extern "C" int32_t foundation_pt_wait(foundation_pt_context* ctx);
namespace { void settle(Ctx* ctx) { if (ctx->render_pending) foundation_pt_wait(ctx); } }

extern "C" {

uint32_t foundation_pt_version(void) { return 0x00010000u; }

const char* foundation_pt_last_error(const foundation_pt_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t foundation_pt_create(const foundation_pt_config* config, const foundation_pt_allocator* host_alloc, foundation_pt_context** out_ctx) {
    if (!out_ctx) { g_create_error = "out_ctx is NULL"; return FOUNDATION_PT_ERR_ARGUMENT; }
    *out_ctx = nullptr;
    if (!config || config->struct_size != sizeof(foundation_pt_config)) { g_create_error = "config NULL or struct_size mismatch"; return FOUNDATION_PT_ERR_ARGUMENT; }
    if (config->width == 0 || config->height == 0 || (uint64_t)config->width * config->height > (1ull << 30)) { g_create_error = "bad render target size"; return FOUNDATION_PT_ERR_ARGUMENT; }
    if (config->max_leaf_tris > 3) { g_create_error = "max_leaf_tris must be 0..3"; return FOUNDATION_PT_ERR_ARGUMENT; }
    if (host_alloc && ((host_alloc->alloc != nullptr) != (host_alloc->free != nullptr))) {
        g_create_error = "host allocator must provide both alloc and free (or neither)"; return FOUNDATION_PT_ERR_ARGUMENT;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (") + cudaGetErrorString(e) + "); this backend has no CPU fallback";
        cudaGetLastError();
        return FOUNDATION_PT_ERR_NO_DEVICE;
    }
    if (config->device < 0 || config->device >= ndev) { g_create_error = "device ordinal out of range"; return FOUNDATION_PT_ERR_ARGUMENT; }
    foundation_pt_context* ctx = new (std::nothrow) foundation_pt_context();
    if (!ctx) { g_create_error = "host allocation failed"; return FOUNDATION_PT_ERR_OOM; }
    ctx->cfg = *config;
    if (ctx->cfg.max_leaf_tris == 0) ctx->cfg.max_leaf_tris = PT_MAX_LEAF;
    if (host_alloc) ctx->host_alloc = *host_alloc;
    ctx->device = config->device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, ctx->device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev2)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev3)) != cudaSuccess) {
        g_create_error = std::string("CUDA init failed: ") + cudaGetErrorString(e);
        delete ctx;
        return FOUNDATION_PT_ERR_CUDA;
    }
    ctx->num_sms = prop.multiProcessorCount;
    {   // stream-ordered pool for build scratch: keep freed blocks cached instead of returning them to the driver at every sync
        cudaMemPool_t pool; int supported = 0;
        if (!getenv("FOUNDATION_PT_NO_POOL") && cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, ctx->device) == cudaSuccess && supported &&
            cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr) == cudaSuccess) ctx->use_pool = true;
        }
        cudaGetLastError();
    }
    if (const char* e = getenv("FOUNDATION_PT_FETCH_THRESH")) { int v = atoi(e); if (v >= 0 && v <= 32) ctx->fetch_thresh = v; }
    if (const char* e = getenv("FOUNDATION_PT_E2E_CHUNK_LOG2")) { int v = atoi(e); if (v >= 12 && v <= 30) ctx->e2e_chunk_log2 = v; }
    if (const char* e = getenv("FOUNDATION_PT_E2E_TAIL_LOG2")) { int v = atoi(e); if (v >= 12 && v <= 30) ctx->e2e_tail_log2 = v; }
    if (const char* e = getenv("FOUNDATION_PT_TRACE_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 16) ctx->trace_blocks_per_sm = v; }
    ctx->mats.push_back(PtMaterial{0.8f, 0.8f, 0.8f, 0.5f, 0, 0, 0, 0});
    *out_ctx = ctx;
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_destroy(foundation_pt_context* ctx) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    if (ctx->stream3) cudaStreamSynchronize(ctx->stream3);
    if (ctx->stream_b) cudaStreamSynchronize(ctx->stream_b);
    for (cudaEvent_t e : {ctx->ev_fork, ctx->ev_join, ctx->ev_acc[0], ctx->ev_acc[1]}) if (e) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0); if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2); if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    for (cudaEvent_t e : ctx->stage_events) cudaEventDestroy(e);
    cudaStream_t s1 = ctx->stream, s2 = ctx->stream2, s3 = ctx->stream3, s4 = ctx->stream_b;
    delete ctx;   // frees device buffers
    if (s1) cudaStreamDestroy(s1); if (s2) cudaStreamDestroy(s2); if (s3) cudaStreamDestroy(s3); if (s4) cudaStreamDestroy(s4);
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_materials_set(foundation_pt_context* ctx, const foundation_pt_material* materials, uint32_t count) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!materials || count == 0) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "materials NULL or empty");
    PT_TRY
    ctx->mats.resize(count);
    memcpy(ctx->mats.data(), materials, (size_t)count * 32);
    ctx->committed = false;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_mesh_create(foundation_pt_context* ctx, const void* positions, size_t pos_stride_bytes, uint32_t num_vertices, const void* indices,
                                  uint32_t index_format, uint32_t num_triangles, const uint32_t* material_ids, uint32_t* out_mesh_id) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!positions || num_vertices == 0 || num_triangles == 0) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_create: empty mesh");
    if (pos_stride_bytes < 12 || (pos_stride_bytes & 3) || pos_stride_bytes > 4096) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_create: stride must be a multiple of 4, >= 12");
    if (index_format != FOUNDATION_PT_INDEX_U16 && index_format != FOUNDATION_PT_INDEX_U32 && index_format != FOUNDATION_PT_INDEX_NONE)
        return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_create: index_format must be 16, 32 or 0");
    if ((index_format != FOUNDATION_PT_INDEX_NONE) != (indices != nullptr)) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_create: indices / index_format mismatch");
    if (index_format == FOUNDATION_PT_INDEX_NONE && (uint64_t)num_triangles * 3 > num_vertices) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_create: not enough vertices");
    if (num_triangles > 0x7fffffffu / 3) return ctx->fail(FOUNDATION_PT_ERR_UNSUPPORTED, "mesh_create: too many triangles in one mesh");
    PT_TRY
    cudaSetDevice(ctx->device);
    Mesh m(ctx->host_alloc.alloc ? &ctx->host_alloc : nullptr);
    m.stride = (uint32_t)pos_stride_bytes; m.idx_fmt = index_format; m.nverts = num_vertices; m.ntris = num_triangles;
    size_t pos_bytes = (size_t)(num_vertices - 1) * pos_stride_bytes + 12;
    m.h_pos.assign((const uint8_t*)positions, (const uint8_t*)positions + pos_bytes);
    size_t idx_bytes = index_format ? (size_t)num_triangles * 3 * (index_format / 8) : 0;
    if (idx_bytes) m.h_idx.assign((const uint8_t*)indices, (const uint8_t*)indices + idx_bytes);
    // validate indices (the reference trusts its callers; a C ABI should not read out of bounds on the device)
    for (size_t k = 0; k < (size_t)num_triangles * 3 && index_format; ++k) {
        uint32_t v = index_format == 32 ? reinterpret_cast<const uint32_t*>(m.h_idx.data())[k] : reinterpret_cast<const uint16_t*>(m.h_idx.data())[k];
        if (v >= num_vertices) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_create: index out of range");
    }
    m.h_mat.assign(num_triangles, 0u);
    if (material_ids) memcpy(m.h_mat.data(), material_ids, (size_t)num_triangles * 4);
    PT_CK(m.d_pos.alloc(pos_bytes + 16)); PT_CK(m.d_mat.alloc((size_t)num_triangles * 4));
    PT_CK(cudaMemcpyAsync(m.d_pos.p, m.h_pos.data(), pos_bytes, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaMemcpyAsync(m.d_mat.p, m.h_mat.data(), (size_t)num_triangles * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (idx_bytes) { PT_CK(m.d_idx.alloc(idx_bytes)); PT_CK(cudaMemcpyAsync(m.d_idx.p, m.h_idx.data(), idx_bytes, cudaMemcpyHostToDevice, ctx->stream)); }
    PT_CK(cudaStreamSynchronize(ctx->stream));
    ctx->meshes.push_back(std::move(m));
    ctx->flat_valid = false;
    if (out_mesh_id) *out_mesh_id = (uint32_t)ctx->meshes.size() - 1;
    ctx->committed = false;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_mesh_update_positions(foundation_pt_context* ctx, uint32_t mesh_id, const void* positions, size_t pos_stride_bytes, uint32_t num_vertices) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (mesh_id >= ctx->meshes.size()) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_update_positions: mesh_id out of range");
    Mesh& m = ctx->meshes[mesh_id];
    if (!positions || num_vertices != m.nverts || pos_stride_bytes != m.stride)
        return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_update_positions: vertex count and stride must equal those of mesh_create (topology is fixed)");
    PT_TRY
    cudaSetDevice(ctx->device);
    const size_t pos_bytes = (size_t)(num_vertices - 1) * pos_stride_bytes + 12;
    m.h_pos.assign((const uint8_t*)positions, (const uint8_t*)positions + pos_bytes);
    PT_CK(cudaMemcpyAsync(m.d_pos.p, m.h_pos.data(), pos_bytes, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    // a deformed mesh gets a fresh BLAS at the next commit: the full build runs at ~1.9 G triangles/s (5 ms for 10 M), so there is no separate
    // refit path whose tree would degrade with the deformation
    m.d_nodes.release(); m.d_tris.release(); m.d_order.release(); m.num_nodes = 0;
    ctx->flat_valid = false; ctx->committed = false;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_instances_set(foundation_pt_context* ctx, const foundation_pt_instance* instances, uint32_t count) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!instances || count == 0) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "instances NULL or empty");
    PT_TRY
    for (uint32_t i = 0; i < count; ++i)      // singular transforms are detected on the device at scene_commit (ERR_ARGUMENT there)
        if (instances[i].mesh_id >= ctx->meshes.size()) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "instances_set: mesh_id out of range");
    ctx->insts.assign(instances, instances + count);
    ctx->has_insts = true; ctx->committed = false;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_mesh_attributes_set(foundation_pt_context* ctx, uint32_t mesh_id, const void* uv, size_t uv_stride_bytes, const void* color, size_t color_stride_bytes) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (mesh_id >= ctx->meshes.size()) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_attributes_set: mesh_id out of range");
    if ((uv && (uv_stride_bytes < 8 || (uv_stride_bytes & 3))) || (color && (color_stride_bytes < 12 || (color_stride_bytes & 3))))
        return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "mesh_attributes_set: strides must be multiples of 4, >= 8 (uv) / >= 12 (colour)");
    PT_TRY
    cudaSetDevice(ctx->device);
    Mesh& m = ctx->meshes[mesh_id];
    // the streams are compacted on the way in (float2 / float3 per vertex): the caller's layout — e.g. the reference's interleaved,
    // over-aligned vertex_input (Renderer.cpp:23-27) — ends at this call
    HostVec<float> tmp{CbAlloc<float>(ctx->host_alloc.alloc ? &ctx->host_alloc : nullptr)};
    auto upload = [&](const void* src, size_t stride, int comps, DevBuf& dst) -> int32_t {
        if (!src) { dst.release(); return 0; }
        tmp.resize((size_t)m.nverts * comps);
        for (uint32_t v = 0; v < m.nverts; ++v) memcpy(&tmp[(size_t)v * comps], (const uint8_t*)src + (size_t)v * stride, 4 * (size_t)comps);
        PT_CK(dst.alloc(tmp.size() * 4));
        PT_CK(cudaMemcpyAsync(dst.p, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        PT_CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    };
    int32_t rc = upload(uv, uv_stride_bytes, 2, m.d_uv);
    if (!rc) rc = upload(color, color_stride_bytes, 3, m.d_col);
    if (rc) return rc;
    ctx->committed = false;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_texture_create(foundation_pt_context* ctx, const void* rgba8, uint32_t width, uint32_t height, size_t row_pitch_bytes, uint32_t* out_texture_id) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!rgba8 || width == 0 || height == 0 || width > 32768 || height > 32768) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "texture_create: bad image");
    if (row_pitch_bytes == 0) row_pitch_bytes = (size_t)width * 4;
    if (row_pitch_bytes < (size_t)width * 4) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "texture_create: row pitch smaller than a row");
    PT_TRY
    cudaSetDevice(ctx->device);
    foundation_pt_context::Texture t;
    t.width = width; t.height = height;
    PT_CK(t.texels.alloc((size_t)width * height * 4));
    PT_CK(cudaMemcpy2DAsync(t.texels.p, (size_t)width * 4, rgba8, row_pitch_bytes, (size_t)width * 4, height, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    ctx->textures.push_back(std::move(t));
    if (out_texture_id) *out_texture_id = (uint32_t)ctx->textures.size() - 1;
    ctx->committed = false;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_material_textures_set(foundation_pt_context* ctx, const uint32_t* texture_ids, uint32_t count) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!texture_ids && count) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "material_textures_set: NULL ids");
    PT_TRY
    for (uint32_t i = 0; i < count; ++i)
        if (texture_ids[i] != FOUNDATION_PT_NO_TEXTURE && texture_ids[i] >= ctx->textures.size()) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "material_textures_set: texture id out of range");
    ctx->mat_tex.assign(texture_ids, texture_ids + count);
    ctx->committed = false;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_scene_commit(foundation_pt_context* ctx, foundation_pt_build_stats* stats) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (ctx->meshes.empty()) return ctx->fail(FOUNDATION_PT_ERR_STATE, "scene_commit: no meshes");
    if (stats && stats->struct_size != sizeof(foundation_pt_build_stats)) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "build_stats struct_size mismatch");
    PT_TRY
    cudaSetDevice(ctx->device);
    PoolScope pool_scope(ctx->use_pool ? ctx->stream : nullptr);
    ctx->call_launches = 0;
    PT_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    float sort_ms = 0;
    for (auto& m : ctx->meshes) {
        if (m.d_nodes.p) continue;   // BLAS reuse: a mesh is immutable once created
        int32_t rc = build_blas(ctx, m, &sort_ms);
        if (rc) return rc;
    }
    ctx->two_level = ctx->has_insts || ctx->meshes.size() > 1;
    uint64_t total_tris = 0, eff_tris = 0, total_nodes = 0;
    for (auto& m : ctx->meshes) { total_tris += m.ntris; total_nodes += m.num_nodes; }
    ctx->mesh_info.assign(ctx->meshes.size(), PtMeshInfo{});
    if (ctx->two_level) {
        if (!ctx->has_insts) {
            ctx->insts.resize(ctx->meshes.size());
            for (size_t i = 0; i < ctx->meshes.size(); ++i) {
                memset(&ctx->insts[i], 0, sizeof(foundation_pt_instance));
                ctx->insts[i].mesh_id = (uint32_t)i;
                ctx->insts[i].transform[0] = ctx->insts[i].transform[5] = ctx->insts[i].transform[10] = 1.0f;
            }
        }
        uint32_t ni = (uint32_t)ctx->insts.size();
        ctx->num_inst = ni;
        for (uint32_t i = 0; i < ni; ++i) eff_tris += ctx->meshes[ctx->insts[i].mesh_id].ntris;
        for (size_t k = 0; k < ctx->meshes.size(); ++k) {
            memcpy(ctx->mesh_info[k].lo, ctx->meshes[k].lo, 12); memcpy(ctx->mesh_info[k].hi, ctx->meshes[k].hi, 12);
            ctx->mesh_info[k].pad = ctx->meshes[k].pad; ctx->mesh_info[k].ntris = ctx->meshes[k].ntris; ctx->mesh_info[k].nnodes = ctx->meshes[k].num_nodes;
        }
        uint32_t blas_nodes = 0, blas_tris = 0;
        for (size_t k = 0; k < ctx->meshes.size(); ++k) {
            ctx->mesh_info[k].node_base = blas_nodes; ctx->mesh_info[k].tri_base = blas_tris;
            blas_nodes += ctx->meshes[k].num_nodes; blas_tris += ctx->meshes[k].ntris;
        }
        PT_CK(ctx->d_inst_in.alloc((size_t)ni * sizeof(PtInstance))); PT_CK(ctx->d_mesh_info.alloc(ctx->mesh_info.size() * sizeof(PtMeshInfo)));
        {   // instance records (inverse transforms in double) are derived on the device from the raw 64-byte ABI records
            int32_t rcs = ensure_status(ctx);
            if (rcs) return rcs;
            DevBuf d_raw; PT_CK(d_raw.alloc((size_t)ni * sizeof(foundation_pt_instance)));
            PT_CK(cudaMemcpyAsync(d_raw.p, ctx->insts.data(), (size_t)ni * sizeof(foundation_pt_instance), cudaMemcpyHostToDevice, ctx->stream));
            PT_LAUNCH(ctx, k_inst_prepare, grid_for(ctx, ni, 256, 8), 256, d_raw.as<PtInstanceIn>(), ni, ctx->d_inst_in.as<PtInstance>(), ctx->d_status.as<uint32_t>());
            uint32_t st = 0;
            PT_CK(cudaMemcpyAsync(&st, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
            PT_CK(cudaStreamSynchronize(ctx->stream));
            if (st & 2u) { cudaMemsetAsync(ctx->d_status.p, 0, 4, ctx->stream); return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "scene_commit: singular instance transform"); }
        }
        PT_CK(cudaMemcpyAsync(ctx->d_mesh_info.p, ctx->mesh_info.data(), ctx->mesh_info.size() * sizeof(PtMeshInfo), cudaMemcpyHostToDevice, ctx->stream));
        // A6: TLAS = the same LBVH -> BVH8 pipeline over instance world boxes, one instance per leaf slot
        DevBuf prim_box, bounds, bp, keys, vals;
        PT_CK(prim_box.alloc((size_t)ni * sizeof(PtBox))); PT_CK(bounds.alloc(32)); PT_CK(bp.alloc(sizeof(PtBuildParams)));
        PT_CK(keys.alloc((size_t)ni * 8)); PT_CK(vals.alloc((size_t)ni * 4));
        uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        PT_CK(cudaMemcpyAsync(bounds.p, init, 24, cudaMemcpyHostToDevice, ctx->stream));
        PT_LAUNCH(ctx, k_inst_boxes, grid_for(ctx, ni, 256, 8), 256, ctx->d_inst_in.as<PtInstance>(), ni, ctx->d_mesh_info.as<PtMeshInfo>(), prim_box.as<PtBox>(),
                  bounds.as<uint32_t>());
        PT_LAUNCH(ctx, k_build_params, 1, 32, bounds.as<uint32_t>(), bp.as<PtBuildParams>());
        PT_LAUNCH(ctx, k_morton_boxes, grid_for(ctx, ni, 256, 8), 256, prim_box.as<PtBox>(), ni, bp.as<PtBuildParams>(), keys.as<uint64_t>(), vals.as<uint32_t>());
        BuildOut out;
        int32_t rc = build_bvh8(ctx, ni, prim_box.as<PtBox>(), keys, vals, bp.as<PtBuildParams>(), 1, &out);
        if (rc) return rc;
        sort_ms += out.sort_ms;
        ctx->tlas_nodes = out.num_nodes;
        PT_CK(ctx->d_instances.alloc((size_t)ni * sizeof(PtInstance)));
        PT_LAUNCH(ctx, k_write_instances, grid_for(ctx, ni, 256, 8), 256, ctx->d_inst_in.as<PtInstance>(), ni, out.order.as<uint32_t>(), out.leaf_seq.as<uint32_t>(),
                  ctx->d_mesh_info.as<PtMeshInfo>(), ctx->d_instances.as<PtInstance>());
        // flat arrays [BLAS 0 | BLAS 1 | ... | TLAS]: the BLAS part only depends on the meshes, so a commit that follows a mere
        // instances_set (moving objects) rebuilds and rewrites the TLAS tail only
        const uint32_t tlas_cap = ni > 64 ? ni : 64;     // a TLAS over ni single-instance leaves never has more than ni nodes
        const bool reuse = ctx->flat_valid && ctx->flat_meshes == ctx->meshes.size() && ctx->flat_tlas_cap >= out.num_nodes;
        if (!reuse) {
            PT_CK(ctx->d_nodes_all.alloc(((size_t)blas_nodes + tlas_cap) * sizeof(PtNode8))); PT_CK(ctx->d_tris_all.alloc((size_t)blas_tris * sizeof(PtTri)));
            for (size_t k = 0; k < ctx->meshes.size(); ++k) {
                PT_CK(cudaMemcpyAsync(ctx->d_nodes_all.as<PtNode8>() + ctx->mesh_info[k].node_base, ctx->meshes[k].d_nodes.p, (size_t)ctx->meshes[k].num_nodes * sizeof(PtNode8),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
                PT_CK(cudaMemcpyAsync(ctx->d_tris_all.as<PtTri>() + ctx->mesh_info[k].tri_base, ctx->meshes[k].d_tris.p, (size_t)ctx->meshes[k].ntris * sizeof(PtTri),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
            }
            ctx->flat_valid = true; ctx->flat_meshes = ctx->meshes.size(); ctx->flat_tlas_cap = tlas_cap;
        }
        ctx->tlas_only_commit = reuse;
        PT_CK(cudaMemcpyAsync(ctx->d_nodes_all.as<PtNode8>() + blas_nodes, out.nodes.p, (size_t)out.num_nodes * sizeof(PtNode8), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->view.tlas_base = blas_nodes;
        PtBuildParams hbp;
        PT_CK(cudaMemcpyAsync(&hbp, bp.p, sizeof hbp, cudaMemcpyDeviceToHost, ctx->stream));
        PT_CK(cudaStreamSynchronize(ctx->stream));
        memcpy(ctx->wlo, hbp.lo, 12); memcpy(ctx->whi, hbp.hi, 12);
        ctx->d_tlas_order = std::move(out.order);
        ctx->view.nodes = ctx->d_nodes_all.as<PtU4>(); ctx->view.tris = ctx->d_tris_all.as<PtU4>(); ctx->view.instances = ctx->d_instances.as<PtU4>();
        total_nodes += out.num_nodes;
    } else {
        Mesh& m = ctx->meshes[0];
        ctx->num_inst = 1; eff_tris = m.ntris; ctx->tlas_nodes = 0;
        memcpy(ctx->wlo, m.lo, 12); memcpy(ctx->whi, m.hi, 12);
        memcpy(ctx->mesh_info[0].lo, m.lo, 12); memcpy(ctx->mesh_info[0].hi, m.hi, 12);
        ctx->mesh_info[0].pad = m.pad; ctx->mesh_info[0].ntris = m.ntris; ctx->mesh_info[0].nnodes = m.num_nodes;
        ctx->d_nodes_all.release(); ctx->d_tris_all.release(); ctx->d_instances.release(); ctx->flat_valid = false; ctx->view.tlas_base = 0;
        ctx->view.nodes = m.d_nodes.as<PtU4>(); ctx->view.tris = m.d_tris.as<PtU4>(); ctx->view.instances = nullptr;
    }
    // lights — instance order, then input triangle order.  The emissive triangles of each mesh are found on the device
    // (one streaming pass over the material ids); only those few are transformed on the host.
    HostVec<PtLight> lights{CbAlloc<PtLight>(ctx->host_alloc.alloc ? &ctx->host_alloc : nullptr)};
    bool any_emissive = false;
    std::vector<uint8_t> mat_em(ctx->mats.size());
    for (size_t k = 0; k < ctx->mats.size(); ++k) { const PtMaterial& mt = ctx->mats[k]; mat_em[k] = (mt.er > 0 || mt.eg > 0 || mt.eb > 0) ? 1 : 0; any_emissive |= mat_em[k] != 0; }
    if (any_emissive) {
        DevBuf d_em, d_cnt;
        bool uploaded = false;
        for (size_t k = 0; k < ctx->meshes.size(); ++k) {
            Mesh& m = ctx->meshes[k];
            if (m.em_valid && m.em_for == mat_em) continue;              // same mesh (material ids are fixed at mesh_create), same emissive materials
            if (!uploaded) {
                PT_CK(d_em.alloc(mat_em.size())); PT_CK(d_cnt.alloc(16));
                PT_CK(cudaMemcpyAsync(d_em.p, mat_em.data(), mat_em.size(), cudaMemcpyHostToDevice, ctx->stream));
                uploaded = true;
            }
            DevBuf d_list; PT_CK(d_list.alloc((size_t)m.ntris * 4));
            PT_CK(cudaMemsetAsync(d_cnt.p, 0, 4, ctx->stream));
            PT_LAUNCH(ctx, k_emissive_list, grid_for(ctx, m.ntris, 256, 8), 256, m.d_mat.as<uint32_t>(), m.ntris, d_em.as<uint8_t>(), (uint32_t)mat_em.size(), d_list.as<uint32_t>(),
                      d_cnt.as<uint32_t>());
            uint32_t cnt = 0;
            PT_CK(cudaMemcpyAsync(&cnt, d_cnt.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
            PT_CK(cudaStreamSynchronize(ctx->stream));
            m.em_tris.resize(cnt);
            if (cnt) PT_CK(cudaMemcpyAsync(m.em_tris.data(), d_list.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PT_CK(cudaStreamSynchronize(ctx->stream));
            std::sort(m.em_tris.begin(), m.em_tris.end());
            m.em_for = mat_em; m.em_valid = true;
        }
        auto add_mesh_lights = [&](size_t mesh_index, const float* o2w) {
            const Mesh& m = ctx->meshes[mesh_index];
            for (uint32_t t : m.em_tris) {
                uint32_t mid = m.h_mat[t];
                const PtMaterial& mt = ctx->mats[mid < ctx->mats.size() ? mid : 0];
                float v[9]; m.host_tri(t, v);
                pt_v3 v0 = pt_mk(v[0], v[1], v[2]), e1 = pt_mk(v[3] - v[0], v[4] - v[1], v[5] - v[2]), e2 = pt_mk(v[6] - v[0], v[7] - v[1], v[8] - v[2]);
                if (o2w) { v0 = pt_xform_point(o2w, v0); e1 = pt_xform_vec(o2w, e1); e2 = pt_xform_vec(o2w, e2); }
                PtLight l; pt_light_make(&l, v0, e1, e2, mt.er, mt.eg, mt.eb);
                lights.push_back(l);
            }
        };
        if (ctx->two_level) {
            for (uint32_t i = 0; i < ctx->num_inst; ++i)
                if (!ctx->meshes[ctx->insts[i].mesh_id].em_tris.empty()) add_mesh_lights(ctx->insts[i].mesh_id, ctx->insts[i].transform);
        }
        else add_mesh_lights(0, nullptr);
    }
    ctx->num_lights = (uint32_t)lights.size();
    ctx->light_area = pt_lights_finalize(lights.data(), &ctx->num_lights);   // degenerate emitters are dropped here
    lights.resize(ctx->num_lights);
    PT_CK(ctx->d_lights.alloc(lights.size() * sizeof(PtLight)));
    if (!lights.empty()) PT_CK(cudaMemcpyAsync(ctx->d_lights.p, lights.data(), lights.size() * sizeof(PtLight), cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(ctx->d_mats.alloc(ctx->mats.size() * sizeof(PtMaterial)));
    PT_CK(cudaMemcpyAsync(ctx->d_mats.p, ctx->mats.data(), ctx->mats.size() * sizeof(PtMaterial), cudaMemcpyHostToDevice, ctx->stream));
    {   // attribute tables: per mesh {uv, colour, indices}, the texture descriptors, the per-material texture ids
        bool any = false;
        std::vector<PtMeshAttr> ma(ctx->meshes.size());
        for (size_t k = 0; k < ctx->meshes.size(); ++k) {
            const Mesh& m = ctx->meshes[k];
            ma[k].uv = m.d_uv.as<uint8_t>(); ma[k].col = m.d_col.as<uint8_t>(); ma[k].idx = m.d_idx.p; ma[k].uv_stride = 8; ma[k].col_stride = 12; ma[k].idx_fmt = m.idx_fmt; ma[k].pad = 0;
            any |= m.d_uv.p != nullptr || m.d_col.p != nullptr;
        }
        const bool was = ctx->attr_enabled;
        ctx->attr_enabled = any;
        if (any != was) ctx->wave_ready = false;                  // the wavefront state grows / loses the per-slot barycentrics
        if (any) {
            std::vector<PtTexture> td(ctx->textures.size() ? ctx->textures.size() : 1);
            for (size_t k = 0; k < ctx->textures.size(); ++k) { td[k].texels = ctx->textures[k].texels.as<uint32_t>(); td[k].width = ctx->textures[k].width; td[k].height = ctx->textures[k].height; td[k].pad = 0; }
            std::vector<uint32_t> mt(ctx->mats.size(), PT_NONE);
            for (size_t k = 0; k < mt.size() && k < ctx->mat_tex.size(); ++k) mt[k] = ctx->mat_tex[k] < ctx->textures.size() ? ctx->mat_tex[k] : PT_NONE;
            PT_CK(ctx->d_mesh_attr.alloc(ma.size() * sizeof(PtMeshAttr))); PT_CK(ctx->d_textures.alloc(td.size() * sizeof(PtTexture))); PT_CK(ctx->d_mat_tex.alloc(mt.size() * 4));
            PT_CK(cudaMemcpyAsync(ctx->d_mesh_attr.p, ma.data(), ma.size() * sizeof(PtMeshAttr), cudaMemcpyHostToDevice, ctx->stream));
            PT_CK(cudaMemcpyAsync(ctx->d_textures.p, td.data(), td.size() * sizeof(PtTexture), cudaMemcpyHostToDevice, ctx->stream));
            PT_CK(cudaMemcpyAsync(ctx->d_mat_tex.p, mt.data(), mt.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            PT_CK(cudaStreamSynchronize(ctx->stream));           // ma / td / mt are locals
        }
    }
    float ext = 0;
    for (int k = 0; k < 3; ++k) ext = pt_max(ext, ctx->whi[k] - ctx->wlo[k]);
    ctx->ray_eps = ext * PT_RAY_EPS_REL;
    int32_t rc = ensure_status(ctx);
    if (rc) return rc;
    pin_nodes_in_l2(ctx, total_nodes * sizeof(PtNode8));
    PT_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    foundation_pt_build_stats& bs = ctx->bstats;
    bs.struct_size = sizeof bs; bs.num_meshes = (uint32_t)ctx->meshes.size(); bs.num_instances = ctx->num_inst; bs.num_triangles = total_tris;
    bs.effective_triangles = eff_tris; bs.num_nodes8 = total_nodes; bs.device_bytes = total_nodes * sizeof(PtNode8) + total_tris * sizeof(PtTri);
    bs.build_ms = ms; bs.sort_ms = sort_ms;
    memcpy(bs.scene_lo, ctx->wlo, 12); memcpy(bs.scene_hi, ctx->whi, 12);
    if (stats) *stats = bs;
    ctx->stats.kernel_launches = ctx->call_launches; ctx->stats.last_ms = ms; ctx->stats.total_launches = ctx->total_launches;
    ctx->committed = true;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_camera_set(foundation_pt_context* ctx, const float view[16], const float proj[16]) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!view || !proj) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "camera_set: NULL matrix");
    if (!pt_camera_derive(view, proj, &ctx->cam)) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "camera_set: singular view/projection");
    ctx->cam_set = true;
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_partition_set(foundation_pt_context* ctx, uint32_t rank, uint32_t count, uint32_t tile_size) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (count == 0 || rank >= count) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "partition_set: rank must be < count");
    if (ctx->comm && (rank != ctx->comm_rank || count != ctx->comm_count)) return ctx->fail(FOUNDATION_PT_ERR_STATE, "partition_set: the partition of a communicator member is fixed by comm_init / group_create");
    ctx->part_rank = rank; ctx->part_count = count; ctx->part_tile = tile_size ? tile_size : 32;
    ctx->wave_ready = false;
    return FOUNDATION_PT_OK;
}

// render = render_async + wait.  The asynchronous pair only enqueues the wavefront loop on the context's stream (no host round trip
// happens inside the loop: the bounce count is fixed and the queue sizes live on the device), so a host Draw() can overlap its own work.
int32_t foundation_pt_render_async(foundation_pt_context* ctx, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    if (!ctx->committed) return ctx->fail(FOUNDATION_PT_ERR_STATE, "render: scene not committed");
    if (!ctx->cam_set) return ctx->fail(FOUNDATION_PT_ERR_STATE, "render: camera not set");
    if (max_bounces > 64) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "render: max_bounces > 64");
    // the PCG stream id is (sample << 32 | pixel) << 1 | 1: bit 31 of the sample index would be shifted out and alias an earlier sample,
    // and a 32-bit sample_begin + sample_count must not wrap
    if ((uint64_t)sample_begin + sample_count > (1ull << 31)) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "render: sample_begin + sample_count must not exceed 2^31");
    if (ctx->render_pending) return ctx->fail(FOUNDATION_PT_ERR_STATE, "render_async: the previous asynchronous render has not been waited for");
    PT_TRY
    cudaSetDevice(ctx->device);
    int32_t rc = setup_wave(ctx);
    if (rc) return rc;
    begin_call(ctx);
    if (sample_begin == 0) {
        // direct mode: the pixels of the root's frame that belong to other ranks are written by THEIR accumulate kernels over NVLink,
        // possibly before this stream gets here — every rank clears only the pixels it owns
        if (ctx->comm_count > 1 && (ctx->comm_flags & FOUNDATION_PT_COMM_DIRECT)) {
            if (ctx->num_slots) PT_LAUNCH(ctx, k_clear_owned, grid_for(ctx, ctx->num_slots, 256, 8), 256, ctx->d_accum.as<float4>(), ctx->w_slot_pixel.as<uint32_t>(), ctx->num_slots);
        } else PT_CK(cudaMemsetAsync(ctx->d_accum.p, 0, (size_t)ctx->cfg.width * ctx->cfg.height * 16, ctx->stream));
    }
    PT_CK(cudaMemsetAsync(ctx->w_ctr.p, 0, 2 * sizeof(PtWaveCounters), ctx->stream));
    if (ctx->num_slots) {
        rc = ctx->two_level ? render_impl<true>(ctx, sample_begin, sample_count, max_bounces) : render_impl<false>(ctx, sample_begin, sample_count, max_bounces);
        if (rc) return rc;
    }
    PT_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->render_pending = true;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_wait(foundation_pt_context* ctx) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    if (!ctx->render_pending) return FOUNDATION_PT_OK;      // nothing in flight: every other entry point is blocking
    PT_TRY
    cudaSetDevice(ctx->device);
    ctx->render_pending = false;
    PT_CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.last_ms = ms; ctx->stats.kernel_launches = ctx->call_launches; ctx->stats.total_launches = ctx->total_launches;
    PtWaveCounters c[2];
    PT_CK(cudaMemcpy(c, ctx->w_ctr.p, sizeof c, cudaMemcpyDeviceToHost));
    ctx->stats.rays_extend = c[0].total_extend + c[1].total_extend; ctx->stats.rays_shadow = c[0].total_shadow + c[1].total_shadow;
    for (float& v : ctx->stats.stage_ms) v = 0.0f;
    for (size_t i = 1; i < ctx->stage_marks.size(); ++i) {
        float dt = 0.0f;
        if (ctx->stage_marks[i] >= 0 && cudaEventElapsedTime(&dt, ctx->stage_events[i - 1], ctx->stage_events[i]) == cudaSuccess) ctx->stats.stage_ms[ctx->stage_marks[i]] += dt;
    }
    cudaGetLastError();
    return check_status(ctx);
    PT_CATCH(ctx)
}

int32_t foundation_pt_render(foundation_pt_context* ctx, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces) {
    int32_t rc = foundation_pt_render_async(ctx, sample_begin, sample_count, max_bounces);
    if (rc) return rc;
    return foundation_pt_wait(ctx);
}

int32_t foundation_pt_read_accum(foundation_pt_context* ctx, float* rgba, size_t size_bytes) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    size_t need = (size_t)ctx->cfg.width * ctx->cfg.height * 16;
    if (!rgba || size_bytes < need) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "read_accum: buffer too small");
    cudaSetDevice(ctx->device);
    if (!ctx->d_accum.p) { memset(rgba, 0, need); return FOUNDATION_PT_OK; }
    PT_CK(cudaMemcpyAsync(rgba, ctx->d_accum.p, need, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_write_accum(foundation_pt_context* ctx, const float* rgba, size_t size_bytes) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    size_t need = (size_t)ctx->cfg.width * ctx->cfg.height * 16;
    if (!rgba || size_bytes < need) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "write_accum: buffer too small");
    cudaSetDevice(ctx->device);
    if (!ctx->d_accum.p) PT_CK(ctx->d_accum.alloc(need));
    PT_CK(cudaMemcpyAsync(ctx->d_accum.p, rgba, need, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_resolve_rgba8(foundation_pt_context* ctx, uint8_t* rgba8, size_t size_bytes) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    uint32_t n = ctx->cfg.width * ctx->cfg.height;
    if (!rgba8 || size_bytes < (size_t)n * 4) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "resolve_rgba8: buffer too small");
    if (!ctx->d_accum.p) return ctx->fail(FOUNDATION_PT_ERR_STATE, "resolve_rgba8: nothing rendered");
    PT_TRY
    cudaSetDevice(ctx->device);
    DevBuf tmp; PT_CK(tmp.alloc((size_t)n * 4));
    PT_LAUNCH(ctx, k_resolve_rgba8, grid_for(ctx, n, 256, 8), 256, ctx->d_accum.as<float4>(), n, tmp.as<uint32_t>());
    PT_CK(cudaMemcpyAsync(rgba8, tmp.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_accum_device_ptr(foundation_pt_context* ctx, void** out_device_ptr, size_t* out_size_bytes) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!out_device_ptr) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "accum_device_ptr: NULL out pointer");
    cudaSetDevice(ctx->device);
    size_t need = (size_t)ctx->cfg.width * ctx->cfg.height * 16;
    if (!ctx->d_accum.p) { PT_CK(ctx->d_accum.alloc(need)); PT_CK(cudaMemset(ctx->d_accum.p, 0, need)); }
    *out_device_ptr = ctx->d_accum.p;
    if (out_size_bytes) *out_size_bytes = need;
    return FOUNDATION_PT_OK;
}

// ---- explicit ray sets --------------------------------------------------------------------------------
int32_t foundation_pt_rays_upload(foundation_pt_context* ctx, const foundation_pt_ray* rays, uint64_t count) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!rays && count) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "rays_upload: NULL rays");
    PT_TRY
    cudaSetDevice(ctx->device);
    if (ctx->d_rays.bytes < count * 32) {
        PT_CK(ctx->d_rays.alloc(count * 32)); PT_CK(ctx->d_hits.alloc(count * 16)); PT_CK(ctx->d_hit_inst.alloc(count * 4)); PT_CK(ctx->d_occ.alloc(count));
    }
    ctx->num_rays = count;
    if (count) PT_CK(cudaMemcpyAsync(ctx->d_rays.p, rays, count * 32, cudaMemcpyHostToDevice, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

static int32_t rays_trace_common(foundation_pt_context* ctx, uint64_t first, uint64_t count, int mode) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    if (!ctx->committed) return ctx->fail(FOUNDATION_PT_ERR_STATE, "trace: scene not committed");
    if (first + count > ctx->num_rays) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "trace: range exceeds the uploaded ray set");
    PT_TRY
    cudaSetDevice(ctx->device);
    begin_call(ctx);
    const float4* rays = ctx->d_rays.as<float4>() + 2 * first;
    int32_t rc = 0;
    if (mode == 0) rc = launch_trace<false>(ctx, rays, count, ctx->d_hits.as<float4>() + first, ctx->d_hit_inst.as<uint32_t>() + first, nullptr);
    else if (mode == 1) rc = launch_trace<true>(ctx, rays, count, nullptr, nullptr, ctx->d_occ.as<uint8_t>() + first);
    else if (count) {
        uint32_t grid = (uint32_t)((count + 127) / 128);
        if (ctx->two_level)
            PT_LAUNCH(ctx, k_trace_brute<true>, grid, 128, ctx->view, ctx->d_inst_in.as<PtInstance>(), ctx->num_inst, ctx->d_mesh_info.as<PtMeshInfo>(), 0u, rays,
                      (unsigned long long)count, ctx->d_hits.as<float4>() + first, ctx->d_hit_inst.as<uint32_t>() + first);
        else
            PT_LAUNCH(ctx, k_trace_brute<false>, grid, 128, ctx->view, nullptr, 1u, nullptr, ctx->meshes[0].ntris, rays, (unsigned long long)count,
                      ctx->d_hits.as<float4>() + first, ctx->d_hit_inst.as<uint32_t>() + first);
        PT_CK(cudaGetLastError());
    }
    if (rc) return rc;
    rc = end_call(ctx);
    if (rc) return rc;
    ctx->stats.trace_ms = ctx->stats.last_ms;
    return check_status(ctx);
    PT_CATCH(ctx)
}
int32_t foundation_pt_rays_trace_closest(foundation_pt_context* ctx, uint64_t first, uint64_t count) { return rays_trace_common(ctx, first, count, 0); }
int32_t foundation_pt_rays_trace_any(foundation_pt_context* ctx, uint64_t first, uint64_t count) { return rays_trace_common(ctx, first, count, 1); }
int32_t foundation_pt_rays_trace_brute(foundation_pt_context* ctx, uint64_t first, uint64_t count) { return rays_trace_common(ctx, first, count, 2); }

int32_t foundation_pt_rays_download_hits(foundation_pt_context* ctx, uint64_t first, uint64_t count, foundation_pt_hit* out_hits, uint32_t* out_inst) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (first + count > ctx->num_rays) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "download_hits: range exceeds the uploaded ray set");
    cudaSetDevice(ctx->device);
    if (out_hits && count) PT_CK(cudaMemcpyAsync(out_hits, ctx->d_hits.as<float4>() + first, count * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_inst && count) PT_CK(cudaMemcpyAsync(out_inst, ctx->d_hit_inst.as<uint32_t>() + first, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PT_CK(cudaStreamSynchronize(ctx->stream));
    return FOUNDATION_PT_OK;
}

// Host-buffer variants: H2D + kernel + D2H inside the call, chunked so copies of one chunk overlap the
// traversal of another when the caller's buffers are pinned (closest-hit and any-hit share the pipeline).
}  // extern "C"
namespace {
template <bool ANY>
int32_t trace_host_pipelined(Ctx* ctx, const foundation_pt_ray* rays, uint64_t count, foundation_pt_hit* out_hits, uint32_t* out_inst, uint8_t* out_occ) {
    PT_TRY
    cudaSetDevice(ctx->device);
    if (ctx->d_rays.bytes < count * 32) {
        PT_CK(ctx->d_rays.alloc(count * 32)); PT_CK(ctx->d_hits.alloc(count * 16)); PT_CK(ctx->d_hit_inst.alloc(count * 4)); PT_CK(ctx->d_occ.alloc(count));
    }
    ctx->num_rays = count;
    begin_call(ctx);
    cudaEvent_t done_h2d = ctx->ev2, done_k = ctx->ev3;
    // 2^22-ray chunks (128 MB up, 64 MB down) while the upload is the bottleneck; the last 2^23 rays go in 2^20-ray chunks so that the part of the pipeline
    // nothing overlaps — the last chunk's traversal and download — is a quarter as long
    const uint64_t big = 1ull << ctx->e2e_chunk_log2, small = 1ull << (ctx->e2e_tail_log2 < ctx->e2e_chunk_log2 ? ctx->e2e_tail_log2 : ctx->e2e_chunk_log2);
    for (uint64_t b = 0, chunk; b < count; b += chunk) {
        chunk = count - b > 2 * big ? big : small;
        uint64_t n = count - b < chunk ? count - b : chunk;
        PT_CK(cudaMemcpyAsync(ctx->d_rays.as<uint8_t>() + b * 32, rays + b, n * 32, cudaMemcpyHostToDevice, ctx->stream2));
        PT_CK(cudaEventRecord(done_h2d, ctx->stream2));
        PT_CK(cudaStreamWaitEvent(ctx->stream, done_h2d, 0));
        int32_t rc = ANY ? launch_trace<true>(ctx, ctx->d_rays.as<float4>() + 2 * b, n, nullptr, nullptr, ctx->d_occ.as<uint8_t>() + b)
                         : launch_trace<false>(ctx, ctx->d_rays.as<float4>() + 2 * b, n, ctx->d_hits.as<float4>() + b, ctx->d_hit_inst.as<uint32_t>() + b, nullptr);
        if (rc) return rc;
        PT_CK(cudaEventRecord(done_k, ctx->stream));
        PT_CK(cudaStreamWaitEvent(ctx->stream3, done_k, 0));
        if (ANY) PT_CK(cudaMemcpyAsync(out_occ + b, ctx->d_occ.as<uint8_t>() + b, n, cudaMemcpyDeviceToHost, ctx->stream3));
        else {
            PT_CK(cudaMemcpyAsync(out_hits + b, ctx->d_hits.as<float4>() + b, n * 16, cudaMemcpyDeviceToHost, ctx->stream3));
            if (out_inst) PT_CK(cudaMemcpyAsync(out_inst + b, ctx->d_hit_inst.as<uint32_t>() + b, n * 4, cudaMemcpyDeviceToHost, ctx->stream3));
        }
    }
    PT_CK(cudaStreamSynchronize(ctx->stream2));
    PT_CK(cudaStreamSynchronize(ctx->stream3));
    int32_t rc = end_call(ctx);
    if (rc) return rc;
    return check_status(ctx);
    PT_CATCH(ctx)
}
}  // namespace
extern "C" {

int32_t foundation_pt_trace_closest(foundation_pt_context* ctx, const foundation_pt_ray* rays, uint64_t count, foundation_pt_hit* out_hits, uint32_t* out_inst) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    if ((!rays || !out_hits) && count) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "trace_closest: NULL buffer");
    if (!ctx->committed) return ctx->fail(FOUNDATION_PT_ERR_STATE, "trace: scene not committed");
    return trace_host_pipelined<false>(ctx, rays, count, out_hits, out_inst, nullptr);
}

int32_t foundation_pt_trace_any(foundation_pt_context* ctx, const foundation_pt_ray* rays, uint64_t count, uint8_t* out_occluded) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    if ((!rays || !out_occluded) && count) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "trace_any: NULL buffer");
    if (!ctx->committed) return ctx->fail(FOUNDATION_PT_ERR_STATE, "trace: scene not committed");
    return trace_host_pipelined<true>(ctx, rays, count, nullptr, nullptr, out_occluded);
}

int32_t foundation_pt_stats_get(foundation_pt_context* ctx, foundation_pt_stats* stats) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!stats || stats->struct_size != sizeof(foundation_pt_stats)) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "stats_get: NULL or struct_size mismatch");
    ctx->stats.struct_size = sizeof(foundation_pt_stats);
    ctx->stats.total_launches = ctx->total_launches;
    *stats = ctx->stats;
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_blas_download(foundation_pt_context* ctx, uint32_t mesh_id, void* nodes, size_t nodes_bytes, void* tris, size_t tris_bytes, uint32_t* order,
                                    size_t order_bytes, uint64_t* out_num_nodes, uint64_t* out_num_tris) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (mesh_id >= ctx->meshes.size()) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "blas_download: mesh_id out of range");
    Mesh& m = ctx->meshes[mesh_id];
    if (!m.d_nodes.p) return ctx->fail(FOUNDATION_PT_ERR_STATE, "blas_download: scene not committed");
    cudaSetDevice(ctx->device);
    if (out_num_nodes) *out_num_nodes = m.num_nodes;
    if (out_num_tris) *out_num_tris = m.ntris;
    if (nodes) { if (nodes_bytes < (size_t)m.num_nodes * 80) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "blas_download: nodes buffer too small"); PT_CK(cudaMemcpy(nodes, m.d_nodes.p, (size_t)m.num_nodes * 80, cudaMemcpyDeviceToHost)); }
    if (tris) { if (tris_bytes < (size_t)m.ntris * 48) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "blas_download: tris buffer too small"); PT_CK(cudaMemcpy(tris, m.d_tris.p, (size_t)m.ntris * 48, cudaMemcpyDeviceToHost)); }
    if (order) { if (order_bytes < (size_t)m.ntris * 4) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "blas_download: order buffer too small"); PT_CK(cudaMemcpy(order, m.d_order.p, (size_t)m.ntris * 4, cudaMemcpyDeviceToHost)); }
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_tlas_download(foundation_pt_context* ctx, void* nodes, size_t nodes_bytes, uint32_t* order, size_t order_bytes, uint64_t* out_num_nodes,
                                    uint64_t* out_num_instances) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!ctx->committed) return ctx->fail(FOUNDATION_PT_ERR_STATE, "tlas_download: scene not committed");
    cudaSetDevice(ctx->device);
    if (out_num_nodes) *out_num_nodes = ctx->tlas_nodes;
    if (out_num_instances) *out_num_instances = ctx->two_level ? ctx->num_inst : 0;
    if (!ctx->two_level) return FOUNDATION_PT_OK;
    if (nodes) { if (nodes_bytes < (size_t)ctx->tlas_nodes * 80) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "tlas_download: nodes buffer too small"); PT_CK(cudaMemcpy(nodes, ctx->d_nodes_all.as<PtNode8>() + ctx->view.tlas_base, (size_t)ctx->tlas_nodes * 80, cudaMemcpyDeviceToHost)); }
    if (order) { if (order_bytes < (size_t)ctx->num_inst * 4) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "tlas_download: order buffer too small"); PT_CK(cudaMemcpy(order, ctx->d_tlas_order.p, (size_t)ctx->num_inst * 4, cudaMemcpyDeviceToHost)); }
    return FOUNDATION_PT_OK;
}


// ---- multi-GPU frame (stage C1) ---------------------------------------------------------------------------
int32_t foundation_pt_comm_unique_id(uint8_t* id, size_t size_bytes) {
    if (!id || size_bytes < FOUNDATION_PT_COMM_ID_BYTES) { g_create_error = "comm_unique_id: buffer smaller than FOUNDATION_PT_COMM_ID_BYTES"; return FOUNDATION_PT_ERR_ARGUMENT; }
    const NcclApi& api = nccl_api();
    if (!api.ok()) { g_create_error = api.err; return FOUNDATION_PT_ERR_COMM; }
    static_assert(sizeof(ncclUniqueId) == FOUNDATION_PT_COMM_ID_BYTES, "NCCL unique id size");
    ncclUniqueId u;
    ncclResult_t r = api.GetUniqueId(&u);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + api.GetErrorString(r); return FOUNDATION_PT_ERR_COMM; }
    memcpy(id, &u, sizeof u);
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_comm_init(foundation_pt_context* ctx, const uint8_t* id, size_t size_bytes, uint32_t rank, uint32_t count, uint32_t tile_size, uint32_t flags) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    if (!id || size_bytes < FOUNDATION_PT_COMM_ID_BYTES || count == 0 || rank >= count) return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "comm_init: bad id / rank / count");
    if (ctx->comm) return ctx->fail(FOUNDATION_PT_ERR_STATE, "comm_init: this context already belongs to a communicator");
    const NcclApi& api = nccl_api();
    if (!api.ok()) return ctx->fail(FOUNDATION_PT_ERR_COMM, api.err);
    PT_TRY
    PT_CK(cudaSetDevice(ctx->device));
    ncclUniqueId u; memcpy(&u, id, sizeof u);
    PT_NCCL(api.CommInitRank(&ctx->comm, (int)count, u, (int)rank));
    ctx->comm_owned = true; ctx->comm_rank = rank; ctx->comm_count = count; ctx->comm_flags = flags;
    ctx->part_rank = rank; ctx->part_count = count; ctx->part_tile = tile_size ? tile_size : 32; ctx->wave_ready = false;
    int32_t rc = ensure_accum(ctx);
    if (!rc) rc = plan_gather(ctx);
    if (rc) { comm_release(ctx); return rc; }
    if ((flags & FOUNDATION_PT_COMM_DIRECT) && count > 1) {
        // the root's frame, mapped into every other rank's address space: IPC handle broadcast over the communicator itself
        cudaIpcMemHandle_t h; memset(&h, 0, sizeof h);
        static_assert(sizeof(cudaIpcMemHandle_t) <= 128, "IPC handle fits the scratch buffer");
        if (rank == 0) {
            cudaError_t e = cudaIpcGetMemHandle(&h, ctx->d_accum.p);
            if (e != cudaSuccess) { cudaGetLastError(); memset(&h, 0, sizeof h); }   // all-zero handle = "no IPC": every rank then fails alike instead of hanging
            PT_CK(cudaMemcpyAsync(ctx->d_comm_scratch.as<uint8_t>() + 64, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
        }
        PT_NCCL(api.Broadcast(ctx->d_comm_scratch.as<uint8_t>() + 64, ctx->d_comm_scratch.as<uint8_t>() + 64, sizeof h, ncclUint8, 0, ctx->comm, ctx->stream));
        PT_CK(cudaMemcpyAsync(&h, ctx->d_comm_scratch.as<uint8_t>() + 64, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        PT_CK(cudaStreamSynchronize(ctx->stream));
        bool zero = true; for (size_t k = 0; k < sizeof h; ++k) zero &= reinterpret_cast<const uint8_t*>(&h)[k] == 0;
        if (zero) { comm_release(ctx); return ctx->fail(FOUNDATION_PT_ERR_COMM, "comm_init: the root could not export its frame (cudaIpcGetMemHandle)"); }
        if (rank != 0) {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); comm_release(ctx); return ctx->fail(FOUNDATION_PT_ERR_COMM, std::string("comm_init: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
            ctx->remote_accum = static_cast<float4*>(p); ctx->remote_is_ipc = true;
        }
    }
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

int32_t foundation_pt_gather(foundation_pt_context* ctx, uint32_t root) {
    if (!ctx) return FOUNDATION_PT_ERR_ARGUMENT;
    settle(ctx);
    PT_TRY
    if (ctx->comm_count <= 1 && ctx->comm) return FOUNDATION_PT_OK;
    ctx->call_launches = 0;
    int32_t rc = gather_enqueue(&ctx, 1, root);
    if (rc) return rc;
    rc = gather_finish(ctx);
    ctx->stats.kernel_launches = ctx->call_launches; ctx->stats.total_launches = ctx->total_launches;
    return rc;
    PT_CATCH(ctx)
}


// Same-device form of the gather: several partitions of one frame rendered by several contexts on ONE device (tests on a single-GPU
// box; also a way to split a frame over streams).  Scatters src's owned tiles into dst's frame with the same pack / scatter kernels
// as the NCCL form, a device-to-device copy standing in for ncclSend / ncclRecv.
int32_t foundation_pt_gather_local(foundation_pt_context* dst, foundation_pt_context* src) {
    if (!dst || !src) return FOUNDATION_PT_ERR_ARGUMENT;
    foundation_pt_context* ctx = dst;
    settle(dst); settle(src);
    if (dst == src || dst->device != src->device || dst->part_count != src->part_count || dst->part_tile != src->part_tile || dst->cfg.width != src->cfg.width ||
        dst->cfg.height != src->cfg.height || dst->part_rank == src->part_rank || dst->part_count < 2)
        return ctx->fail(FOUNDATION_PT_ERR_ARGUMENT, "gather_local: the two contexts must be different partitions (same count / tile / frame size) on the same device");
    PT_TRY
    PT_CK(cudaSetDevice(dst->device));
    int32_t rc = setup_wave(src);
    if (rc) { dst->err = src->err; return rc; }
    if (dst->rank_pixels.size() != dst->part_count) { rc = ensure_accum(dst); if (!rc) rc = plan_gather(dst); if (rc) return rc; }
    dst->call_launches = 0;
    const size_t total = (size_t)dst->stage_off.back() + dst->rank_pixels.back();
    if (dst->d_stage.bytes < total * 16) PT_CK(dst->d_stage.alloc(total * 16 + 16));
    if (src->d_pack.bytes < (size_t)src->num_slots * 16) { ctx = src; PT_CK(src->d_pack.alloc((size_t)src->num_slots * 16 + 16)); ctx = dst; }
    if (src->num_slots) {
        k_pack_owned<<<grid_for(src, src->num_slots, 256, 8), 256, 0, src->stream>>>(src->d_accum.as<float4>(), src->w_slot_pixel.as<uint32_t>(), src->num_slots, src->d_pack.as<float4>());
        dst->call_launches++; dst->total_launches++;
        PT_CK(cudaMemcpyAsync(dst->d_stage.as<float4>() + dst->stage_off[src->part_rank], src->d_pack.p, (size_t)src->num_slots * 16, cudaMemcpyDeviceToDevice, src->stream));
    }
    PT_CK(cudaStreamSynchronize(src->stream));
    if (src->num_slots != dst->rank_pixels[src->part_rank]) return ctx->fail(FOUNDATION_PT_ERR_STATE, "gather_local: partition plans disagree");
    PtGatherPlan g; g.width = dst->cfg.width; g.height = dst->cfg.height; g.tile = dst->part_tile ? dst->part_tile : 32; g.count = dst->part_count; g.root = dst->part_rank; g.only = src->part_rank;
    PT_LAUNCH(dst, k_unpack_gathered, grid_for(dst, (uint64_t)g.width * g.height, 256, 8), 256, g, dst->d_stage.as<float4>(), dst->d_stage_off.as<uint32_t>(),
              dst->d_row_base.as<uint32_t>(), dst->d_accum.as<float4>());
    PT_CK(cudaStreamSynchronize(dst->stream));
    dst->stats.kernel_launches = dst->call_launches; dst->stats.total_launches = dst->total_launches;
    return FOUNDATION_PT_OK;
    PT_CATCH(ctx)
}

}  // extern "C"

// ---- single-process group: one host thread drives N devices (SURVEY.md section 8e "Process model") -----------------------
struct foundation_pt_group {
    std::vector<foundation_pt_context*> members;
    std::string err = "no error";
};

extern "C" {

int32_t foundation_pt_group_create(const foundation_pt_config* config, const int32_t* devices, uint32_t count, uint32_t tile_size, uint32_t flags,
                                   const foundation_pt_allocator* host_alloc, foundation_pt_group** out_group) {
    if (!out_group) { g_create_error = "out_group is NULL"; return FOUNDATION_PT_ERR_ARGUMENT; }
    *out_group = nullptr;
    if (!config || !devices || count == 0 || count > 64) { g_create_error = "group_create: bad config / device list"; return FOUNDATION_PT_ERR_ARGUMENT; }
    foundation_pt_group* g = new (std::nothrow) foundation_pt_group();
    if (!g) { g_create_error = "host allocation failed"; return FOUNDATION_PT_ERR_OOM; }
    auto fail = [&](int32_t code, const std::string& msg) { g_create_error = msg; for (auto* m : g->members) foundation_pt_destroy(m); delete g; return code; };
    for (uint32_t i = 0; i < count; ++i) {
        foundation_pt_config c = *config; c.device = devices[i];
        foundation_pt_context* m = nullptr;
        int32_t rc = foundation_pt_create(&c, host_alloc, &m);
        if (rc) return fail(rc, g_create_error);
        g->members.push_back(m);
    }
    if (count > 1) {
        const NcclApi& api = nccl_api();
        if (!api.ok()) return fail(FOUNDATION_PT_ERR_COMM, api.err);
        std::vector<ncclComm_t> comms(count);
        std::vector<int> devs(devices, devices + count);
        ncclResult_t r = api.CommInitAll(comms.data(), (int)count, devs.data());
        if (r != ncclSuccess) return fail(FOUNDATION_PT_ERR_COMM, std::string("ncclCommInitAll: ") + api.GetErrorString(r));
        for (uint32_t i = 0; i < count; ++i) {
            Ctx* m = g->members[i];
            m->comm = comms[i]; m->comm_owned = true; m->comm_rank = i; m->comm_count = count; m->comm_flags = flags;
            m->part_rank = i; m->part_count = count; m->part_tile = tile_size ? tile_size : 32; m->wave_ready = false;
            cudaSetDevice(m->device);
            int32_t rc = ensure_accum(m);
            if (!rc) rc = plan_gather(m);
            if (rc) return fail(rc, m->err);
        }
        if (flags & FOUNDATION_PT_COMM_DIRECT) {
            for (uint32_t i = 1; i < count; ++i) {
                Ctx* m = g->members[i];
                if (m->device != g->members[0]->device) {
                    cudaSetDevice(m->device);
                    cudaError_t e = cudaDeviceEnablePeerAccess(g->members[0]->device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return fail(FOUNDATION_PT_ERR_COMM, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
                    cudaGetLastError();
                }
                m->remote_accum = g->members[0]->d_accum.as<float4>(); m->remote_is_ipc = false;
            }
        }
    }
    *out_group = g;
    return FOUNDATION_PT_OK;
}

int32_t foundation_pt_group_destroy(foundation_pt_group* group) {
    if (!group) return FOUNDATION_PT_ERR_ARGUMENT;
    for (auto* m : group->members) { cudaSetDevice(m->device); if (m->stream) cudaStreamSynchronize(m->stream); }
    for (auto* m : group->members) foundation_pt_destroy(m);
    delete group;
    return FOUNDATION_PT_OK;
}
uint32_t foundation_pt_group_size(const foundation_pt_group* group) { return group ? (uint32_t)group->members.size() : 0u; }
foundation_pt_context* foundation_pt_group_context(foundation_pt_group* group, uint32_t index) {
    return (group && index < group->members.size()) ? group->members[index] : nullptr;
}
const char* foundation_pt_group_last_error(const foundation_pt_group* group) { return group ? group->err.c_str() : g_create_error.c_str(); }

int32_t foundation_pt_group_render(foundation_pt_group* group, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces) {
    if (!group) return FOUNDATION_PT_ERR_ARGUMENT;
    auto fail = [&](int32_t rc, foundation_pt_context* m) { group->err = m->err; for (auto* o : group->members) foundation_pt_wait(o); return rc; };
    for (auto* m : group->members) { int32_t rc = foundation_pt_render_async(m, sample_begin, sample_count, max_bounces); if (rc) return fail(rc, m); }
    if (group->members.size() > 1) {
        int32_t rc = gather_enqueue(group->members.data(), (uint32_t)group->members.size(), 0);
        if (rc) return fail(rc, group->members[0]);
    }
    int32_t first = FOUNDATION_PT_OK;
    for (auto* m : group->members) {
        int32_t rc = foundation_pt_wait(m);
        if (!rc && group->members.size() > 1) rc = gather_finish(m);
        if (rc && !first) { first = rc; group->err = m->err; }
    }
    return first;
}

}  // extern "C"

/* foundation_pt.h — C ABI of the B200-native path-tracing backend for Foundation's src/Renderer.
 *
 * Plain C99: opaque context, POD descriptor structs with a leading struct_size, explicit byte sizes and
 * strides, int32 status returns (0 = OK, negative = error class; text via foundation_pt_last_error).
 * Nothing here throws, aborts, or exposes a torch / CUDA type.  Each entry cites the reference
 * interface it stands in for (paths relative to mos9527/Foundation); where the reference has no
 * counterpart (it ships no ray tracer — SURVEY.md §0) the citation is the slot in src/Renderer the
 * call occupies.
 *
 * Conventions taken from the reference:
 *   - matrices are column-major float[16], exactly `struct uniform_buffer` (src/Renderer/Renderer.cpp:28-33);
 *   - world is Z-up, right-handed; clip space is Vulkan's with the Y flip `proj[1][1] *= -1`
 *     (Renderer.cpp:373-380), so pixel (0,0) is the top-left corner;
 *   - vertex data arrives as R32G32B32_SIGNED_FLOAT with an explicit byte stride, indices as R16_UINT or
 *     R32_UINT (src/Platform/RHI/Common.hpp:18-27; index switch src/Platform/RHI/Vulkan/Command.cpp:292-302);
 *   - the colour target is R8G8B8A8_UNORM 1920x1080 (Renderer.cpp:40-41);
 *   - uploads are blocking and input pointers are borrowed only for the duration of the call, like the
 *     staging uploads at Renderer.cpp:133-197;
 *   - a context is single-caller (the reference is single-threaded) and calls return after the GPU work
 *     is complete, like Renderer::Draw (Renderer.cpp:394).
 */
#ifndef FOUNDATION_PT_H
#define FOUNDATION_PT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FOUNDATION_PT_API __declspec(dllexport)
#else
#define FOUNDATION_PT_API __attribute__((visibility("default")))
#endif

typedef struct foundation_pt_context foundation_pt_context;

/* status codes */
enum {
    FOUNDATION_PT_OK = 0,
    FOUNDATION_PT_ERR_ARGUMENT = -1,   /* bad pointer / size / enum / struct_size                       */
    FOUNDATION_PT_ERR_STATE = -2,      /* call order (e.g. render before scene_commit)                  */
    FOUNDATION_PT_ERR_CUDA = -3,       /* a CUDA runtime call failed; text has the cudaError string     */
    FOUNDATION_PT_ERR_OOM = -4,        /* host or device allocation failed                              */
    FOUNDATION_PT_ERR_NO_DEVICE = -5,  /* no CUDA device: there is deliberately NO CPU fallback         */
    FOUNDATION_PT_ERR_UNSUPPORTED = -6,
    FOUNDATION_PT_ERR_COMM = -7        /* NCCL missing / failed, or peer memory could not be mapped     */
};

/* Index formats: values mirror RHIResourceFormat usage at Vulkan/Command.cpp:292-302. */
enum { FOUNDATION_PT_INDEX_U16 = 16, FOUNDATION_PT_INDEX_U32 = 32, FOUNDATION_PT_INDEX_NONE = 0 /* unindexed soup */ };

/* Host-scratch allocation callbacks: same (user, size, alignment) triple as the reference's only extern "C"
 * precedent, vkCustomCpuAllocation / vkCustomCpuFree over Core::Allocator
 * (src/Platform/RHI/Vulkan/Application.hpp:11-26, Application.cpp:93-110).  May be NULL (malloc/free). */
typedef struct foundation_pt_allocator {
    void* user;
    void* (*alloc)(void* user, size_t size, size_t alignment);
    void (*free)(void* user, void* ptr);
} foundation_pt_allocator;

typedef struct foundation_pt_config {
    uint32_t struct_size;   /* = sizeof(foundation_pt_config)                                           */
    int32_t device;         /* CUDA device ordinal; the reference takes EnumerateDevices()[0] (Editor.cpp:18) */
    uint32_t width, height; /* render target; reference: swapchain extent 1920x1080 (Renderer.cpp:41)   */
    uint64_t seed;          /* PCG stream seed for the progressive render                                */
    uint32_t max_leaf_tris; /* 1..3 triangles per BVH8 leaf slot (0 = default 1)                        */
    uint32_t flags;         /* FOUNDATION_PT_FLAG_*                                                      */
    float background[3];    /* constant environment radiance returned on a miss                          */
    uint32_t reserved;
} foundation_pt_config;

enum {
    FOUNDATION_PT_FLAG_NO_MATERIAL_SORT = 1u << 0, /* never sort (this is also the default, see MATERIAL_SORT)               */
    FOUNDATION_PT_FLAG_NO_NEE = 1u << 1,           /* BSDF sampling only (oracle self-checks)                  */
    FOUNDATION_PT_FLAG_NO_BSDF_EMISSION = 1u << 2, /* NEE only: emitters hit by BSDF rays after bounce 0 add nothing */
    FOUNDATION_PT_FLAG_SOBOL_JITTER = 1u << 4,     /* sub-pixel positions from a per-pixel scrambled Sobol (0,2)-sequence instead of PCG32 */
    FOUNDATION_PT_FLAG_STAGE_TIMING = 1u << 6,     /* record CUDA events between the wavefront stages of a render -> foundation_pt_stats.stage_ms */
    FOUNDATION_PT_FLAG_SOBOL_PATH = 1u << 5,       /* light-point and BSDF-direction samples of every path vertex from padded, Owen-scrambled (0,2)-sequences */
    FOUNDATION_PT_FLAG_MATERIAL_SORT = 1u << 3     /* counting-sort live paths by material id before shading.  Off by default: with ONE
                                                      surface model for all materials the sort costs 8-14 % of a pass and buys nothing
                                                      (measured); results are identical either way.              */
};

/* Material: a two-lobe "diffuse + GGX" surface and an emitter.  32 bytes. */
typedef struct foundation_pt_material {
    float base_color[3];
    float roughness;   /* GGX alpha = roughness^2, clamped to >= 1e-3 */
    float emission[3]; /* radiance leaving the front side (geometric normal side) */
    float metallic;    /* 0: Lambert(base) + dielectric GGX coat F0=0.04; 1: conductor GGX F0=base_color */
} foundation_pt_material;

/* Instance: mesh id + 3x4 object->world, row-major rows [m00 m01 m02 tx] (the affine part of a glm::mat4
 * model matrix such as the one Renderer::Draw animates at Renderer.cpp:373, transposed to rows). 64 bytes. */
typedef struct foundation_pt_instance {
    uint32_t mesh_id;
    uint32_t reserved[3];
    float transform[12];
} foundation_pt_instance;

/* Ray and hit records of the parity / bench interface.  32 and 16 (+4) bytes, the SoA element sizes that
 * SURVEY.md §8(d)'s algorithmic-bytes formula counts. */
typedef struct foundation_pt_ray {
    float origin[3];
    float tmin;
    float direction[3]; /* need not be normalised; t is in units of |direction| */
    float tmax;         /* exclusive; +inf for unbounded */
} foundation_pt_ray;

typedef struct foundation_pt_hit {
    float t;            /* +inf on a miss */
    float u, v;         /* barycentrics of the hit (weights of vertex 1 and 2) */
    uint32_t prim;      /* triangle index inside its mesh (as passed to mesh_create); 0xFFFFFFFF on a miss */
} foundation_pt_hit;

typedef struct foundation_pt_build_stats {
    uint32_t struct_size;
    uint32_t num_meshes, num_instances;
    uint64_t num_triangles;       /* sum over meshes (unique) */
    uint64_t effective_triangles; /* sum over instances */
    uint64_t num_nodes8;          /* BVH8 nodes over all BLAS + TLAS */
    uint64_t device_bytes;        /* nodes + triangles resident in HBM */
    float build_ms;               /* device time of A1..A6 (CUDA events) */
    float sort_ms;                /* of which: radix sort */
    float scene_lo[3], scene_hi[3];
} foundation_pt_build_stats;

typedef struct foundation_pt_stats {
    uint32_t struct_size;
    uint32_t kernel_launches;  /* kernels of THIS library launched by the last render/trace call */
    uint64_t rays_extend;      /* closest-hit rays traced by the last render call */
    uint64_t rays_shadow;      /* any-hit rays traced by the last render call */
    float last_ms;             /* device time of the last render/trace call (CUDA events on the context's stream) */
    float trace_ms;            /* device time of the last rays_trace_* call (one traversal kernel)                            */
    float gather_ms;           /* device time of the last multi-GPU gather on this rank (pack + NCCL send/recv + scatter, or the direct mode's barrier) */
    uint32_t reserved;
    uint64_t total_launches;   /* since create */
    float stage_ms[6];         /* FOUNDATION_PT_FLAG_STAGE_TIMING: device time of the last render per wavefront stage (CUDA events between the
                                  launches): [0] ray generation, [1] extend (closest hit), [2] shade + NEE, [3] connect (any hit),
                                  [4] material sort, [5] accumulate.  All 0 without the flag.                                      */
} foundation_pt_stats;

/* ---- lifetime (reference slot: Renderer::Renderer / ~Renderer, src/Renderer/Renderer.cpp:34-311, 402-406) ---- */
FOUNDATION_PT_API int32_t foundation_pt_create(const foundation_pt_config* config, const foundation_pt_allocator* host_alloc,
                                               foundation_pt_context** out_ctx);
FOUNDATION_PT_API int32_t foundation_pt_destroy(foundation_pt_context* ctx);
/* Never NULL. ctx may be NULL (returns the message of the last failed create on this thread). */
FOUNDATION_PT_API const char* foundation_pt_last_error(const foundation_pt_context* ctx);

/* ---- scene upload (reference slot: the blocking vertex/index staging uploads, Renderer.cpp:133-197) ---- */
FOUNDATION_PT_API int32_t foundation_pt_materials_set(foundation_pt_context* ctx, const foundation_pt_material* materials, uint32_t count);
/* positions: float3 at `pos_stride_bytes` intervals (the reference's vertex_input is over-aligned by glm,
 * Renderer.cpp:23-27,110-115 — hence the explicit stride).  indices: 3 per triangle, or NULL with
 * INDEX_NONE (3 consecutive vertices per triangle).  material_ids: one uint32 per triangle or NULL (all 0). */
FOUNDATION_PT_API int32_t foundation_pt_mesh_create(foundation_pt_context* ctx, const void* positions, size_t pos_stride_bytes,
                                                    uint32_t num_vertices, const void* indices, uint32_t index_format,
                                                    uint32_t num_triangles, const uint32_t* material_ids, uint32_t* out_mesh_id);
/* Deforming meshes (the reference animates its model every Draw, Renderer.cpp:373): new vertex positions for an
 * existing mesh, same vertex count, stride and index buffer.  The mesh's BLAS is REBUILT by the next scene_commit (all other BLAS are kept): at
 * ~1.9 G triangles/s a rebuild costs what a refit would and the tree does not degrade.  Restart the accumulation with render(0, ...). */
FOUNDATION_PT_API int32_t foundation_pt_mesh_update_positions(foundation_pt_context* ctx, uint32_t mesh_id, const void* positions, size_t pos_stride_bytes,
                                                              uint32_t num_vertices);
/* Optional per-vertex attributes of a mesh — the rest of the reference's `vertex_input {pos, color, texCoord}` (Renderer.cpp:23-27; the
 * quad's values at :153-157): uv = float2, color = float3, each with its own byte stride (pointers into one interleaved, over-aligned
 * vertex buffer are fine), one entry per vertex of mesh_create.  Either may be NULL.  The hit's base colour becomes
 * material.base_color * colour(u, v) * texel(uv(u, v)), the path-traced counterpart of the reference's fragment shader
 * `texture.Sample(sampler, uv) * float4(color, 1)` (src/Renderer/Triangle.slang:34-37). */
FOUNDATION_PT_API int32_t foundation_pt_mesh_attributes_set(foundation_pt_context* ctx, uint32_t mesh_id, const void* uv, size_t uv_stride_bytes,
                                                            const void* color, size_t color_stride_bytes);
/* R8G8B8A8_UNORM image, rows from the top (what the reference uploads at Renderer.cpp:200-270; row_pitch_bytes 0 = tightly packed).  Sampled
 * bilinearly with REPEAT addressing: the RHI sampler's defaults (src/Platform/RHI/Device.hpp:71-99, created at Renderer.cpp:272-277;
 * one mip level, so its anisotropy setting has nothing to act on).  Alpha is ignored. */
#define FOUNDATION_PT_NO_TEXTURE 0xFFFFFFFFu
FOUNDATION_PT_API int32_t foundation_pt_texture_create(foundation_pt_context* ctx, const void* rgba8, uint32_t width, uint32_t height, size_t row_pitch_bytes,
                                                       uint32_t* out_texture_id);
/* One texture id (or FOUNDATION_PT_NO_TEXTURE) per material of materials_set: the albedo texture of that material. */
FOUNDATION_PT_API int32_t foundation_pt_material_textures_set(foundation_pt_context* ctx, const uint32_t* texture_ids, uint32_t count);
/* Optional.  Without it every mesh is instanced once with the identity transform.  A singular transform is reported by the
 * following scene_commit (FOUNDATION_PT_ERR_ARGUMENT): the inverse matrices are computed on the device. */
FOUNDATION_PT_API int32_t foundation_pt_instances_set(foundation_pt_context* ctx, const foundation_pt_instance* instances, uint32_t count);
/* Builds every BLAS (Morton LBVH -> BVH8) and the TLAS on the device.  stats may be NULL. */
FOUNDATION_PT_API int32_t foundation_pt_scene_commit(foundation_pt_context* ctx, foundation_pt_build_stats* stats);

/* ---- camera (reference: uniform_buffer.view / .proj written each Draw, Renderer.cpp:372-380) ---- */
FOUNDATION_PT_API int32_t foundation_pt_camera_set(foundation_pt_context* ctx, const float view[16], const float proj[16]);

/* ---- multi-GPU partition: this context renders only the pixels of tiles k with (k + row rotation) % count == rank.
 *      tile_size in pixels (0 = 32).  rank 0 / count 1 is the default (whole frame).  No reference counterpart:
 *      the reference is single-device (Editor.cpp:18). ---- */
FOUNDATION_PT_API int32_t foundation_pt_partition_set(foundation_pt_context* ctx, uint32_t rank, uint32_t count, uint32_t tile_size);

/* ---- multi-GPU frame (stage C1).  The frame is split into interleaved tiles (same rule as partition_set), the
 *      scene is replicated (every member uploads and commits the same scene: the device build is deterministic), and the owned tiles
 *      are gathered into ONE member's accumulation buffer over NVLink.  No reference counterpart: the reference is single-device
 *      (src/Editor/Editor.cpp:18); the slot is still Renderer::Draw (src/Renderer/Renderer.cpp:367-401), which stays one blocking call.
 *
 *      Two host shapes, one mechanism:
 *        one process per GPU   comm_unique_id on one rank -> distribute the 128 bytes by any channel -> comm_init on every rank
 *                              (collective) -> render ... -> gather (collective);
 *        one process, N GPUs   group_create(devices[], n) -> per-member scene upload through group_context(i) -> group_render.
 *      NCCL (libnccl.so.2) is loaded at run time by the first of these calls; a single-GPU host never needs it.
 *
 *      Gather modes (flags):
 *        0                          each rank packs its owned pixels and ncclSend()s them; the root ncclRecv()s and scatters.  At
 *                                   1080p on 8 GPUs that is 4.1 MB per rank instead of the 33.2 MB zero-padded frame of an all-reduce.
 *        FOUNDATION_PT_COMM_DIRECT  fused compute + collective: every rank's accumulate kernel stores its finished pixels straight into
 *                                   rank 0's frame over NVLink peer memory (peer access in a group, CUDA IPC across processes); gather
 *                                   is then only a 4-byte all-reduce that orders those stores.  root must be 0.
 *      Either way the gathered frame is bit-identical to the single-GPU frame (tests/test_gpu_multi.py, bench.py
 *      n_gpu_vs_1_gpu_max_abs_diff). ---- */
#define FOUNDATION_PT_COMM_ID_BYTES 128
enum { FOUNDATION_PT_COMM_DIRECT = 1u << 0 };
FOUNDATION_PT_API int32_t foundation_pt_comm_unique_id(uint8_t* id, size_t size_bytes);
/* Collective over `count` contexts (one per rank).  Fixes this context's tile partition to (rank, count, tile_size; 0 = 32). */
FOUNDATION_PT_API int32_t foundation_pt_comm_init(foundation_pt_context* ctx, const uint8_t* id, size_t size_bytes, uint32_t rank, uint32_t count,
                                                  uint32_t tile_size, uint32_t flags);
/* Collective.  Afterwards rank `root`'s accumulation buffer holds the whole frame (read_accum / resolve_rgba8 there). */
FOUNDATION_PT_API int32_t foundation_pt_gather(foundation_pt_context* ctx, uint32_t root);

/* Same-device form: scatters `src`'s owned tiles into `dst`'s frame (two different partitions of the same frame on ONE device) with the
 * pack / scatter kernels of the NCCL form, a device-to-device copy in place of ncclSend / ncclRecv.  Needs no NCCL. */
FOUNDATION_PT_API int32_t foundation_pt_gather_local(foundation_pt_context* dst, foundation_pt_context* src);

typedef struct foundation_pt_group foundation_pt_group;
FOUNDATION_PT_API int32_t foundation_pt_group_create(const foundation_pt_config* config /* .device ignored */, const int32_t* devices, uint32_t count,
                                                     uint32_t tile_size, uint32_t flags, const foundation_pt_allocator* host_alloc,
                                                     foundation_pt_group** out_group);
FOUNDATION_PT_API int32_t foundation_pt_group_destroy(foundation_pt_group* group);
FOUNDATION_PT_API uint32_t foundation_pt_group_size(const foundation_pt_group* group);
/* Member i (borrowed): upload the scene / set the camera on every member with the per-context calls.  Member 0 receives the frame. */
FOUNDATION_PT_API foundation_pt_context* foundation_pt_group_context(foundation_pt_group* group, uint32_t index);
/* All members render their tiles concurrently (render_async on every stream), the frame is gathered into member 0, then all are
 * waited for: one blocking call, like Draw(). */
FOUNDATION_PT_API int32_t foundation_pt_group_render(foundation_pt_group* group, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces);
FOUNDATION_PT_API const char* foundation_pt_group_last_error(const foundation_pt_group* group);

/* ---- render (reference slot: Renderer::Record's pass body, Renderer.cpp:332-351, driven by Draw :367-401) ----
 * Adds samples [sample_begin, sample_begin + sample_count) of every owned pixel to the accumulation buffer.
 * sample_begin == 0 clears the buffer first.
 * Memory: the first render of a context allocates the wavefront state — up to 32 M path slots (about 5.2 GB) per copy, two copies so that two waves of a
 * call can be in flight (FOUNDATION_PT_DUAL_WAVE=0 or FOUNDATION_PT_FLAG_STAGE_TIMING: one copy; one copy is also the fallback when the second does not fit).
 * The result does not depend on the number of copies or on how a call is cut into waves: every pixel adds its samples in ascending sample order. */
FOUNDATION_PT_API int32_t foundation_pt_render(foundation_pt_context* ctx, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces);
/* Asynchronous form (SURVEY.md section 8b "an async variant (render_async + wait) is optional"): render_async only enqueues the same work on
 * the context's stream and returns; wait blocks until it has finished, fills the stats and reports a traversal-stack overflow exactly like
 * render.  render == render_async + wait.  Any other entry point called in between first waits, so the context always behaves as if
 * it were idle (the reference's Draw() is blocking, Renderer.cpp:394; a host that wants to overlap its own per-frame work uses this pair). */
FOUNDATION_PT_API int32_t foundation_pt_render_async(foundation_pt_context* ctx, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces);
FOUNDATION_PT_API int32_t foundation_pt_wait(foundation_pt_context* ctx);
/* Linear radiance SUM (not yet divided by spp) as float4 per pixel, row-major from the top-left; w = sample count. */
FOUNDATION_PT_API int32_t foundation_pt_read_accum(foundation_pt_context* ctx, float* rgba, size_t size_bytes);
/* Restores an accumulation buffer saved with read_accum (checkpoint / resume of a progressive render; the
 * reference serialises nothing, §5).  Follow with render(sample_begin = samples already in the buffer, ...). */
FOUNDATION_PT_API int32_t foundation_pt_write_accum(foundation_pt_context* ctx, const float* rgba, size_t size_bytes);
/* accum / spp, clamped to [0,1], packed R8G8B8A8_UNORM (the reference's swapchain format, Renderer.cpp:40). */
FOUNDATION_PT_API int32_t foundation_pt_resolve_rgba8(foundation_pt_context* ctx, uint8_t* rgba8, size_t size_bytes);
/* Device address of the accumulation buffer (width*height float4) for zero-copy hand-off to another CUDA consumer in the same
 * process.  Valid until destroy. */
FOUNDATION_PT_API int32_t foundation_pt_accum_device_ptr(foundation_pt_context* ctx, void** out_device_ptr, size_t* out_size_bytes);

/* ---- parity / bench interface on explicit ray sets (no reference counterpart) ----
 * Host buffers: copies are inside the call (this is the `e2e` path of bench.py).  out_inst may be NULL. */
FOUNDATION_PT_API int32_t foundation_pt_trace_closest(foundation_pt_context* ctx, const foundation_pt_ray* rays, uint64_t count,
                                                      foundation_pt_hit* out_hits, uint32_t* out_inst);
FOUNDATION_PT_API int32_t foundation_pt_trace_any(foundation_pt_context* ctx, const foundation_pt_ray* rays, uint64_t count, uint8_t* out_occluded);
/* Device-resident ray sets: upload once, trace many times (this is the `value` path of bench.py). */
FOUNDATION_PT_API int32_t foundation_pt_rays_upload(foundation_pt_context* ctx, const foundation_pt_ray* rays, uint64_t count);
FOUNDATION_PT_API int32_t foundation_pt_rays_trace_closest(foundation_pt_context* ctx, uint64_t first, uint64_t count);
FOUNDATION_PT_API int32_t foundation_pt_rays_trace_any(foundation_pt_context* ctx, uint64_t first, uint64_t count);
FOUNDATION_PT_API int32_t foundation_pt_rays_download_hits(foundation_pt_context* ctx, uint64_t first, uint64_t count,
                                                           foundation_pt_hit* out_hits, uint32_t* out_inst);
/* Exhaustive O(rays x triangles) closest hit on the device — ground truth that uses no BVH (test utility). */
FOUNDATION_PT_API int32_t foundation_pt_rays_trace_brute(foundation_pt_context* ctx, uint64_t first, uint64_t count);

/* ---- introspection ---- */
FOUNDATION_PT_API int32_t foundation_pt_stats_get(foundation_pt_context* ctx, foundation_pt_stats* stats);
/* Copies the device-built acceleration structure of one mesh back to the host so tests can compare it
 * byte-for-byte with the oracle's build.  Any pointer may be NULL; counts are always written.
 * nodes: 80 bytes each; tris: 48 bytes each; order: uint32 per triangle (sorted position -> input triangle). */
FOUNDATION_PT_API int32_t foundation_pt_blas_download(foundation_pt_context* ctx, uint32_t mesh_id, void* nodes, size_t nodes_bytes,
                                                      void* tris, size_t tris_bytes, uint32_t* order, size_t order_bytes,
                                                      uint64_t* out_num_nodes, uint64_t* out_num_tris);
FOUNDATION_PT_API int32_t foundation_pt_tlas_download(foundation_pt_context* ctx, void* nodes, size_t nodes_bytes, uint32_t* order,
                                                      size_t order_bytes, uint64_t* out_num_nodes, uint64_t* out_num_instances);
FOUNDATION_PT_API uint32_t foundation_pt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FOUNDATION_PT_H */

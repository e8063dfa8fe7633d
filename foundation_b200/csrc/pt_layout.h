// pt_layout.h — HBM data layout of the acceleration structure and the wavefront state, plus the
// few arithmetic rules that define it (outward quantisation, padding).  Shared by the CUDA kernels and,
// for the struct definitions and rounding rules only, by the CPU oracle.
//
// The reference has no acceleration structure (only device extension is VK_KHR_swapchain,
// src/Platform/RHI/Vulkan/Device.cpp:13-15); the layout follows north_star: "compressed 8-wide BVH nodes
// with quantised AABBs", 80-byte nodes = 5 x 128-bit loads, 48-byte triangles = 3 x 128-bit loads
// (SURVEY.md §8a2 rows A5/B2, §8d bytes_ray = 32 + 16 + 80 V + 48 T).
#pragma once
#include "pt_math.h"

// ---- BVH8 node, 80 bytes --------------------------------------------------------------------
// word 0: px py pz | ex ey ez imask      origin (= padded node lo), per-axis scale 2^(e-127), internal mask
// word 1: child_base | tri_base | meta[0..3] | meta[4..7]
// word 2: qlo_x[0..7] qlo_y[0..7]        child boxes, 8 bit per plane on the node grid, rounded OUTWARD
// word 3: qlo_z[0..7] qhi_x[0..7]
// word 4: qhi_y[0..7] qhi_z[0..7]
// meta[s]: 0 = empty slot; internal child: 0b001_11sss (= 0x20 | 24 + s); leaf: (unary count 1/3/7) << 5 | first
// triangle offset (0..23) inside the node's triangle block.  Slot s holds the child that lies towards
// octant s (bit 2 = +x, bit 1 = +y, bit 0 = +z) so `slot ^ ray_octant` orders children front to back.
struct PtNode8 {
    float px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t child_base, tri_base;
    uint8_t meta[8];
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
#if defined(__cplusplus)
static_assert(sizeof(PtNode8) == 80, "PtNode8 must be 80 bytes");
#endif

// ---- triangle, 48 bytes ----------------------------------------------------------------------
// BLAS: prim = input triangle index, mat = material id.  Stored in BVH8 leaf order.
struct PtTri {
    float v0x, v0y, v0z; uint32_t prim;
    float e1x, e1y, e1z; uint32_t mat;
    float e2x, e2y, e2z; uint32_t pad;
};

// ---- instance record, 112 bytes (7 x 16) ------------------------------------------------------
struct PtInstance {
    float w2o[12];      // world -> object, rows
    float o2w[12];      // object -> world, rows
    uint32_t node_base; // BLAS root node index in the global node array
    uint32_t tri_base;  // BLAS first triangle in the global triangle array
    uint32_t mesh_id, inst_id;  // inst_id: index in the caller's instance array (records are stored in TLAS leaf order)
};

// ---- BVH2 node produced by the LBVH emit (build-time only) --------------------------------------
// children: index < n-1 -> internal node, else leaf (index - (n-1)) in sorted order.
struct PtBox { float lox, loy, loz, hix, hiy, hiz; };

// ---- light triangle (world space), 64 bytes -----------------------------------------------------
struct PtLight {
    float v0x, v0y, v0z, cdf;     // cdf: cumulative area fraction up to and including this triangle
    float e1x, e1y, e1z, area;
    float e2x, e2y, e2z, pad0;
    float emr, emg, emb, pad1;
};

struct PtMaterial { float r, g, b, roughness, er, eg, eb, metallic; };

// camera: dir(x,y) = d0 + x*dx + y*dy for NDC x,y in [-1,1]; derived on the host (double) from
// inverse(proj*view) with the reference's conventions (Renderer.cpp:373-380).
struct PtCamera { float eye[3], d0[3], dx[3], dy[3]; };

#define PT_PAD_REL 1.9073486328125e-06f  // 2^-19: child boxes are padded by this x max |coordinate| before quantisation
#define PT_MAX_LEAF 1   // default triangles per leaf slot (the format allows 3): measured 7-9 % faster than 3 on configs 2-4, 1.9x on Cornell:
                       // every extra triangle of a leaf is one more random 64-byte DRAM fetch, a tighter box avoids it
// Ray-dependent slack of the slab test.  A plane distance is evaluated as t = fma(QBIAS + q, a, c) with a = scale * idir,
// b = (p - o) * idir and c = b - QBIAS * a (the bias lets the byte q be turned into a float with one byte permute, pt_qfloat).  That is
// five roundings, each at most 2^-24 of (|b| + (QBIAS + 255) |a|); every plane is widened by PT_SLAB_EPS = 2^-21 times that bound
// (x1.6 margin): 1.6 % of one quantisation step plus 2^-21 of the distance term.  The build-time pad covers coordinates near the mesh;
// this covers rays whose origin is far away compared with the mesh (instanced BLAS in object space, distant cameras), where the error
// of (p - o) * idir grows with the distance.
#define PT_SLAB_EPS 4.76837158203125e-07f
#define PT_QBIAS 32768.0f
#define PT_QBIAS_BITS 0x47000000u
#define PT_SLAB_QMAX 33023.0f   // QBIAS + 255

// relative costs of the collapse plan (pt_build.h): visiting a wide node / testing one triangle, per unit of surface area
#define PT_COST_NODE 1.0f
#ifndef PT_COST_TRI
#define PT_COST_TRI 0.3f
#endif

// ---- quantisation rules -------------------------------------------------------------------------
// biased exponent e such that 255 * 2^(e-127) >= extent (smallest such e, clamped to [1,253])
PT_HD uint32_t pt_quant_exp(float extent) {
    uint32_t k = (pt_f2u(extent) >> 23) & 0xffu;          // extent in [2^(k-127), 2^(k-126))
    int e = (int)k - 7;                                    // try scale 2^(k-7-127): 255*scale in [..)
    if (e < 1) e = 1;
    float s = pt_u2f((uint32_t)e << 23);
    if (!(255.0f * s >= extent)) e += 1;                   // 255*2^(k-6) >= 2^(k+1) > extent always
    if (e > 253) e = 253;                                  // keep 2^(127-e) (the inverse scale) a normal float
    return (uint32_t)e;
}
PT_HD float pt_floor(float x) {  // exact floor for |x| < 2^23 (all our operands are in [0, 256])
    float t = (float)(int)x;
    return t > x ? t - 1.0f : t;
}
PT_HD float pt_ceil(float x) {
    float t = (float)(int)x;
    return t < x ? t + 1.0f : t;
}
// quantise one plane pair; inv = 2^-(e-127) (exact power of two), p = node origin
PT_HD uint32_t pt_quant_lo(float lo, float p, float inv) {
    float q = pt_floor((lo - p) * inv);
    return (uint32_t)pt_clamp(q, 0.0f, 255.0f);
}
PT_HD uint32_t pt_quant_hi(float hi, float p, float inv) {
    float q = pt_ceil((hi - p) * inv);
    return (uint32_t)pt_clamp(q, 0.0f, 255.0f);
}
PT_HD float pt_box_area(float lx, float ly, float lz, float hx, float hy, float hz) {
    float dx = hx - lx, dy = hy - ly, dz = hz - lz;
    return pt_fma(dx, dy, pt_fma(dy, dz, dz * dx));
}

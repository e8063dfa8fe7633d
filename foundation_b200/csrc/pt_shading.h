// pt_shading.h — the per-vertex shading arithmetic of the wavefront path tracer (stages B1, B3, B4
// of SURVEY.md §8a2): camera ray generation, the diffuse + GGX surface model, one-sample next-event
// estimation on emissive triangles with power-heuristic MIS, Russian roulette.
//
// Same rule as pt_math.h: only + - * / sqrt fma and the polynomial sincos, so the CUDA kernels and the
// CPU oracle produce bit-identical radiance for identical PCG streams (north_star asks RMSE <= 1e-3; the
// tests additionally record how many pixels differ at all).  The reference contains no shading model of
// its own (its fragment shader is a texture fetch, src/Renderer/Triangle.slang:34-37); the camera
// conventions come from src/Renderer/Renderer.cpp:373-380.
#pragma once
#include "pt_layout.h"

#define PT_FLAG_NO_MATERIAL_SORT 1u
#define PT_FLAG_NO_NEE 2u
#define PT_FLAG_NO_BSDF_EMISSION 4u
#define PT_FLAG_MATERIAL_SORT 8u
#define PT_FLAG_SOBOL_JITTER 16u
#define PT_FLAG_SOBOL_PATH 32u

struct PtShadeConsts {
    const PtLight* lights;
    uint32_t num_lights;
    float light_area;      // total emissive area (world space)
    float ray_eps;         // origin offset along the normal for secondary rays
    uint32_t flags;
    uint32_t max_bounces;
    float bg[3];
    uint64_t seed;         // the render seed (per-pixel scramble keys of PT_FLAG_SOBOL_PATH)
};

// Everything a path carries between bounces (SoA on the device, a local struct in the oracle).
struct PtPath {
    pt_v3 o, d;        // current ray (d normalised)
    pt_v3 beta;        // throughput
    pt_v3 L;           // radiance gathered by this sample so far
    pt_rng rng;
    float pdf_prev;    // solid-angle pdf of the BSDF sample that produced the current ray (0 on the primary ray)
    uint32_t pixel;
    uint32_t bounce;   // index of the vertex the current ray will hit (0 = primary)
};

struct PtShadowRay {
    pt_v3 o, d;        // d unnormalised: the light point is at t = 1
    float tmax;
    pt_v3 contrib;     // added to L if unoccluded
    bool valid;
};

// ---- B1: ray generation ---------------------------------------------------------------------------
PT_HD void pt_camera_ray(const PtCamera& cam, uint32_t px, uint32_t py, uint32_t width, uint32_t height, float jx, float jy,
                         pt_v3* o, pt_v3* d) {
    // NDC of the sample: x right, y DOWN (Vulkan clip space after the reference's proj[1][1] *= -1), top-left pixel 0,0
    float nx = pt_fma(((float)px + jx), pt_div(2.0f, (float)width), -1.0f);
    float ny = pt_fma(((float)py + jy), pt_div(2.0f, (float)height), -1.0f);
    pt_v3 dir = pt_mk(pt_fma(ny, cam.dy[0], pt_fma(nx, cam.dx[0], cam.d0[0])), pt_fma(ny, cam.dy[1], pt_fma(nx, cam.dx[1], cam.d0[1])),
                      pt_fma(ny, cam.dy[2], pt_fma(nx, cam.dx[2], cam.d0[2])));
    *o = pt_mk(cam.eye[0], cam.eye[1], cam.eye[2]);
    *d = pt_normalize(dir);
}

// Sub-pixel positions from the Sobol (0,2)-sequence: dimension 0 is the radical inverse in base 2 (van der Corput), dimension 1 the
// second Sobol dimension (direction numbers v_k = v_{k-1} ^ (v_{k-1} >> 1), i.e. the Pascal-triangle generator matrix), each XOR-
// scrambled with a per-pixel key (random digit scrambling keeps the (0,2) stratification: any 2^k consecutive samples starting at
// a multiple of 2^k put one point in every elementary interval of area 2^-k).  Integer-only, so both machines agree to the bit.
PT_HD uint32_t pt_reverse_bits32(uint32_t x) {
    x = (x << 16) | (x >> 16);
    x = ((x & 0x00ff00ffu) << 8) | ((x & 0xff00ff00u) >> 8);
    x = ((x & 0x0f0f0f0fu) << 4) | ((x & 0xf0f0f0f0u) >> 4);
    x = ((x & 0x33333333u) << 2) | ((x & 0xccccccccu) >> 2);
    x = ((x & 0x55555555u) << 1) | ((x & 0xaaaaaaaau) >> 1);
    return x;
}
PT_HD uint32_t pt_sobol2(uint32_t i) {
    uint32_t r = 0;
    for (uint32_t v = 0x80000000u; i; i >>= 1, v ^= v >> 1)
        if (i & 1u) r ^= v;
    return r;
}
PT_HD void pt_sobol02(uint32_t sample, uint32_t key0, uint32_t key1, float* x, float* y) {
    *x = (float)((pt_reverse_bits32(sample) ^ key0) >> 8) * 5.9604644775390625e-08f;
    *y = (float)((pt_sobol2(sample) ^ key1) >> 8) * 5.9604644775390625e-08f;
}

// Path dimensions (PT_FLAG_SOBOL_PATH): the two 2-D decisions of every path vertex — the point on the light (u1, u2) and the BSDF
// direction (u4, u5) — take sample i of a (0,2)-sequence that is "padded" per (pixel, bounce, decision): the sample index is shuffled
// and both coordinates are Owen-scrambled with the hash-based nested uniform scramble of Laine & Karras ("Stratified sampling for
// stochastic transparency", 2011) as used by Burley, "Practical Hash-based Owen Scrambling" (JCGT 2020).  x ^= x * even only moves
// information from low to high bits, so reverse -> hash -> reverse is a nested (Owen) permutation: every 2^k-sample prefix of every
// decision stays a (0,2)-net, and different decisions are decorrelated by their keys.  Integer-only: bit-identical on both machines.
// The 1-D decisions (light pick, lobe pick, Russian roulette) stay on the PCG32 stream, whose layout does not depend on the flag.
PT_HD uint32_t pt_lk_hash(uint32_t x, uint32_t seed) {
    x += seed;
    x ^= x * 0x6c50b47cu; x ^= x * 0xb82f1e52u; x ^= x * 0xc7afe638u; x ^= x * 0x8d22f6e6u;
    return x;
}
PT_HD uint32_t pt_owen32(uint32_t x, uint32_t seed) { return pt_reverse_bits32(pt_lk_hash(pt_reverse_bits32(x), seed)); }
PT_HD void pt_sobol02_padded(uint32_t sample, uint64_t key, float* x, float* y) {
    const uint64_t k2 = pt_mix64(key);
    const uint32_t i = pt_owen32(sample, (uint32_t)key);                         // shuffled index (nested: aligned 2^k blocks stay blocks)
    *x = (float)(pt_owen32(pt_reverse_bits32(i), (uint32_t)(key >> 32)) >> 8) * 5.9604644775390625e-08f;
    *y = (float)(pt_owen32(pt_sobol2(i), (uint32_t)k2) >> 8) * 5.9604644775390625e-08f;
}
// key of decision `which` (0 = light point, 1 = BSDF direction) at path vertex `bounce` of `pixel`
PT_HD uint64_t pt_path_dim_key(uint64_t seed, uint32_t pixel, uint32_t bounce, uint32_t which) {
    return pt_mix64(seed ^ pt_mix64(0x50b02ull + pixel + ((uint64_t)(2u * bounce + which + 1u) << 32)));
}

PT_HD void pt_path_init(PtPath* p, const PtCamera& cam, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t width, uint32_t height,
                        uint32_t flags = 0) {
    p->rng = pt_rng_for(seed, pixel, sample);
    float jx = pt_rng_f(&p->rng), jy = pt_rng_f(&p->rng);     // always drawn: the stream layout does not depend on the flag
    if (flags & PT_FLAG_SOBOL_JITTER) {
        uint64_t k = pt_mix64(seed ^ pt_mix64(0x50b01ull + pixel));
        pt_sobol02(sample, (uint32_t)k, (uint32_t)(k >> 32), &jx, &jy);
    }
    pt_camera_ray(cam, pixel % width, pixel / width, width, height, jx, jy, &p->o, &p->d);
    p->beta = pt_mk(1.0f, 1.0f, 1.0f);
    p->L = pt_mk(0.0f, 0.0f, 0.0f);
    p->pdf_prev = 0.0f;
    p->pixel = pixel;
    p->bounce = 0;
}

// ---- per-vertex attributes and the albedo texture -------------------------------------------------------
// The only material inputs the reference's renderer has: `vertex_input {pos, color, texCoord}` (mos9527/Foundation
// src/Renderer/Renderer.cpp:23-27, the quad at :153-157) and one RGBA8 texture sampled in the fragment shader,
// `texture.Sample(sampler, uv) * float4(color, 1)` (src/Renderer/Triangle.slang:34-37), through a sampler with the RHI's defaults —
// linear filter, REPEAT addressing (src/Platform/RHI/Device.hpp:71-99; the image has one mip level, so anisotropy / mip modes are moot).
// Here the product enters the surface model as the base colour:  base = material.base_color * colour(u, v) * texel(uv(u, v)).
// Same arithmetic rule as the rest of this file (+ - * / fma, fixed order), so the device and the oracle agree to the bit.
struct PtTexture { const uint32_t* texels; uint32_t width, height, pad; };   // R8G8B8A8_UNORM (format of Renderer.cpp:205), rows from the top, tightly packed
struct PtMeshAttr {                                                          // per mesh; NULL stream = attribute absent (colour 1, no uv)
    const uint8_t* uv; const uint8_t* col; const void* idx;
    uint32_t uv_stride, col_stride, idx_fmt /* 32, 16, 0 = unindexed */, pad;
};
PT_HD float pt_wrap01(float x) {      // x - floor(x) in [0, 1); 0 for NaN / inf / |x| >= 2^23 (no fractional bits left)
    if (!(pt_abs(x) < 8388608.0f)) return 0.0f;
    float f = x - pt_floor(x);
    return f < 1.0f ? f : 0.0f;       // -1e-9 - floor(-1e-9) rounds to 1
}
PT_HD pt_v3 pt_texel(const PtTexture& t, uint32_t x, uint32_t y) {
    uint32_t p = t.texels[(size_t)y * t.width + x];
    const float k = 0.00392156885936856270f;   // float(1 / 255)
    return pt_mk((float)(p & 0xffu) * k, (float)((p >> 8) & 0xffu) * k, (float)((p >> 16) & 0xffu) * k);
}
// bilinear, REPEAT in both directions, texel centres at half-integers (Vulkan's unnormalised coordinate u * W - 0.5)
PT_HD pt_v3 pt_texture_sample(const PtTexture& t, float u, float v) {
    float x = pt_fma(pt_wrap01(u), (float)t.width, -0.5f), y = pt_fma(pt_wrap01(v), (float)t.height, -0.5f);   // in [-0.5, W - 0.5)
    float fx0 = pt_floor(x), fy0 = pt_floor(y);
    float fx = x - fx0, fy = y - fy0;
    int ix = (int)fx0, iy = (int)fy0;                                                                          // -1 .. W - 1
    uint32_t x0 = ix < 0 ? t.width - 1u : (uint32_t)ix, y0 = iy < 0 ? t.height - 1u : (uint32_t)iy;
    uint32_t x1 = x0 + 1u < t.width ? x0 + 1u : 0u, y1 = y0 + 1u < t.height ? y0 + 1u : 0u;
    pt_v3 c00 = pt_texel(t, x0, y0), c10 = pt_texel(t, x1, y0), c01 = pt_texel(t, x0, y1), c11 = pt_texel(t, x1, y1);
    pt_v3 a = pt_madd(c00, fx, pt_sub(c10, c00)), b = pt_madd(c01, fx, pt_sub(c11, c01));
    return pt_madd(a, fy, pt_sub(b, a));
}
PT_HD uint32_t pt_attr_index(const PtMeshAttr& a, uint32_t prim, uint32_t k) {
    if (a.idx_fmt == 32u) return ((const uint32_t*)a.idx)[3 * (size_t)prim + k];
    if (a.idx_fmt == 16u) return ((const uint16_t*)a.idx)[3 * (size_t)prim + k];
    return 3u * prim + k;
}
// base colour of the hit: material x interpolated vertex colour x texel.  bu, bv: barycentric weights of vertex 1 and 2 (the hit record's u, v).
PT_HD void pt_material_apply_attributes(PtMaterial* m, const PtMeshAttr& a, const PtTexture* textures, uint32_t tex_id, uint32_t prim, float bu, float bv) {
    if (!a.uv && !a.col) return;
    uint32_t i0 = pt_attr_index(a, prim, 0), i1 = pt_attr_index(a, prim, 1), i2 = pt_attr_index(a, prim, 2);
    float w0 = (1.0f - bu) - bv;
    if (a.col) {
        const float *c0 = (const float*)(a.col + (size_t)i0 * a.col_stride), *c1 = (const float*)(a.col + (size_t)i1 * a.col_stride),
                    *c2 = (const float*)(a.col + (size_t)i2 * a.col_stride);
        m->r *= pt_fma(bv, c2[0], pt_fma(bu, c1[0], w0 * c0[0]));
        m->g *= pt_fma(bv, c2[1], pt_fma(bu, c1[1], w0 * c0[1]));
        m->b *= pt_fma(bv, c2[2], pt_fma(bu, c1[2], w0 * c0[2]));
    }
    if (a.uv && tex_id != PT_NONE) {
        const float *t0 = (const float*)(a.uv + (size_t)i0 * a.uv_stride), *t1 = (const float*)(a.uv + (size_t)i1 * a.uv_stride),
                    *t2 = (const float*)(a.uv + (size_t)i2 * a.uv_stride);
        float u = pt_fma(bv, t2[0], pt_fma(bu, t1[0], w0 * t0[0])), v = pt_fma(bv, t2[1], pt_fma(bu, t1[1], w0 * t0[1]));
        pt_v3 c = pt_texture_sample(textures[tex_id], u, v);
        m->r *= c.x; m->g *= c.y; m->b *= c.z;
    }
}

// ---- surface model ----------------------------------------------------------------------------------
struct PtBsdf {
    pt_v3 kd, f0;
    float alpha, a2, p_spec;
};
PT_HD PtBsdf pt_bsdf_make(const PtMaterial& m) {
    PtBsdf b;
    float om = 1.0f - m.metallic;
    b.kd = pt_mk(m.r * om, m.g * om, m.b * om);
    b.f0 = pt_mk(pt_fma(m.r - 0.04f, m.metallic, 0.04f), pt_fma(m.g - 0.04f, m.metallic, 0.04f), pt_fma(m.b - 0.04f, m.metallic, 0.04f));
    b.alpha = pt_max(m.roughness * m.roughness, 1e-3f);
    b.a2 = b.alpha * b.alpha;
    b.p_spec = pt_fma(0.5f, m.metallic, 0.5f);
    return b;
}
PT_HD float pt_lambda(float a2, float cz) {  // Smith Lambda for GGX, cz = cos(theta) > 0
    float c2 = cz * cz;
    float t2 = pt_div(pt_max(1.0f - c2, 0.0f), c2);
    return 0.5f * (pt_sqrt(pt_fma(a2, t2, 1.0f)) - 1.0f);
}
// f (without the cosine) and the mixture pdf, local frame (z = shading normal), wo.z > 0, wi.z > 0
PT_HD void pt_bsdf_eval(const PtBsdf& b, pt_v3 wo, pt_v3 wi, pt_v3* f, float* pdf) {
    float woz = pt_max(wo.z, 1e-6f), wiz = pt_max(wi.z, 1e-6f);
    pt_v3 h = pt_normalize(pt_add(wo, wi));
    float dh = pt_max(pt_dot(wo, h), 0.0f);
    float k = pt_fma(h.z * h.z, b.a2 - 1.0f, 1.0f);
    float D = pt_div(b.a2, PT_PI * k * k);
    float lo = pt_lambda(b.a2, woz), li = pt_lambda(b.a2, wiz);
    float G2 = pt_div(1.0f, 1.0f + lo + li);
    float G1 = pt_div(1.0f, 1.0f + lo);
    float m = 1.0f - dh, m2 = m * m, m5 = m2 * m2 * m;
    pt_v3 F = pt_mk(pt_fma(1.0f - b.f0.x, m5, b.f0.x), pt_fma(1.0f - b.f0.y, m5, b.f0.y), pt_fma(1.0f - b.f0.z, m5, b.f0.z));
    float sp = pt_div(D * G2, 4.0f * woz * wiz);
    // diffuse base under the coat: weighted by the energy NOT reflected specularly on the way in and out,
    // (1 - F(n.wo)) (1 - F(n.wi)) — symmetric in wo, wi (reciprocal) and keeps the two-lobe albedo <= 1 at grazing angles
    float mo = 1.0f - woz, mo2 = mo * mo, mo5 = mo2 * mo2 * mo;
    float mi = 1.0f - wiz, mi2 = mi * mi, mi5 = mi2 * mi2 * mi;
    pt_v3 Td = pt_mk((1.0f - pt_fma(1.0f - b.f0.x, mo5, b.f0.x)) * (1.0f - pt_fma(1.0f - b.f0.x, mi5, b.f0.x)),
                     (1.0f - pt_fma(1.0f - b.f0.y, mo5, b.f0.y)) * (1.0f - pt_fma(1.0f - b.f0.y, mi5, b.f0.y)),
                     (1.0f - pt_fma(1.0f - b.f0.z, mo5, b.f0.z)) * (1.0f - pt_fma(1.0f - b.f0.z, mi5, b.f0.z)));
    *f = pt_mk(pt_fma(b.kd.x * PT_INV_PI, Td.x, F.x * sp), pt_fma(b.kd.y * PT_INV_PI, Td.y, F.y * sp), pt_fma(b.kd.z * PT_INV_PI, Td.z, F.z * sp));
    float pdf_s = pt_div(G1 * D, 4.0f * woz);
    float pdf_d = wiz * PT_INV_PI;
    *pdf = pt_fma(b.p_spec, pdf_s, (1.0f - b.p_spec) * pdf_d);
}
// Samples wi (local). Returns false when the sample falls below the horizon.
PT_HD bool pt_bsdf_sample(const PtBsdf& b, pt_v3 wo, float ul, float u1, float u2, pt_v3* wi) {
    float s, c;
    pt_sincos2pi(u2, &s, &c);
    float r = pt_sqrt(u1);
    if (ul < b.p_spec) {
        // GGX visible-normal sampling (Heitz 2018)
        pt_v3 vh = pt_normalize(pt_mk(b.alpha * wo.x, b.alpha * wo.y, wo.z));
        float lensq = pt_fma(vh.x, vh.x, vh.y * vh.y);
        pt_v3 t1 = pt_mk(1.0f, 0.0f, 0.0f);
        if (lensq > 0.0f) { float il = pt_div(1.0f, pt_sqrt(lensq)); t1 = pt_mk(-vh.y * il, vh.x * il, 0.0f); }
        pt_v3 t2 = pt_cross(vh, t1);
        float a = r * c, bb = r * s;
        float sw = 0.5f * (1.0f + vh.z);
        bb = pt_fma(1.0f - sw, pt_sqrt(pt_max(pt_fma(-a, a, 1.0f), 0.0f)), sw * bb);
        float cz = pt_sqrt(pt_max(1.0f - pt_fma(a, a, bb * bb), 0.0f));
        pt_v3 nh = pt_madd(pt_madd(pt_scale(t1, a), bb, t2), cz, vh);
        pt_v3 h = pt_normalize(pt_mk(b.alpha * nh.x, b.alpha * nh.y, pt_max(nh.z, 0.0f)));
        float dh = pt_dot(wo, h);
        *wi = pt_sub(pt_scale(h, 2.0f * dh), wo);
    } else {
        *wi = pt_mk(r * c, r * s, pt_sqrt(pt_max(1.0f - u1, 0.0f)));
    }
    return wi->z > 0.0f;
}

// first light whose cumulative area fraction reaches u
PT_HD uint32_t pt_pick_light(const PtLight* lights, uint32_t n, float u) {
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (lights[mid].cdf < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- B3 + B4: shade one path vertex -------------------------------------------------------------------
// Inputs: the path (current ray), the hit distance, the hit triangle in WORLD space (v0, e1, e2) and its
// material.  Effects: adds emission to path->L, optionally emits one shadow ray, samples the BSDF and
// rewrites the path's ray.  Returns true while the path is alive.  Consumes exactly 7 random numbers.
PT_HD bool pt_shade_vertex(PtPath* path, const PtShadeConsts& sc, float t, pt_v3 e1, pt_v3 e2, const PtMaterial& mat, PtShadowRay* sh) {
    sh->valid = false;
    pt_v3 ng = pt_normalize(pt_cross(e1, e2));
    float dn = pt_dot(ng, path->d);
    bool front = dn < 0.0f;
    pt_v3 n = front ? ng : pt_neg(ng);
    bool emissive = (mat.er > 0.0f) || (mat.eg > 0.0f) || (mat.eb > 0.0f);
    if (emissive) {
        if (front) {
            float w = 1.0f;
            if (path->bounce > 0) {
                if (sc.flags & PT_FLAG_NO_BSDF_EMISSION) w = 0.0f;
                else if (!(sc.flags & PT_FLAG_NO_NEE) && sc.num_lights > 0) {
                    float pl = pt_div(t * t, -dn * sc.light_area);
                    float pb2 = path->pdf_prev * path->pdf_prev;
                    w = pt_div(pb2, pt_fma(pl, pl, pb2));
                }
            }
            path->L = pt_mk(pt_fma(path->beta.x * mat.er, w, path->L.x), pt_fma(path->beta.y * mat.eg, w, path->L.y),
                            pt_fma(path->beta.z * mat.eb, w, path->L.z));
        }
        return false;  // emitters do not scatter
    }
    if (path->bounce >= sc.max_bounces) return false;

    pt_v3 p = pt_madd(path->o, t, path->d);
    pt_v3 po = pt_madd(p, sc.ray_eps, n);
    pt_v3 T, B;
    pt_onb(n, &T, &B);
    pt_v3 wo_w = pt_neg(path->d);
    pt_v3 wo = pt_mk(pt_dot(wo_w, T), pt_dot(wo_w, B), pt_dot(wo_w, n));
    PtBsdf bsdf = pt_bsdf_make(mat);

    // B4: next-event estimation (always consumes 3 numbers so the stream layout is fixed)
    float u0 = pt_rng_f(&path->rng), u1 = pt_rng_f(&path->rng), u2 = pt_rng_f(&path->rng);
    const uint32_t sample = (uint32_t)(path->rng.inc >> 33);   // the stream id is (sample << 32 | pixel) << 1 | 1 (pt_rng_for)
    if (sc.flags & PT_FLAG_SOBOL_PATH) pt_sobol02_padded(sample, pt_path_dim_key(sc.seed, path->pixel, path->bounce, 0u), &u1, &u2);
    if (!(sc.flags & PT_FLAG_NO_NEE) && sc.num_lights > 0) {
        PtLight lt = sc.lights[pt_pick_light(sc.lights, sc.num_lights, u0)];
        float su = pt_sqrt(u1);
        float b1 = 1.0f - su, b2 = u2 * su;
        pt_v3 le1 = pt_mk(lt.e1x, lt.e1y, lt.e1z), le2 = pt_mk(lt.e2x, lt.e2y, lt.e2z);
        pt_v3 pl = pt_madd(pt_madd(pt_mk(lt.v0x, lt.v0y, lt.v0z), b1, le1), b2, le2);
        pt_v3 wu = pt_sub(pl, po);
        float d2 = pt_dot(wu, wu);
        float inv_d = pt_div(1.0f, pt_sqrt(d2));
        pt_v3 wi_w = pt_scale(wu, inv_d);
        pt_v3 nl = pt_normalize(pt_cross(le1, le2));
        float cl = -pt_dot(nl, wi_w);
        float cs = pt_dot(n, wi_w);
        if (cl > 0.0f && cs > 0.0f && d2 > 0.0f) {
            pt_v3 wi = pt_mk(pt_dot(wi_w, T), pt_dot(wi_w, B), cs);
            pt_v3 f; float pb;
            pt_bsdf_eval(bsdf, wo, wi, &f, &pb);
            float plight = pt_div(d2, cl * sc.light_area);
            float w = pt_div(plight * plight, pt_fma(plight, plight, pb * pb));
            float g = pt_div(cs * w, plight);
            sh->o = po; sh->d = wu; sh->tmax = 0.999f;
            sh->contrib = pt_mk(path->beta.x * f.x * lt.emr * g, path->beta.y * f.y * lt.emg * g, path->beta.z * f.z * lt.emb * g);
            sh->valid = true;
        }
    }

    // B3: BSDF sample
    float u3 = pt_rng_f(&path->rng), u4 = pt_rng_f(&path->rng), u5 = pt_rng_f(&path->rng), u6 = pt_rng_f(&path->rng);
    if (sc.flags & PT_FLAG_SOBOL_PATH) pt_sobol02_padded(sample, pt_path_dim_key(sc.seed, path->pixel, path->bounce, 1u), &u4, &u5);
    pt_v3 wi;
    if (!pt_bsdf_sample(bsdf, wo, u3, u4, u5, &wi)) return false;
    pt_v3 f; float pdf;
    pt_bsdf_eval(bsdf, wo, wi, &f, &pdf);
    if (!(pdf > 0.0f)) return false;
    float g = pt_div(wi.z, pdf);
    path->beta = pt_mk(path->beta.x * f.x * g, path->beta.y * f.y * g, path->beta.z * f.z * g);
    if (path->bounce >= 2) {  // Russian roulette
        float q = pt_min(pt_maxcomp(path->beta), 0.95f);
        if (!(u6 < q)) return false;
        float iq = pt_div(1.0f, q);
        path->beta = pt_scale(path->beta, iq);
    }
    if (!(pt_maxcomp(path->beta) > 0.0f)) return false;
    path->o = po;
    path->d = pt_normalize(pt_madd(pt_madd(pt_scale(T, wi.x), wi.y, B), wi.z, n));
    path->pdf_prev = pdf;
    path->bounce += 1;
    return true;
}

PT_HD void pt_shade_miss(PtPath* path, const PtShadeConsts& sc) {
    path->L = pt_mk(pt_fma(path->beta.x, sc.bg[0], path->L.x), pt_fma(path->beta.y, sc.bg[1], path->L.y), pt_fma(path->beta.z, sc.bg[2], path->L.z));
}

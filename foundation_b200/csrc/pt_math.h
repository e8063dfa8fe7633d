// pt_math.h — the arithmetic contract of the path-tracing hot path.
//
// Everything that decides a hit ID, a hit distance, a sampled direction or a radiance value is
// written here ONCE, as __host__ __device__ functions built only from IEEE-754 correctly-rounded
// primitives: + - * / sqrt and explicit fused multiply-add.  No libm transcendentals, no implicit
// contraction (the CUDA TUs are compiled with --fmad=false, the CPU TUs with -ffp-contract=off), so
// an sm_100a thread and an x86 core evaluate the SAME sequence of roundings.  That is what makes the
// north_star's "bit-exact hit IDs, t within 2 ulp" achievable (we get 0 ulp) — see SURVEY.md §7
// "Hard parts".  The reference (mos9527/Foundation) has no counterpart for any of this: its only
// arithmetic on the path is the glm camera maths at src/Renderer/Renderer.cpp:373-380, which
// arrives here as data (view/proj matrices).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#define PT_HDM __host__ __device__ __forceinline__
#else
#define PT_HD static inline __attribute__((always_inline))
#define PT_HDM inline
#endif

// ---------------------------------------------------------------------------------------------
// primitives
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
PT_HD float pt_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
PT_HD float pt_sqrt(float a) { return __fsqrt_rn(a); }
PT_HD float pt_div(float a, float b) { return __fdiv_rn(a, b); }
PT_HD uint32_t pt_f2u(float f) { return __float_as_uint(f); }
PT_HD float pt_u2f(uint32_t u) { return __uint_as_float(u); }
PT_HD int pt_clz64(uint64_t x) { return __clzll((long long)x); }
PT_HD int pt_clz32(uint32_t x) { return __clz((int)x); }
PT_HD int pt_popc(uint32_t x) { return __popc(x); }
PT_HD int pt_ffs0(uint32_t x) { return __ffs((int)x) - 1; }  // index of lowest set bit (x != 0)
#else
PT_HD float pt_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
PT_HD float pt_sqrt(float a) { return __builtin_sqrtf(a); }
PT_HD float pt_div(float a, float b) { return a / b; }
PT_HD uint32_t pt_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
PT_HD float pt_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
PT_HD int pt_clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
PT_HD int pt_clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
PT_HD int pt_popc(uint32_t x) { return __builtin_popcount(x); }
PT_HD int pt_ffs0(uint32_t x) { return __builtin_ctz(x); }
#endif

// min/max with a fixed NaN rule (result is b when the comparison is false), identical on both sides.
PT_HD float pt_min(float a, float b) { return a < b ? a : b; }
PT_HD float pt_max(float a, float b) { return a > b ? a : b; }
PT_HD float pt_abs(float a) { return pt_u2f(pt_f2u(a) & 0x7fffffffu); }
PT_HD float pt_copysign(float mag, float sgn) {
    return pt_u2f((pt_f2u(mag) & 0x7fffffffu) | (pt_f2u(sgn) & 0x80000000u));
}
PT_HD float pt_clamp(float x, float lo, float hi) { return pt_min(pt_max(x, lo), hi); }

#define PT_INF_BITS 0x7f800000u
#define PT_PI 3.14159274101257324f        // float(pi)
#define PT_INV_PI 0.318309873342514038f   // float(1/pi)
#define PT_NONE 0xffffffffu

struct pt_v3 { float x, y, z; };

PT_HD pt_v3 pt_mk(float x, float y, float z) { pt_v3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_HD pt_v3 pt_add(pt_v3 a, pt_v3 b) { return pt_mk(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD pt_v3 pt_sub(pt_v3 a, pt_v3 b) { return pt_mk(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD pt_v3 pt_mul(pt_v3 a, pt_v3 b) { return pt_mk(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD pt_v3 pt_scale(pt_v3 a, float s) { return pt_mk(a.x * s, a.y * s, a.z * s); }
PT_HD pt_v3 pt_neg(pt_v3 a) { return pt_mk(-a.x, -a.y, -a.z); }
// a + s*b, one rounding per component
PT_HD pt_v3 pt_madd(pt_v3 a, float s, pt_v3 b) {
    return pt_mk(pt_fma(s, b.x, a.x), pt_fma(s, b.y, a.y), pt_fma(s, b.z, a.z));
}
PT_HD float pt_dot(pt_v3 a, pt_v3 b) { return pt_fma(a.x, b.x, pt_fma(a.y, b.y, a.z * b.z)); }
PT_HD pt_v3 pt_cross(pt_v3 a, pt_v3 b) {
    return pt_mk(pt_fma(a.y, b.z, -(a.z * b.y)), pt_fma(a.z, b.x, -(a.x * b.z)), pt_fma(a.x, b.y, -(a.y * b.x)));
}
PT_HD pt_v3 pt_normalize(pt_v3 a) {
    float l = pt_sqrt(pt_dot(a, a));
    float inv = pt_div(1.0f, l);
    return pt_scale(a, inv);
}
PT_HD float pt_maxcomp(pt_v3 a) { return pt_max(a.x, pt_max(a.y, a.z)); }

// ---------------------------------------------------------------------------------------------
// ray / triangle.  Moller-Trumbore on the (v0, e1, e2) form, two-sided, decided on UNDIVIDED
// barycentrics so the accept/reject decision never depends on a reciprocal.
//   accept  <=>  |det| > 0,  U >= 0,  V >= 0,  U + V <= |det|      (U,V sign-normalised by det)
//   t = T / |det|   (one correctly rounded division, only evaluated for accepted candidates)
// Closest-hit order is lexicographic on (t, id): ties on shared edges / coplanar duplicates resolve
// to the smaller primitive id, independent of traversal order or BVH shape.
// ---------------------------------------------------------------------------------------------
PT_HD bool pt_ray_tri(pt_v3 o, pt_v3 d, pt_v3 v0, pt_v3 e1, pt_v3 e2, float* t, float* U, float* V, float* ad) {
    pt_v3 p = pt_cross(d, e2);
    float det = pt_dot(e1, p);
    pt_v3 tv = pt_sub(o, v0);
    float u = pt_dot(tv, p);
    pt_v3 q = pt_cross(tv, e1);
    float v = pt_dot(d, q);
    uint32_t s = pt_f2u(det) & 0x80000000u;
    float a = pt_u2f(pt_f2u(det) ^ s);
    float un = pt_u2f(pt_f2u(u) ^ s);
    float vn = pt_u2f(pt_f2u(v) ^ s);
    if (!(a > 0.0f) || !(un >= 0.0f) || !(vn >= 0.0f) || !(un + vn <= a)) return false;
    float tn = pt_u2f(pt_f2u(pt_dot(e2, q)) ^ s);
    *t = pt_div(tn, a);
    *U = un; *V = vn; *ad = a;
    return true;
}

// lexicographic (t, id) improvement test against the current best (best_id == PT_NONE: no hit yet,
// best_t is then the ray's tmax, which is exclusive).
PT_HD bool pt_closer(float t, uint64_t id, float tmin, float best_t, uint64_t best_id, bool have_best) {
    if (!(t > tmin)) return false;
    if (t < best_t) return true;
    return have_best && t == best_t && id < best_id;
}

// ---------------------------------------------------------------------------------------------
// Morton keys: 21 bits per axis interleaved to 63 bits (x highest).
// ---------------------------------------------------------------------------------------------
PT_HD uint64_t pt_expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x001f00000000ffffull;
    x = (x | (x << 16)) & 0x001f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
PT_HD uint32_t pt_quant21(float c, float lo, float inv_ext) {
    // (c - lo) * (2^21 / extent), clamped to [0, 2^21-1]; NaN -> 0
    float f = (c - lo) * inv_ext;
    if (!(f > 0.0f)) return 0u;
    if (f >= 2097151.0f) return 2097151u;
    return (uint32_t)f;
}
PT_HD uint64_t pt_morton63(pt_v3 c, pt_v3 lo, pt_v3 inv_ext) {
    return (pt_expand21(pt_quant21(c.x, lo.x, inv_ext.x)) << 2) | (pt_expand21(pt_quant21(c.y, lo.y, inv_ext.y)) << 1) |
           pt_expand21(pt_quant21(c.z, lo.z, inv_ext.z));
}
// centroid used for the Morton key of a triangle: ((v0+v1)+v2) * (1/3 rounded to float)
PT_HD pt_v3 pt_tri_centroid(pt_v3 a, pt_v3 b, pt_v3 c) {
    return pt_scale(pt_add(pt_add(a, b), c), 0.333333343267440796f);
}
// Karras-2012 delta: common-prefix length of keys i and j with the index as tie-break.
PT_HD int pt_delta(uint64_t ki, uint64_t kj, uint32_t i, uint32_t j) {
    uint64_t x = ki ^ kj;
    if (x) return pt_clz64(x);
    return 64 + pt_clz32(i ^ j);
}

// ---------------------------------------------------------------------------------------------
// PCG32 (XSH-RR 64/32), O'Neill 2014.  One stream per (pixel, sample).
// ---------------------------------------------------------------------------------------------
struct pt_rng { uint64_t state, inc; };
#define PT_PCG_MULT 6364136223846793005ull
PT_HD uint32_t pt_rng_next(pt_rng* r) {
    uint64_t old = r->state;
    r->state = old * PT_PCG_MULT + r->inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((32u - rot) & 31u));
}
PT_HD pt_rng pt_rng_seed(uint64_t initstate, uint64_t initseq) {
    pt_rng r; r.state = 0u; r.inc = (initseq << 1u) | 1u;
    pt_rng_next(&r); r.state += initstate; pt_rng_next(&r);
    return r;
}
// uniform float in [0,1): top 24 bits * 2^-24 (exact)
PT_HD float pt_rng_f(pt_rng* r) { return (float)(pt_rng_next(r) >> 8) * 5.9604644775390625e-08f; }
// splitmix64 finaliser: decorrelates (seed, pixel, sample) into a PCG initstate
PT_HD uint64_t pt_mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
PT_HD pt_rng pt_rng_for(uint64_t seed, uint32_t pixel, uint32_t sample) {
    uint64_t h = pt_mix64(seed ^ pt_mix64(((uint64_t)sample << 32) | pixel));
    return pt_rng_seed(h, ((uint64_t)sample << 32) | pixel);
}

// ---------------------------------------------------------------------------------------------
// sincos(2*pi*u), u in [0,1): quadrant reduction (exact) + cephes-style minimax polynomials on
// [-pi/4, pi/4], every step an explicit fma/mul so both machines agree to the bit.  |err| < 2e-7.
// ---------------------------------------------------------------------------------------------
PT_HD void pt_sincos2pi(float u, float* s, float* c) {
    float a = u * 4.0f;                         // exact
    int q = (int)(a + 0.5f);                    // nearest quadrant 0..4
    float r = a - (float)q;                     // exact, in [-0.5, 0.5]
    float x = r * 1.57079637050628662f;         // * float(pi/2)
    float x2 = x * x;
    float sp = pt_fma(x2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = pt_fma(x2, sp, -1.6666654611e-1f);
    float sn = pt_fma(x * x2, sp, x);
    float cp = pt_fma(x2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = pt_fma(x2, cp, 4.166664568298827e-2f);
    float cs = pt_fma(x2 * x2, cp, pt_fma(x2, -0.5f, 1.0f));
    switch (q & 3) {
        case 0: *s = sn; *c = cs; break;
        case 1: *s = cs; *c = -sn; break;
        case 2: *s = -sn; *c = -cs; break;
        default: *s = -cs; *c = sn; break;
    }
}

// Orthonormal basis around unit n (Duff et al. 2017), branch-free in the arithmetic.
PT_HD void pt_onb(pt_v3 n, pt_v3* t, pt_v3* b) {
    float sg = pt_copysign(1.0f, n.z);
    float a = pt_div(-1.0f, sg + n.z);
    float bb = n.x * n.y * a;
    *t = pt_mk(pt_fma(sg * n.x, n.x * a, 1.0f), sg * bb, -sg * n.x);
    *b = pt_mk(bb, pt_fma(n.y, n.y * a, sg), -n.y);
}

// 3x4 affine transform (row-major rows r0,r1,r2 = [m00 m01 m02 tx]) applied to points / vectors.
PT_HD pt_v3 pt_xform_point(const float* m, pt_v3 p) {
    return pt_mk(pt_fma(m[0], p.x, pt_fma(m[1], p.y, pt_fma(m[2], p.z, m[3]))),
                 pt_fma(m[4], p.x, pt_fma(m[5], p.y, pt_fma(m[6], p.z, m[7]))),
                 pt_fma(m[8], p.x, pt_fma(m[9], p.y, pt_fma(m[10], p.z, m[11]))));
}
PT_HD pt_v3 pt_xform_vec(const float* m, pt_v3 v) {
    return pt_mk(pt_fma(m[0], v.x, pt_fma(m[1], v.y, m[2] * v.z)), pt_fma(m[4], v.x, pt_fma(m[5], v.y, m[6] * v.z)),
                 pt_fma(m[8], v.x, pt_fma(m[9], v.y, m[10] * v.z)));
}

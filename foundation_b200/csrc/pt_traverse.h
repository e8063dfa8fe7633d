// pt_traverse.h — per-ray traversal of the compressed BVH8 (stages B2 "extend" and B5 "connect" of
// SURVEY.md §8a2), written as __host__ __device__ templates so the exact device control flow can also be
// executed by the CPU-side emulation harness under tests/ (never by the product: the product only ever
// instantiates these inside __global__ kernels).
//
// Algorithm: stack of (base, mask) groups as in Ylitie, Karras, Laine, "Efficient Incoherent Ray
// Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017): a node group holds the still-unvisited
// internal children of one node ordered front-to-back by `slot ^ ray_octant`, a triangle group holds the
// still-untested triangles of one node.  Results are order independent: the closest hit is the
// lexicographic minimum of (t, instance, primitive) and box culling is conservative (child boxes are
// padded and rounded outward at build time, see pt_layout.h).
// No reference counterpart exists (the reference's RHI has no ray query / AS, src/Platform/RHI/Command.hpp:38-115).
#pragma once
#include "pt_layout.h"

#if defined(__CUDA_ARCH__)
PT_HD float pt_fmin(float a, float b) { return fminf(a, b); }
PT_HD float pt_fmax(float a, float b) { return fmaxf(a, b); }
#else
PT_HD float pt_fmin(float a, float b) { return __builtin_fminf(a, b); }
PT_HD float pt_fmax(float a, float b) { return __builtin_fmaxf(a, b); }
#endif

#define PT_STACK_SIZE 64
#define PT_IDIR_CLAMP 9.094947017729282e-13f  // 2^-40

struct PtU4 { uint32_t x, y, z, w; };
struct PtU2 { uint32_t x, y; };

// The whole scene as the traversal sees it.
struct PtSceneView {
    const PtU4* nodes;        // all BVH8 nodes: [TLAS | BLAS 0 | BLAS 1 ...] when two_level, else one BLAS
    const PtU4* tris;         // all triangles, 3 x 16 B each, BLAS after BLAS, leaf order
    const PtU4* instances;    // PtInstance records in TLAS leaf order, 7 x 16 B each (two_level only)
};

struct PtRayCtx {
    pt_v3 o, d, idir;
    uint32_t oct_inv;  // bit 2/1/0 set when d.x/d.y/d.z is non-negative
};

PT_HD float pt_safe_rcp_dir(float d) {
    float a = pt_abs(d) < PT_IDIR_CLAMP ? pt_copysign(PT_IDIR_CLAMP, d) : d;
    return pt_div(1.0f, a);
}
PT_HD void pt_ray_ctx(PtRayCtx* c, pt_v3 o, pt_v3 d) {
    c->o = o; c->d = d;
    c->idir = pt_mk(pt_safe_rcp_dir(d.x), pt_safe_rcp_dir(d.y), pt_safe_rcp_dir(d.z));
    c->oct_inv = ((pt_f2u(d.x) >> 31) ? 0u : 4u) | ((pt_f2u(d.y) >> 31) ? 0u : 2u) | ((pt_f2u(d.z) >> 31) ? 0u : 1u);
}

PT_HD uint32_t pt_byte(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

// Intersects the 8 quantised child boxes of one node.  Returns the hit mask: bits 24..31 = internal
// children that are hit, at position 24 + (slot ^ oct_inv); bits 0..23 = triangles of leaf slots that are hit.
PT_HD uint32_t pt_node_hits(const PtU4& n0, const PtU4& n1, const PtU4& n2, const PtU4& n3, const PtU4& n4, const PtRayCtx& r, float tmin,
                            float tbest) {
    float sx = pt_u2f((n0.w & 0xffu) << 23), sy = pt_u2f(((n0.w >> 8) & 0xffu) << 23), sz = pt_u2f(((n0.w >> 16) & 0xffu) << 23);
    float ax = sx * r.idir.x, ay = sy * r.idir.y, az = sz * r.idir.z;
    float bx = (pt_u2f(n0.x) - r.o.x) * r.idir.x, by = (pt_u2f(n0.y) - r.o.y) * r.idir.y, bz = (pt_u2f(n0.z) - r.o.z) * r.idir.z;
    bool negx = !(r.oct_inv & 4u), negy = !(r.oct_inv & 2u), negz = !(r.oct_inv & 1u);
    uint32_t mask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int half = 0; half < 2; ++half) {
        uint32_t meta4 = half ? n1.w : n1.z;
        uint32_t qlx = half ? n2.y : n2.x, qly = half ? n2.w : n2.z, qlz = half ? n3.y : n3.x;
        uint32_t qhx = half ? n3.w : n3.z, qhy = half ? n4.y : n4.x, qhz = half ? n4.w : n4.z;
        uint32_t nx = negx ? qhx : qlx, fx = negx ? qlx : qhx;
        uint32_t ny = negy ? qhy : qly, fy = negy ? qly : qhy;
        uint32_t nz = negz ? qhz : qlz, fz = negz ? qlz : qhz;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) {
            uint32_t meta = pt_byte(meta4, j);
            float tnx = pt_fma((float)pt_byte(nx, j), ax, bx), tfx = pt_fma((float)pt_byte(fx, j), ax, bx);
            float tny = pt_fma((float)pt_byte(ny, j), ay, by), tfy = pt_fma((float)pt_byte(fy, j), ay, by);
            float tnz = pt_fma((float)pt_byte(nz, j), az, bz), tfz = pt_fma((float)pt_byte(fz, j), az, bz);
            float tn = pt_fmax(pt_fmax(tnx, tny), pt_fmax(tnz, tmin));
            float tf = pt_fmin(pt_fmin(tfx, tfy), pt_fmin(tfz, tbest));
            if (tn <= tf) {
                uint32_t bits = meta >> 5;
                uint32_t inner = ((meta & 0x18u) == 0x18u) ? 7u : 0u;   // internal children carry 24 + slot
                uint32_t idx = (meta & 0x1fu) ^ (r.oct_inv & inner);
                mask |= bits << idx;
            }
        }
    }
    return mask;
}

struct PtHitRec {
    float t, U, V, ad;   // undivided barycentrics U, V and |det|
    uint32_t prim, inst;
    uint32_t tidx, iidx;  // position of the triangle / instance record in the leaf-ordered device arrays (for shading)
};

template <class Counter>
PT_HD void pt_test_tri(const PtU4* tris, uint32_t tri_index, const PtRayCtx& r, float tmin, uint32_t inst, uint32_t iidx, PtHitRec* best, Counter& cnt) {
    const PtU4 a = tris[3 * (size_t)tri_index + 0], b = tris[3 * (size_t)tri_index + 1], c = tris[3 * (size_t)tri_index + 2];
    cnt.tri();
    float t, U, V, ad;
    if (pt_ray_tri(r.o, r.d, pt_mk(pt_u2f(a.x), pt_u2f(a.y), pt_u2f(a.z)), pt_mk(pt_u2f(b.x), pt_u2f(b.y), pt_u2f(b.z)),
                   pt_mk(pt_u2f(c.x), pt_u2f(c.y), pt_u2f(c.z)), &t, &U, &V, &ad)) {
        uint64_t id = ((uint64_t)inst << 32) | a.w, bid = ((uint64_t)best->inst << 32) | best->prim;
        if (pt_closer(t, id, tmin, best->t, bid, best->prim != PT_NONE)) {
            best->t = t; best->U = U; best->V = V; best->ad = ad; best->prim = a.w; best->inst = inst;
            best->tidx = tri_index; best->iidx = iidx;
        }
    }
}

struct PtNoCount { PT_HDM void node() {} PT_HDM void tri() {} PT_HDM void inst() {} };
struct PtCount {
    uint64_t nodes, tris, insts;
    PT_HDM void node() { ++nodes; } PT_HDM void tri() { ++tris; } PT_HDM void inst() { ++insts; }
};

// Resumable traversal state: pt_trav_init + repeated pt_trav_step (one node visit, its triangles, one pop).
// The kernels interleave steps with warp-level dynamic ray fetch; pt_traverse below is init + loop and is what
// the emulation harness runs.
struct PtTravState {
    PtRayCtx world, r;
    float tmin;
    PtU2 ng, tg;
    uint32_t node_base, tri_base, cur_inst, cur_iidx;
    int sp;
    bool in_blas, overflow;
};   // the group stack is a separate array so this struct stays in registers
enum { PT_STEP_RUNNING = 0, PT_STEP_DONE = 1 };

template <bool TWO_LEVEL>
PT_HD void pt_trav_init(PtTravState* s, pt_v3 o, pt_v3 d, float tmin, float tmax, PtHitRec* best) {
    best->t = tmax; best->U = 0.0f; best->V = 0.0f; best->ad = 1.0f; best->prim = PT_NONE; best->inst = PT_NONE;
    best->tidx = 0; best->iidx = 0;
    pt_ray_ctx(&s->world, o, d);
    s->r = s->world;
    s->tmin = tmin;
    s->sp = 0; s->overflow = false;
    s->in_blas = !TWO_LEVEL;
    s->node_base = 0; s->tri_base = 0; s->cur_inst = TWO_LEVEL ? PT_NONE : 0u; s->cur_iidx = 0;
    // root: one pending child of a virtual parent with child_base 0 and an empty imask, so popc(...) = 0 -> node 0
    s->ng.x = 0; s->ng.y = 0x80000000u;
    s->tg.x = 0; s->tg.y = 0;
}

// ANY = true: occlusion query, finishes as soon as any triangle is hit in (tmin, tmax).
template <bool ANY, bool TWO_LEVEL, class Counter>
PT_HD int pt_trav_step(const PtSceneView& sc, PtTravState* s, PtU2* stack, PtHitRec* best, Counter& cnt) {
    if (s->ng.y & 0xff000000u) {
        uint32_t bit = 31u - (uint32_t)pt_clz32(s->ng.y);
        s->ng.y &= ~(1u << bit);
        uint32_t slot = (bit - 24u) ^ s->r.oct_inv;
        uint32_t child = s->ng.x + (uint32_t)pt_popc(s->ng.y & 0xffu & ~(0xffffffffu << slot));
        if (s->ng.y & 0xff000000u) {
            if (s->sp >= PT_STACK_SIZE) { s->overflow = true; return PT_STEP_DONE; }
            stack[s->sp++] = s->ng;
        }
        const PtU4* np = sc.nodes + 5 * (size_t)(s->node_base + child);
        const PtU4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3], n4 = np[4];
        cnt.node();
        uint32_t hits = pt_node_hits(n0, n1, n2, n3, n4, s->r, s->tmin, best->t);
        s->ng.x = n1.x; s->ng.y = (hits & 0xff000000u) | (n0.w >> 24);
        s->tg.x = n1.y; s->tg.y = hits & 0x00ffffffu;
    } else {
        s->tg = s->ng; s->ng.x = 0; s->ng.y = 0;
    }
    while (s->tg.y) {
        uint32_t k = (uint32_t)pt_ffs0(s->tg.y);
        s->tg.y &= s->tg.y - 1u;
        if (!TWO_LEVEL || s->in_blas) {
            pt_test_tri(sc.tris, s->tri_base + s->tg.x + k, s->r, s->tmin, s->cur_inst, s->cur_iidx, best, cnt);
            if (ANY && best->prim != PT_NONE) return PT_STEP_DONE;
        } else {
            // TLAS leaf: enter the instance.  Save the remaining groups, push the return sentinel.
            if (s->sp + 3 > PT_STACK_SIZE) { s->overflow = true; return PT_STEP_DONE; }
            if (s->tg.y) stack[s->sp++] = s->tg;
            if (s->ng.y & 0xff000000u) stack[s->sp++] = s->ng;
            PtU2 sentinel; sentinel.x = PT_NONE; sentinel.y = 0; stack[s->sp++] = sentinel;
            const PtU4* ip = sc.instances + 7 * (size_t)(s->tg.x + k);
            PtU4 m0 = ip[0], m1 = ip[1], m2 = ip[2], m6 = ip[6];
            float w2o[12] = {pt_u2f(m0.x), pt_u2f(m0.y), pt_u2f(m0.z), pt_u2f(m0.w), pt_u2f(m1.x), pt_u2f(m1.y),
                             pt_u2f(m1.z), pt_u2f(m1.w), pt_u2f(m2.x), pt_u2f(m2.y), pt_u2f(m2.z), pt_u2f(m2.w)};
            cnt.inst();
            pt_ray_ctx(&s->r, pt_xform_point(w2o, s->world.o), pt_xform_vec(w2o, s->world.d));
            s->node_base = m6.x; s->tri_base = m6.y; s->cur_inst = m6.w; s->cur_iidx = s->tg.x + k;   // m6 = node_base, tri_base, mesh_id, inst_id
            s->in_blas = true;
            s->ng.x = 0; s->ng.y = 0x80000000u;
            s->tg.x = 0; s->tg.y = 0;
        }
    }
    if (!(s->ng.y & 0xff000000u)) {
        for (;;) {
            if (s->sp == 0) return PT_STEP_DONE;
            s->ng = stack[--s->sp];
            if (TWO_LEVEL && s->ng.x == PT_NONE && s->ng.y == 0) {   // leaving an instance
                s->r = s->world; s->in_blas = false; s->node_base = 0; s->tri_base = 0; s->cur_inst = PT_NONE;
                continue;
            }
            break;
        }
    }
    return PT_STEP_RUNNING;
}

// Whole traversal of one ray.  Returns false on traversal-stack overflow (reported through the status word).
template <bool ANY, bool TWO_LEVEL, class Counter>
PT_HD bool pt_traverse(const PtSceneView& sc, pt_v3 o, pt_v3 d, float tmin, float tmax, PtHitRec* best, Counter& cnt) {
    PtTravState s;
    PtU2 stack[PT_STACK_SIZE];
    pt_trav_init<TWO_LEVEL>(&s, o, d, tmin, tmax, best);
    while (pt_trav_step<ANY, TWO_LEVEL>(sc, &s, stack, best, cnt) == PT_STEP_RUNNING) {}
    return !s.overflow;
}

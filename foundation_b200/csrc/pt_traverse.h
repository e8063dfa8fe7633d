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

struct alignas(16) PtU4 { uint32_t x, y, z, w; };   // 16-byte aligned so node / triangle words move as single 128-bit loads
struct alignas(8) PtU2 { uint32_t x, y; };

#if defined(__CUDA_ARCH__)
PT_HD PtU4 pt_load4(const PtU4* p) {   // read-only 128-bit load (LDG.E.128.CONSTANT)
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    PtU4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
#else
PT_HD PtU4 pt_load4(const PtU4* p) { return *p; }
#endif

// The whole scene as the traversal sees it.
struct PtSceneView {
    const PtU4* nodes;        // all BVH8 nodes: [BLAS 0 | BLAS 1 ... | TLAS] when two_level (any order works: bases are explicit), else one BLAS
    const PtU4* tris;         // all triangles, 3 x 16 B each, BLAS after BLAS, leaf order
    const PtU4* instances;    // PtInstance records in TLAS leaf order, 7 x 16 B each (two_level only)
    uint32_t tlas_base;       // index of the TLAS root in `nodes` (the TLAS sits after the BLAS so it can be rebuilt alone)
    uint32_t zero;            // always 0, but only known at run time: see pt_test_tri_words
    uint32_t qbias = PT_QBIAS_BITS;   // float bits of PT_QBIAS as run-time data: kept in ONE register it serves all 48 byte permutes of a node
                                      // test with immediate selectors (as a literal, ptxas puts it in the immediate slot and moves the selectors
                                      // through registers instead: +36 instructions per node)
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
// byte i of w as the float PT_QBIAS + q: q lands in mantissa bits 8..15 of 2^15 (0x47000000), exact for q in 0..255
#if defined(__CUDA_ARCH__)
PT_HD float pt_qfloat(uint32_t w, int i, uint32_t qbias) { return __uint_as_float(__byte_perm(w, qbias, 0x7404u | ((uint32_t)i << 4))); }
#else
PT_HD float pt_qfloat(uint32_t w, int i, uint32_t qbias) { return pt_u2f(qbias | (pt_byte(w, i) << 8)); }
#endif

#ifndef PT_NODE_LOADS_LATE
#define PT_NODE_LOADS_LATE 0
#endif
#ifndef PT_NODE_F32X2
#define PT_NODE_F32X2 1      // 24 FFMA2 instead of 48 FFMA per node test.  Same-box A/B (profiles/r02_ab_traversal_build.log): +1.9 % spp/s, +2.5 % any-hit,
                             // -1 % on the closest-hit ray sets: the node test is bound by the ALU pipe (PRMT / FMNMX / LOP3) and by load latency, not by the fma pipe
#endif
#if defined(__CUDA_ARCH__)
// {lo, hi} = {q0 * a0 + c0, q1 * a1 + c1} in ONE instruction (fma.rn.f32x2, sm_100+)
__device__ __forceinline__ void pt_fma2(float q0, float q1, float a0, float a1, float c0, float c1, float* lo, float* hi) {
    unsigned long long q, m, c, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(q0), "f"(q1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(q), "l"(m), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(*lo), "=f"(*hi) : "l"(r));
}
#endif

// Intersects the 8 quantised child boxes of one node.  Returns the hit mask: bits 24..31 = internal
// children that are hit, at position 24 + (slot ^ oct_inv); bits 0..23 = triangles of leaf slots that are hit.
PT_HD uint32_t pt_node_hits(const PtU4& n0, const PtU4& n1, const PtU4& n2, const PtU4& n3, const PtU4& n4, const PtRayCtx& r, float tmin,
                            float tbest, uint32_t qbias) {
    float sx = pt_u2f((n0.w & 0xffu) << 23), sy = pt_u2f(((n0.w >> 8) & 0xffu) << 23), sz = pt_u2f(((n0.w >> 16) & 0xffu) << 23);
    float ax = sx * r.idir.x, ay = sy * r.idir.y, az = sz * r.idir.z;
    float bx = (pt_u2f(n0.x) - r.o.x) * r.idir.x, by = (pt_u2f(n0.y) - r.o.y) * r.idir.y, bz = (pt_u2f(n0.z) - r.o.z) * r.idir.z;
    // near planes move towards the origin, far planes away from it, by the rounding-error bound of the plane distance (see PT_SLAB_EPS).
    // The quantised byte q enters the fma as the float PT_QBIAS + q (pt_qfloat: one byte permute, no int->float conversion, which
    // is a quarter-rate XU instruction and was the busiest pipe of the kernel); the bias is folded into the constant term.
    float ex = pt_fma(pt_abs(ax), PT_SLAB_QMAX, pt_abs(bx)) * PT_SLAB_EPS, ey = pt_fma(pt_abs(ay), PT_SLAB_QMAX, pt_abs(by)) * PT_SLAB_EPS,
          ez = pt_fma(pt_abs(az), PT_SLAB_QMAX, pt_abs(bz)) * PT_SLAB_EPS;
    float cx = pt_fma(-PT_QBIAS, ax, bx), cy = pt_fma(-PT_QBIAS, ay, by), cz = pt_fma(-PT_QBIAS, az, bz);
    float bnx = cx - ex, bny = cy - ey, bnz = cz - ez, bfx = cx + ex, bfy = cy + ey, bfz = cz + ez;
    bool negx = !(r.oct_inv & 4u), negy = !(r.oct_inv & 2u), negz = !(r.oct_inv & 1u);
    uint32_t mask = 0;
    const uint32_t oct4 = r.oct_inv * 0x01010101u;
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
        // byte-parallel decode of the four meta bytes (no per-child branches): internal children (bits 3 and 4 set) get
        // their bit position 24 + (slot ^ oct), leaves keep their triangle offset; bits = 1 / unary triangle count
        uint32_t inner4 = (((meta4 & (meta4 << 1)) & 0x10101010u) >> 4) * 0xffu;
        uint32_t idx4 = (meta4 ^ (oct4 & inner4)) & 0x1f1f1f1fu;
        uint32_t bits4 = (meta4 >> 5) & 0x07070707u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) {
#if defined(__CUDA_ARCH__) && PT_NODE_F32X2
            // sm_100 packed fp32 (FFMA2): {near x, near y} and {far x, far y} share the multiplier pair {ax, ay}, {near z, far z} the pair
            // {az, az} — four multiplier registers instead of three (pairing near / far on every axis needs six and spills at the
            // 64-register cap).  Each half is an IEEE round-to-nearest fma of its own operands: the same bits as the scalar fmas (and the oracle's).
            float tnx, tfx, tny, tfy, tnz, tfz;
            pt_fma2(pt_qfloat(nx, j, qbias), pt_qfloat(ny, j, qbias), ax, ay, bnx, bny, &tnx, &tny);
            pt_fma2(pt_qfloat(fx, j, qbias), pt_qfloat(fy, j, qbias), ax, ay, bfx, bfy, &tfx, &tfy);
            pt_fma2(pt_qfloat(nz, j, qbias), pt_qfloat(fz, j, qbias), az, az, bnz, bfz, &tnz, &tfz);
#else
            float tnx = pt_fma(pt_qfloat(nx, j, qbias), ax, bnx), tfx = pt_fma(pt_qfloat(fx, j, qbias), ax, bfx);
            float tny = pt_fma(pt_qfloat(ny, j, qbias), ay, bny), tfy = pt_fma(pt_qfloat(fy, j, qbias), ay, bfy);
            float tnz = pt_fma(pt_qfloat(nz, j, qbias), az, bnz), tfz = pt_fma(pt_qfloat(fz, j, qbias), az, bfz);
#endif
            float tn = pt_fmax(pt_fmax(tnx, tny), pt_fmax(tnz, tmin));
            float tf = pt_fmin(pt_fmin(tfx, tfy), pt_fmin(tfz, tbest));
            uint32_t sel = (tn <= tf) ? 0xffffffffu : 0u;
            mask |= (pt_byte(bits4, j) << pt_byte(idx4, j)) & sel;
        }
    }
    return mask;
}

#ifndef PT_SLIM_HIT
#define PT_SLIM_HIT 0      // 1: barycentrics / material id re-derived once per ray instead of carried through the loop (4 registers less).  Measured
                           // (profiles/r02_ab_traversal_build.log, run r2g): -1.8 % closest-hit, -1.4 % spp/s at 8 CTAs/SM (the extra triangle fetch per hitting ray costs
                           // more than the registers buy), and 9 CTAs/SM at 56 registers still spills 12-44 bytes: -7 %.  Off.
#endif
struct PtHitRec {
    float t;
#if !PT_SLIM_HIT
    float U, V, ad;   // undivided barycentrics U, V and |det|
#endif
    uint32_t prim, inst;
    uint32_t tidx, iidx;  // position of the triangle / instance record in the leaf-ordered device arrays (for shading, and for pt_hit_bary)
#if !PT_SLIM_HIT
    uint32_t mat;         // material id of the hit triangle (word 1 .w of the record)
#endif
};   // PT_SLIM_HIT: the barycentrics and the material id of the FINAL hit are re-derived once per ray from (tidx, iidx) — pt_hit_bary, one extra
     // triangle fetch per hitting ray — instead of being carried through the traversal loop: four registers less in a loop that sits at the cap

// one ray / triangle test against the current best; a, b, c are the triangle's three 16-byte words
// `keep` is PtSceneView::zero (0 at run time, opaque to the compiler): OR-ing the two unused .w lanes of the triangle words into
// prim through it keeps those registers LIVE until the hit update.  Without it ptxas recycles the dead lane of the 128-bit load as
// a scratch register for the very next ALU instruction, which then waits (write-after-write) for the load to land BEFORE the
// node loads of the same step are issued — the triangle and node round trips serialise again and the kernel loses 8 %
// (measured on the same box: 6188 vs 6728 Mrays/s; a SASS scan for such hazards is in scripts/sass_waw.py).
template <class Counter>
PT_HD void pt_test_tri_words(const PtU4& a, const PtU4& b, const PtU4& c, uint32_t tri_index, const PtRayCtx& r, float tmin, uint32_t inst, uint32_t iidx,
                             PtHitRec* best, Counter& cnt, uint32_t keep = 0) {
    cnt.tri();
    float t, U, V, ad;
    if (pt_ray_tri(r.o, r.d, pt_mk(pt_u2f(a.x), pt_u2f(a.y), pt_u2f(a.z)), pt_mk(pt_u2f(b.x), pt_u2f(b.y), pt_u2f(b.z)),
                   pt_mk(pt_u2f(c.x), pt_u2f(c.y), pt_u2f(c.z)), &t, &U, &V, &ad)) {
        uint64_t id = ((uint64_t)inst << 32) | a.w, bid = ((uint64_t)best->inst << 32) | best->prim;
        if (pt_closer(t, id, tmin, best->t, bid, best->prim != PT_NONE)) {
            best->t = t; best->prim = a.w | ((b.w | c.w) & keep); best->inst = inst;
            best->tidx = tri_index; best->iidx = iidx;
#if !PT_SLIM_HIT
            best->U = U; best->V = V; best->ad = ad; best->mat = b.w;
#endif
        }
    }
}
template <class Counter>
PT_HD void pt_test_tri(const PtU4* tris, uint32_t tri_index, const PtRayCtx& r, float tmin, uint32_t inst, uint32_t iidx, PtHitRec* best, Counter& cnt) {
    const PtU4 a = pt_load4(tris + 3 * (size_t)tri_index), b = pt_load4(tris + 3 * (size_t)tri_index + 1), c = pt_load4(tris + 3 * (size_t)tri_index + 2);
    pt_test_tri_words(a, b, c, tri_index, r, tmin, inst, iidx, best, cnt, 0u);
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
    int sp;            // stack entries in use; -1 after a stack overflow (the ray is abandoned, the kernels report it through the status word)
    bool in_blas;
};   // the group stack is a separate array so this struct stays in registers
enum { PT_STEP_RUNNING = 0, PT_STEP_DONE = 1 };

template <bool TWO_LEVEL>
PT_HD void pt_trav_init(PtTravState* s, pt_v3 o, pt_v3 d, float tmin, float tmax, PtHitRec* best, uint32_t tlas_base = 0) {
    best->t = tmax; best->prim = PT_NONE; best->inst = PT_NONE;
    best->tidx = 0; best->iidx = 0;
#if !PT_SLIM_HIT
    best->U = 0.0f; best->V = 0.0f; best->ad = 1.0f; best->mat = 0;
#endif
    pt_ray_ctx(&s->world, o, d);
    s->r = s->world;
    s->tmin = tmin;
    s->sp = 0;
    s->in_blas = !TWO_LEVEL;
    s->node_base = TWO_LEVEL ? tlas_base : 0u; s->tri_base = 0; s->cur_inst = TWO_LEVEL ? PT_NONE : 0u; s->cur_iidx = 0;
    // root: one pending child of a virtual parent with child_base 0 and an empty imask, so popc(...) = 0 -> node 0
    s->ng.x = 0; s->ng.y = 0x80000000u;
    s->tg.x = 0; s->tg.y = 0;
}

// ANY = true: occlusion query, finishes as soon as any triangle is hit in (tmin, tmax).
// One step = at most one triangle test (or instance entry) followed, if the lane then has no triangle left, by at most one
// node visit.  A lane whose node produced a single triangle therefore does both in one iteration, lanes with more pending
// triangles spend extra iterations in the (short) triangle block only; the triangles of a node are always tested before
// any of its children is visited, so the visit order — and the node / triangle counters — equal the oracle's.
// The group stack is any type with put(i, e) / get(i): a plain array here, optionally shared memory for the first entries in the kernels.
struct PtArrayStack {
    PtU2 a[PT_STACK_SIZE];
    PT_HDM void put(int i, const PtU2& e) { a[i] = e; }
    PT_HDM PtU2 get(int i) const { return a[i]; }
};
template <bool ANY, bool TWO_LEVEL, class Counter, class Stack>
PT_HD int pt_trav_step(const PtSceneView& sc, PtTravState* s, Stack& stack, PtHitRec* best, Counter& cnt) {
    const bool do_tri = s->tg.y != 0;
    const bool leaf_tri = !TWO_LEVEL || s->in_blas;
    // the node visit of this step happens iff the lane has no triangle left after (at most) one test; whether it does is known
    // up front, so the triangle's and the node's 128-bit loads are issued TOGETHER and their latencies overlap
    const bool do_node = (s->ng.y & 0xff000000u) && ((s->tg.y & (s->tg.y - 1u)) == 0u) && (leaf_tri || !do_tri);
    uint32_t k = 0, tri_index = 0;
    if (do_tri) {
        k = (uint32_t)pt_ffs0(s->tg.y);
        s->tg.y &= s->tg.y - 1u;
        if (leaf_tri) tri_index = s->tri_base + s->tg.x + k;
    }
    // UNCONDITIONAL loads (record 0 when the lane has nothing to fetch: always valid and L1 resident): no branch around the loads, no
    // zero fill of their 32 destination registers (15 CS2R per step before) and, with the register pressure that removes, no spill at
    // the 64-register cap.  Measured on one box: 4.9 -> 6.5 Grays/s for the closest-hit kernel.
    const PtU4* tp = sc.tris + 3 * (size_t)tri_index;
    const PtU4 ta = pt_load4(tp), tb = pt_load4(tp + 1), tc = pt_load4(tp + 2);
    uint32_t child = 0, nbase = 0;
    if (do_node) {
        uint32_t bit = 31u - (uint32_t)pt_clz32(s->ng.y);
        s->ng.y &= ~(1u << bit);
        uint32_t slot = (bit - 24u) ^ s->r.oct_inv;
        child = s->ng.x + (uint32_t)pt_popc(s->ng.y & 0xffu & ~(0xffffffffu << slot));
        nbase = s->node_base;
        if (s->ng.y & 0xff000000u) {
            if (s->sp >= PT_STACK_SIZE) { s->sp = -1; return PT_STEP_DONE; }
            stack.put(s->sp++, s->ng);
        }
    }
    const PtU4* np = sc.nodes + 5 * (size_t)(nbase + child);
#if !PT_NODE_LOADS_LATE
    const PtU4 n0 = pt_load4(np), n1 = pt_load4(np + 1), n2 = pt_load4(np + 2), n3 = pt_load4(np + 3), n4 = pt_load4(np + 4);
#endif
    if (do_tri) {
        if (leaf_tri) {
            pt_test_tri_words(ta, tb, tc, tri_index, s->r, s->tmin, s->cur_inst, s->cur_iidx, best, cnt, sc.zero);
            if (ANY && best->prim != PT_NONE) return PT_STEP_DONE;
        } else {
            // TLAS leaf: enter the instance.  Save the remaining groups, push the return sentinel.
            if (s->sp + 3 > PT_STACK_SIZE) { s->sp = -1; return PT_STEP_DONE; }
            if (s->ng.y & 0xff000000u) stack.put(s->sp++, s->ng);   // popped last: the node's remaining instances come before its internal children
            if (s->tg.y) stack.put(s->sp++, s->tg);
            PtU2 sentinel; sentinel.x = PT_NONE; sentinel.y = 0; stack.put(s->sp++, sentinel);
            const PtU4* ip = sc.instances + 7 * (size_t)(s->tg.x + k);
            PtU4 m0 = pt_load4(ip), m1 = pt_load4(ip + 1), m2 = pt_load4(ip + 2), m6 = pt_load4(ip + 6);
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
#if PT_NODE_LOADS_LATE
    // variant: the node's five words are requested only after the triangle test (20 registers less across the test, but the two round
    // trips of a lane that does both in one step no longer overlap): measured -1 % at 8 CTAs / SM (profiles/r02_ab_traversal_build.log, run r2d)
#if defined(__CUDA_ARCH__)
    asm volatile("" ::: "memory");
#endif
    const PtU4 n0 = pt_load4(np), n1 = pt_load4(np + 1), n2 = pt_load4(np + 2), n3 = pt_load4(np + 3), n4 = pt_load4(np + 4);
#endif
    if (do_node) {   // children are culled against the best hit INCLUDING the triangle tested just above
        cnt.node();
        uint32_t hits = pt_node_hits(n0, n1, n2, n3, n4, s->r, s->tmin, best->t, sc.qbias);
        s->ng.x = n1.x; s->ng.y = (hits & 0xff000000u) | (n0.w >> 24);
        s->tg.x = n1.y; s->tg.y = hits & 0x00ffffffu;
    }
    // both groups empty: pop the next group (a node group keeps its hit bits in the top byte, a triangle group has none)
    while (!s->tg.y && !(s->ng.y & 0xff000000u)) {
        if (s->sp == 0) return PT_STEP_DONE;
        PtU2 e = stack.get(--s->sp);
        if (TWO_LEVEL && e.x == PT_NONE && e.y == 0) {   // leaving an instance
            s->r = s->world; s->in_blas = false; s->node_base = sc.tlas_base; s->tri_base = 0; s->cur_inst = PT_NONE;
            continue;
        }
        if (e.y & 0xff000000u) s->ng = e; else { s->tg = e; s->ng.x = 0; s->ng.y = 0; }
    }
    return PT_STEP_RUNNING;
}

// Whole traversal of one ray.  Returns false on traversal-stack overflow (reported through the status word).
template <bool ANY, bool TWO_LEVEL, class Counter>
PT_HD bool pt_traverse(const PtSceneView& sc, pt_v3 o, pt_v3 d, float tmin, float tmax, PtHitRec* best, Counter& cnt) {
    PtTravState s;
    PtArrayStack stack;
    pt_trav_init<TWO_LEVEL>(&s, o, d, tmin, tmax, best, sc.tlas_base);
    while (pt_trav_step<ANY, TWO_LEVEL>(sc, &s, stack, best, cnt) == PT_STEP_RUNNING) {}
    return s.sp >= 0;
}

// Barycentrics (weights of vertex 1 and 2) and material id of a finished closest hit.  With PT_SLIM_HIT they are re-derived from the hit triangle
// and — in a two-level scene — the instance the hit lies in: the world ray goes through the same world -> object transform as during the
// traversal and through the same pt_ray_tri, so the values are bit for bit the ones the loop would have carried.  wo / wd: the world-space ray.
template <bool TWO_LEVEL>
PT_HD void pt_hit_bary(const PtSceneView& sc, const PtHitRec& h, pt_v3 wo, pt_v3 wd, float* u, float* v, uint32_t* mat) {
#if PT_SLIM_HIT
    pt_v3 o = wo, d = wd;
    if (TWO_LEVEL) {
        const PtU4* ip = sc.instances + 7 * (size_t)h.iidx;
        const PtU4 m0 = pt_load4(ip), m1 = pt_load4(ip + 1), m2 = pt_load4(ip + 2);
        const float w2o[12] = {pt_u2f(m0.x), pt_u2f(m0.y), pt_u2f(m0.z), pt_u2f(m0.w), pt_u2f(m1.x), pt_u2f(m1.y),
                               pt_u2f(m1.z), pt_u2f(m1.w), pt_u2f(m2.x), pt_u2f(m2.y), pt_u2f(m2.z), pt_u2f(m2.w)};
        o = pt_xform_point(w2o, wo); d = pt_xform_vec(w2o, wd);
    }
    const PtU4* tp = sc.tris + 3 * (size_t)h.tidx;
    const PtU4 a = pt_load4(tp), b = pt_load4(tp + 1), c = pt_load4(tp + 2);
    float t, U = 0.0f, V = 0.0f, ad = 1.0f;
    pt_ray_tri(o, d, pt_mk(pt_u2f(a.x), pt_u2f(a.y), pt_u2f(a.z)), pt_mk(pt_u2f(b.x), pt_u2f(b.y), pt_u2f(b.z)), pt_mk(pt_u2f(c.x), pt_u2f(c.y), pt_u2f(c.z)), &t, &U, &V, &ad);
    *u = pt_div(U, ad); *v = pt_div(V, ad); *mat = b.w;
#else
    (void)sc; (void)wo; (void)wd;
    *u = pt_div(h.U, h.ad); *v = pt_div(h.V, h.ad); *mat = h.mat;
#endif
}

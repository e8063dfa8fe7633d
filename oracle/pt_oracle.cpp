// pt_oracle.cpp — CPU oracle for the path-tracing hot path.  TEST INFRASTRUCTURE ONLY.
//
// ** PARITY UNPINNED ** : mos9527/Foundation ships no ray tracer (its renderer is one textured quad,
// src/Renderer/Renderer.cpp:350 `DrawIndexed(6)`; only shader src/Renderer/Triangle.slang:23-37; no RT/AS
// extension, src/Platform/RHI/Vulkan/Device.cpp:13-15) and no fixture, golden vector or test for this path
// (SURVEY.md §0, §4, §8c).  There is nothing in the reference to check this oracle against, so it restates
// the north_star specification instead and validates ITSELF: exhaustive O(N) closest hit (no BVH) is the
// ground truth for hit IDs; the BVH path must agree with it exactly; the shading model is checked by
// furnace / reciprocity / estimator-agreement tests (tests/test_oracle.py).
// What pins this oracle from OUTSIDE the shared headers (nothing in the reference can): the PCG32 demo vector; a numpy float64 exhaustive
// closest hit written from scratch — the ground truth for MISSED hits, with the epsilons written out and the measured non-watertightness of
// the Moller-Trumbore form (tests/test_watertight.py); the Sobol (0,2) prefixes against scipy's Joe-Kuo sequence; the radiance under a square
// emitter against the closed-form configuration factor; the texture sampler against a float64 restatement (tests/test_textures.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library.  The product (foundation_b200/csrc) never includes, links or calls anything in oracle/.
//
// What is shared with the product, and why: the headers pt_math.h / pt_layout.h / pt_shading.h /
// pt_host_shared.h under foundation_b200/csrc hold the ARITHMETIC CONTRACT (ray-triangle test, Morton
// quantisation, PCG32, the BSDF, the outward quantisation rule) as IEEE-exact __host__ __device__
// functions, so a hit ID or a radiance value is the same bits on both machines.  Everything structural is
// written independently here, in the simplest sequential form:
//   - LBVH by top-down radix splits            (product: Karras 2012 parallel emit)
//   - std::stable_sort on (key, index)         (product: hand-written LSD radix sort)
//   - recursive BVH8 traversal                 (product: stack of node/triangle groups)
//   - one scalar loop per path                 (product: wavefront queues, compaction, material sort)
//   - sequential BFS collapse, plan by post-order (product: level-synchronous kernels + scans, plan inside the refit)
// Conventions restated from the reference: column-major view/proj (Renderer.cpp:28-33), Z-up RH camera and
// Vulkan Y flip (Renderer.cpp:373-380), top-left pixel origin, 1920x1080 RGBA8 target (Renderer.cpp:40-41).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../foundation_b200/csrc/pt_host_shared.h"
#include "../foundation_b200/csrc/pt_shading.h"

namespace {

struct Box3 {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
    void grow(const float* p) { for (int k = 0; k < 3; ++k) { lo[k] = pt_min(lo[k], p[k]); hi[k] = pt_max(hi[k], p[k]); } }
    void grow(const Box3& b) { for (int k = 0; k < 3; ++k) { lo[k] = pt_min(lo[k], b.lo[k]); hi[k] = pt_max(hi[k], b.hi[k]); } }
};

// ---------------------------------------------------------------------------------------------------
// LBVH over n primitives given boxes + centroids.  Children: < n-1 internal, else leaf (sorted position).
// Internal node numbering equals Karras': the node covering sorted range [a,b] split at g has children
// "g" (range [a,g]) and "g+1" (range [g+1,b]); the root is 0.
// ---------------------------------------------------------------------------------------------------
struct Bvh2 {
    uint32_t n = 0;
    std::vector<uint32_t> order;        // sorted position -> primitive
    std::vector<uint64_t> keys;         // sorted keys
    std::vector<uint32_t> left, right;  // n-1
    std::vector<uint32_t> first, last;  // n-1: sorted range of each internal node
    std::vector<Box3> box;              // 2n-1: internal nodes then leaves
    uint32_t count(uint32_t ref) const { return ref < n - 1 ? last[ref] - first[ref] + 1 : 1; }
    uint32_t lo_pos(uint32_t ref) const { return ref < n - 1 ? first[ref] : ref - (n - 1); }
};

void build_lbvh(const std::vector<Box3>& pbox, const std::vector<pt_v3>& cent, const Box3& bounds, Bvh2* out) {
    uint32_t n = (uint32_t)pbox.size();
    out->n = n;
    pt_v3 lo = pt_mk(bounds.lo[0], bounds.lo[1], bounds.lo[2]);
    float ext[3] = {bounds.hi[0] - bounds.lo[0], bounds.hi[1] - bounds.lo[1], bounds.hi[2] - bounds.lo[2]};
    pt_v3 inv = pt_mk(ext[0] > 0 ? pt_div(2097152.0f, ext[0]) : 0.0f, ext[1] > 0 ? pt_div(2097152.0f, ext[1]) : 0.0f,
                      ext[2] > 0 ? pt_div(2097152.0f, ext[2]) : 0.0f);
    std::vector<std::pair<uint64_t, uint32_t>> kv(n);
    for (uint32_t i = 0; i < n; ++i) kv[i] = {pt_morton63(cent[i], lo, inv), i};
    std::stable_sort(kv.begin(), kv.end());  // lexicographic (key, index)
    out->order.resize(n); out->keys.resize(n);
    for (uint32_t i = 0; i < n; ++i) { out->keys[i] = kv[i].first; out->order[i] = kv[i].second; }
    out->box.resize(2 * (size_t)n - 1);
    for (uint32_t j = 0; j < n; ++j) out->box[n - 1 + j] = pbox[out->order[j]];
    if (n < 2) return;
    out->left.resize(n - 1); out->right.resize(n - 1); out->first.resize(n - 1); out->last.resize(n - 1);
    const std::vector<uint64_t>& K = out->keys;
    struct Item { uint32_t a, b, idx; };
    std::vector<Item> st; st.push_back({0, n - 1, 0});
    std::vector<uint32_t> post;  // internal nodes in creation order (parents before children)
    while (!st.empty()) {
        Item it = st.back(); st.pop_back();
        uint32_t a = it.a, b = it.b;
        int dn = pt_delta(K[a], K[b], a, b);
        // largest g in [a, b-1] with delta(a, g) > dn
        uint32_t lo_g = a, hi_g = b - 1;
        while (lo_g < hi_g) {
            uint32_t mid = lo_g + (hi_g - lo_g + 1) / 2;
            if (pt_delta(K[a], K[mid], a, mid) > dn) lo_g = mid; else hi_g = mid - 1;
        }
        uint32_t g = lo_g;
        out->first[it.idx] = a; out->last[it.idx] = b;
        out->left[it.idx] = (a == g) ? (n - 1 + g) : g;
        out->right[it.idx] = (g + 1 == b) ? (n - 1 + g + 1) : (g + 1);
        post.push_back(it.idx);
        if (a != g) st.push_back({a, g, g});
        if (g + 1 != b) st.push_back({g + 1, b, g + 1});
    }
    for (size_t k = post.size(); k-- > 0;) {  // children before parents
        uint32_t i = post[k];
        Box3 bx = out->box[out->left[i]];
        bx.grow(out->box[out->right[i]]);
        out->box[i] = bx;
    }
}

// ---------------------------------------------------------------------------------------------------
// Collapse plan: the surface-area-cost dynamic programme of Ylitie, Karras, Laine (HPG 2017, section 3.1).
// cost[ref][i-1], i = 1..7 = cheapest way to represent the subtree of BVH2 node `ref` as at most i children
// of a wide node; plan[ref][j-2], j = 2..8 = how many of j slots go to the left child (low nibble), bit 7 =
// "j-1 slots are as cheap"; plan[ref][7] is unused.
// ---------------------------------------------------------------------------------------------------
struct CollapsePlan { std::vector<float> cost; std::vector<uint8_t> plan; };

static inline float ref_area(const Bvh2& b, uint32_t ref) {
    const Box3& x = b.box[ref];
    return pt_box_area(x.lo[0], x.lo[1], x.lo[2], x.hi[0], x.hi[1], x.hi[2]);
}
static inline float plan_cost(const Bvh2& b, const CollapsePlan& pl, uint32_t ref, int i) {   // i = 1..7
    return ref >= b.n - 1 ? ref_area(b, ref) * PT_COST_TRI : pl.cost[(size_t)ref * 7 + (i - 1)];
}
void plan_collapse(const Bvh2& b, uint32_t max_leaf, CollapsePlan* pl) {
    uint32_t n = b.n;
    if (n < 2) return;
    pl->cost.assign((size_t)(n - 1) * 7, 0.0f); pl->plan.assign((size_t)(n - 1) * 8, 0);
    std::vector<uint32_t> order; order.reserve(n - 1);
    std::vector<uint32_t> st{0};
    while (!st.empty()) {   // parents before children; processed in reverse
        uint32_t r = st.back(); st.pop_back(); order.push_back(r);
        if (b.left[r] < n - 1) st.push_back(b.left[r]);
        if (b.right[r] < n - 1) st.push_back(b.right[r]);
    }
    for (size_t q = order.size(); q-- > 0;) {
        uint32_t r = order[q], L = b.left[r], R = b.right[r];
        float cl[8], cr[8], D[9];
        for (int i = 1; i <= 7; ++i) { cl[i] = plan_cost(b, *pl, L, i); cr[i] = plan_cost(b, *pl, R, i); }
        uint8_t* P = &pl->plan[(size_t)r * 8];
        for (int j = 2; j <= 8; ++j) {
            int bk = 1; float bc = cl[1] + cr[j - 1];
            for (int k = 2; k < j; ++k) { float c = cl[k] + cr[j - k]; if (c < bc) { bc = c; bk = k; } }
            D[j] = bc; P[j - 2] = (uint8_t)bk;
        }
        float A = ref_area(b, r);
        float* C = &pl->cost[(size_t)r * 7];
        uint32_t cnt = b.count(r);
        C[0] = cnt <= max_leaf ? A * ((float)cnt * PT_COST_TRI) : A * PT_COST_NODE + D[8];   // a subtree that fits one leaf slot is always a leaf
        for (int i = 2; i <= 7; ++i) {
            if (C[i - 2] <= D[i]) { C[i - 1] = C[i - 2]; P[i - 2] |= 0x80; } else C[i - 1] = D[i];
        }
    }
}
// children of the wide node that stands for internal BVH2 node `ref`, in left-to-right order
static int plan_children(const Bvh2& b, const CollapsePlan& pl, uint32_t ref, uint32_t* C) {
    struct It { uint32_t r; int j; };
    It st[16]; int sp = 0, nc = 0;
    int k = pl.plan[(size_t)ref * 8 + 6] & 15;
    st[sp++] = {b.right[ref], 8 - k}; st[sp++] = {b.left[ref], k};
    while (sp) {
        It it = st[--sp];
        if (it.r >= b.n - 1 || it.j == 1) { C[nc++] = it.r; continue; }
        uint8_t p = pl.plan[(size_t)it.r * 8 + (it.j - 2)];
        if (p & 0x80) { st[sp++] = {it.r, it.j - 1}; continue; }
        int kk = p & 15;
        st[sp++] = {b.right[it.r], it.j - kk}; st[sp++] = {b.left[it.r], kk};
    }
    return nc;
}

// ---------------------------------------------------------------------------------------------------
// BVH2 -> BVH8 collapse (breadth first).  leaf_seq receives, in leaf order, the sorted positions of the
// primitives; node.tri_base + offset indexes into it.
// ---------------------------------------------------------------------------------------------------
void collapse8(const Bvh2& b, uint32_t max_leaf, float pad, std::vector<PtNode8>* nodes, std::vector<uint32_t>* leaf_seq) {
    nodes->clear(); leaf_seq->clear();
    uint32_t n = b.n;
    if (n == 0) return;
    uint32_t root_ref = (n == 1) ? 0u /* leaf 0 == ref n-1 == 0 */ : 0u;
    std::vector<uint32_t> level{root_ref}, next;
    uint32_t level_start = 0;
    CollapsePlan plan;
    plan_collapse(b, max_leaf, &plan);
    while (!level.empty()) {
        next.clear();
        uint32_t next_start = level_start + (uint32_t)level.size();
        nodes->resize(next_start);
        for (size_t w = 0; w < level.size(); ++w) {
            uint32_t ref = level[w];
            uint32_t C[8]; int nc = 0;
            bool ref_is_leaf = (n == 1) || b.count(ref) <= max_leaf;
            if (ref_is_leaf) C[nc++] = (n == 1) ? 0u : ref;
            else nc = plan_children(b, plan, ref, C);
            const Box3& nb = b.box[(n == 1) ? 0 : ref];
            // greedy octant slot assignment
            float cost[8][8];
            for (int k = 0; k < nc; ++k) {
                const Box3& cb = b.box[C[k]];
                float dx = (cb.lo[0] + cb.hi[0]) * 0.5f - (nb.lo[0] + nb.hi[0]) * 0.5f;
                float dy = (cb.lo[1] + cb.hi[1]) * 0.5f - (nb.lo[1] + nb.hi[1]) * 0.5f;
                float dz = (cb.lo[2] + cb.hi[2]) * 0.5f - (nb.lo[2] + nb.hi[2]) * 0.5f;
                for (int s = 0; s < 8; ++s) cost[k][s] = (((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy)) + ((s & 1) ? dz : -dz);
            }
            int slot_child[8]; bool child_done[8] = {false};
            for (int s = 0; s < 8; ++s) slot_child[s] = -1;
            for (int it = 0; it < nc; ++it) {
                int bk = -1, bs = -1; float bc = -INFINITY;
                for (int k = 0; k < nc; ++k) {
                    if (child_done[k]) continue;
                    for (int s = 0; s < 8; ++s) {
                        if (slot_child[s] >= 0) continue;
                        if (bk < 0 || cost[k][s] > bc) { bc = cost[k][s]; bk = k; bs = s; }
                    }
                }
                slot_child[bs] = bk; child_done[bk] = true;
            }
            PtNode8 nd; memset(&nd, 0, sizeof nd);
            float p[3], inv[3]; uint32_t e[3];
            for (int k = 0; k < 3; ++k) {
                p[k] = nb.lo[k] - pad;
                float ext = (nb.hi[k] + pad) - p[k];
                e[k] = pt_quant_exp(ext);
                inv[k] = pt_u2f((254u - e[k]) << 23);
            }
            nd.px = p[0]; nd.py = p[1]; nd.pz = p[2];
            nd.ex = (uint8_t)e[0]; nd.ey = (uint8_t)e[1]; nd.ez = (uint8_t)e[2];
            nd.child_base = next_start + (uint32_t)next.size();
            nd.tri_base = (uint32_t)leaf_seq->size();
            uint32_t tri_off = 0;
            for (int s = 0; s < 8; ++s) {
                int k = slot_child[s];
                if (k < 0) { nd.qlox[s] = nd.qloy[s] = nd.qloz[s] = 255; nd.qhix[s] = nd.qhiy[s] = nd.qhiz[s] = 0; continue; }
                const Box3& cb = b.box[C[k]];
                nd.qlox[s] = (uint8_t)pt_quant_lo(cb.lo[0] - pad, p[0], inv[0]); nd.qhix[s] = (uint8_t)pt_quant_hi(cb.hi[0] + pad, p[0], inv[0]);
                nd.qloy[s] = (uint8_t)pt_quant_lo(cb.lo[1] - pad, p[1], inv[1]); nd.qhiy[s] = (uint8_t)pt_quant_hi(cb.hi[1] + pad, p[1], inv[1]);
                nd.qloz[s] = (uint8_t)pt_quant_lo(cb.lo[2] - pad, p[2], inv[2]); nd.qhiz[s] = (uint8_t)pt_quant_hi(cb.hi[2] + pad, p[2], inv[2]);
                uint32_t cnt = (n == 1) ? 1 : b.count(C[k]);
                if (cnt <= max_leaf) {
                    uint32_t fp = (n == 1) ? 0 : b.lo_pos(C[k]);
                    nd.meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | tri_off);
                    for (uint32_t q = 0; q < cnt; ++q) leaf_seq->push_back(fp + q);
                    tri_off += cnt;
                } else {
                    nd.imask |= (uint8_t)(1u << s);
                    nd.meta[s] = (uint8_t)(0x20u | (24u + s));
                    next.push_back(C[k]);
                }
            }
            (*nodes)[level_start + w] = nd;
        }
        level_start = next_start;
        level.swap(next);
    }
}

struct Mesh {
    uint32_t ntris = 0;
    std::vector<float> v;        // 9 floats per triangle (resolved vertices)
    std::vector<uint32_t> mat;
    Box3 bounds; float pad = 0;
    std::vector<uint32_t> order;  // sorted position -> input triangle
    std::vector<float> uv, col;   // optional per-CORNER attributes (6 / 9 floats per triangle, resolved through the indices like `v`)
    std::vector<PtNode8> nodes;
    std::vector<PtTri> tris;      // leaf order
};

struct Inst { uint32_t mesh; float o2w[12]; };

struct Scene {
    std::vector<Mesh> meshes;
    std::vector<PtMaterial> mats;
    std::vector<Inst> insts; bool has_insts = false;
    uint32_t max_leaf = PT_MAX_LEAF;
    // TLAS
    std::vector<PtNode8> tnodes; std::vector<PtInstance> tinst; std::vector<uint32_t> torder; Box3 wbounds;
    std::vector<PtLight> lights; float light_area = 0; float ray_eps = 0;
    struct Tex { std::vector<uint32_t> texels; uint32_t w = 0, h = 0; };
    std::vector<Tex> textures; std::vector<uint32_t> mat_tex;
    std::vector<PtNode8> fnodes; std::vector<PtTri> ftris;  // device-like flat arrays: [TLAS | BLAS 0 | BLAS 1 ..]
    PtCamera cam; bool committed = false;
};

void build_mesh(Mesh& m, uint32_t max_leaf) {
    uint32_t n = m.ntris;
    std::vector<Box3> pb(n); std::vector<pt_v3> ce(n);
    m.bounds.reset();
    for (uint32_t i = 0; i < n; ++i) {
        const float* t = &m.v[9 * (size_t)i];
        pb[i].reset(); pb[i].grow(t); pb[i].grow(t + 3); pb[i].grow(t + 6);
        m.bounds.grow(pb[i]);
        ce[i] = pt_tri_centroid(pt_mk(t[0], t[1], t[2]), pt_mk(t[3], t[4], t[5]), pt_mk(t[6], t[7], t[8]));
    }
    m.pad = pt_pad_for(m.bounds.lo, m.bounds.hi);
    Bvh2 b;
    build_lbvh(pb, ce, m.bounds, &b);
    m.order = b.order;
    std::vector<uint32_t> seq;
    collapse8(b, max_leaf, m.pad, &m.nodes, &seq);
    m.tris.resize(n);
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t i = b.order[seq[k]];
        const float* t = &m.v[9 * (size_t)i];
        PtTri& o = m.tris[k];
        o.v0x = t[0]; o.v0y = t[1]; o.v0z = t[2]; o.prim = i;
        o.e1x = t[3] - t[0]; o.e1y = t[4] - t[1]; o.e1z = t[5] - t[2]; o.mat = m.mat[i];
        o.e2x = t[6] - t[0]; o.e2y = t[7] - t[1]; o.e2z = t[8] - t[2]; o.pad = 0;
    }
}

void commit(Scene& s) {
    for (auto& m : s.meshes) build_mesh(m, s.max_leaf);
    s.lights.clear();
    s.wbounds.reset();
    if (s.has_insts) {
        uint32_t ni = (uint32_t)s.insts.size();
        std::vector<Box3> pb(ni); std::vector<pt_v3> ce(ni);
        std::vector<PtInstance> rec(ni);
        for (uint32_t i = 0; i < ni; ++i) {
            const Mesh& m = s.meshes[s.insts[i].mesh];
            float lo[3], hi[3];
            for (int k = 0; k < 3; ++k) { lo[k] = m.bounds.lo[k] - m.pad; hi[k] = m.bounds.hi[k] + m.pad; }
            pt_world_box(s.insts[i].o2w, lo, hi, pb[i].lo, pb[i].hi);
            ce[i] = pt_mk((pb[i].lo[0] + pb[i].hi[0]) * 0.5f, (pb[i].lo[1] + pb[i].hi[1]) * 0.5f, (pb[i].lo[2] + pb[i].hi[2]) * 0.5f);
            s.wbounds.grow(pb[i]);
            memcpy(rec[i].o2w, s.insts[i].o2w, 48);
            pt_invert_affine(rec[i].o2w, rec[i].w2o);
            rec[i].mesh_id = s.insts[i].mesh; rec[i].inst_id = i;
        }
        // node/tri bases: BLAS b starts after the TLAS nodes and the BLAS before it
        Bvh2 b;
        build_lbvh(pb, ce, s.wbounds, &b);
        s.torder = b.order;
        std::vector<uint32_t> seq;
        collapse8(b, 1, pt_pad_for(s.wbounds.lo, s.wbounds.hi), &s.tnodes, &seq);
        std::vector<uint32_t> nbase(s.meshes.size()), tbase(s.meshes.size());
        uint32_t nb = (uint32_t)s.tnodes.size(), tb = 0;
        for (size_t m = 0; m < s.meshes.size(); ++m) { nbase[m] = nb; tbase[m] = tb; nb += (uint32_t)s.meshes[m].nodes.size(); tb += s.meshes[m].ntris; }
        s.tinst.resize(ni);
        for (uint32_t k = 0; k < ni; ++k) {
            s.tinst[k] = rec[b.order[seq[k]]];
            s.tinst[k].node_base = nbase[s.tinst[k].mesh_id]; s.tinst[k].tri_base = tbase[s.tinst[k].mesh_id];
        }
        // lights: instance order, then triangle input order
        for (uint32_t i = 0; i < ni; ++i) {
            const Mesh& m = s.meshes[s.insts[i].mesh];
            for (uint32_t t = 0; t < m.ntris; ++t) {
                const PtMaterial& mt = s.mats[m.mat[t] < s.mats.size() ? m.mat[t] : 0];
                if (!(mt.er > 0 || mt.eg > 0 || mt.eb > 0)) continue;
                const float* v = &m.v[9 * (size_t)t];
                pt_v3 v0 = pt_mk(v[0], v[1], v[2]), e1 = pt_mk(v[3] - v[0], v[4] - v[1], v[5] - v[2]), e2 = pt_mk(v[6] - v[0], v[7] - v[1], v[8] - v[2]);
                PtLight l;
                pt_light_make(&l, pt_xform_point(s.insts[i].o2w, v0), pt_xform_vec(s.insts[i].o2w, e1), pt_xform_vec(s.insts[i].o2w, e2), mt.er, mt.eg, mt.eb);
                s.lights.push_back(l);
            }
        }
    } else {
        for (auto& m : s.meshes) s.wbounds.grow(m.bounds);
        const Mesh& m = s.meshes[0];
        for (uint32_t t = 0; t < m.ntris; ++t) {
            const PtMaterial& mt = s.mats[m.mat[t] < s.mats.size() ? m.mat[t] : 0];
            if (!(mt.er > 0 || mt.eg > 0 || mt.eb > 0)) continue;
            const float* v = &m.v[9 * (size_t)t];
            PtLight l;
            pt_light_make(&l, pt_mk(v[0], v[1], v[2]), pt_mk(v[3] - v[0], v[4] - v[1], v[5] - v[2]), pt_mk(v[6] - v[0], v[7] - v[1], v[8] - v[2]), mt.er, mt.eg, mt.eb);
            s.lights.push_back(l);
        }
    }
    s.fnodes.clear(); s.ftris.clear();
    if (s.has_insts) {
        s.fnodes = s.tnodes;
        for (auto& m : s.meshes) { s.fnodes.insert(s.fnodes.end(), m.nodes.begin(), m.nodes.end()); s.ftris.insert(s.ftris.end(), m.tris.begin(), m.tris.end()); }
    } else { s.fnodes = s.meshes[0].nodes; s.ftris = s.meshes[0].tris; }
    uint32_t nl = (uint32_t)s.lights.size();
    s.light_area = pt_lights_finalize(s.lights.data(), &nl);   // drops degenerate emitters (same rule as the product)
    s.lights.resize(nl);
    float ext = 0;
    for (int k = 0; k < 3; ++k) ext = pt_max(ext, s.wbounds.hi[k] - s.wbounds.lo[k]);
    s.ray_eps = ext * PT_RAY_EPS_REL;
    s.committed = true;
}

// ---------------------------------------------------------------------------------------------------
// traversal (recursive restatement)
// ---------------------------------------------------------------------------------------------------
struct Hit { float t, U, V, ad; uint32_t prim, inst; };
struct Counters { uint64_t nodes = 0, tris = 0, insts = 0; };

struct RayC { pt_v3 o, d, idir; bool neg[3]; uint32_t oct_inv; };
float safe_rcp(float d) {
    float a = pt_abs(d) < 9.094947017729282e-13f ? pt_copysign(9.094947017729282e-13f, d) : d;
    return pt_div(1.0f, a);
}
RayC make_ray(pt_v3 o, pt_v3 d) {
    RayC r; r.o = o; r.d = d; r.idir = pt_mk(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    r.neg[0] = pt_f2u(d.x) >> 31; r.neg[1] = pt_f2u(d.y) >> 31; r.neg[2] = pt_f2u(d.z) >> 31;
    r.oct_inv = (r.neg[0] ? 0 : 4) | (r.neg[1] ? 0 : 2) | (r.neg[2] ? 0 : 1);
    return r;
}

inline void test_tri(const PtTri& tr, const RayC& r, float tmin, uint32_t inst, Hit* best, Counters* c) {
    c->tris++;
    float t, U, V, ad;
    if (!pt_ray_tri(r.o, r.d, pt_mk(tr.v0x, tr.v0y, tr.v0z), pt_mk(tr.e1x, tr.e1y, tr.e1z), pt_mk(tr.e2x, tr.e2y, tr.e2z), &t, &U, &V, &ad)) return;
    uint64_t id = ((uint64_t)inst << 32) | tr.prim, bid = ((uint64_t)best->inst << 32) | best->prim;
    if (pt_closer(t, id, tmin, best->t, bid, best->prim != PT_NONE)) { best->t = t; best->U = U; best->V = V; best->ad = ad; best->prim = tr.prim; best->inst = inst; }
}

struct Trav {
    const Scene* s; bool any; float tmin; Hit* best; Counters* c; bool done = false;
    RayC world;

    // child-box test on the quantised grid: the same roundings as the device (fma(QBIAS + q, scale*idir, (p-o)*idir - QBIAS*scale*idir))
    bool child_hit(const PtNode8& n, int slot, const RayC& r) const {
        const float p[3] = {n.px, n.py, n.pz};
        const uint8_t e[3] = {n.ex, n.ey, n.ez};
        const uint8_t* ql[3] = {n.qlox, n.qloy, n.qloz};
        const uint8_t* qh[3] = {n.qhix, n.qhiy, n.qhiz};
        const float o[3] = {r.o.x, r.o.y, r.o.z}, id[3] = {r.idir.x, r.idir.y, r.idir.z};
        float tn = tmin, tf = best->t;
        float tnk[3], tfk[3];
        for (int k = 0; k < 3; ++k) {
            float a = pt_u2f((uint32_t)e[k] << 23) * id[k];
            float bb = (p[k] - o[k]) * id[k];
            float err = pt_fma(pt_abs(a), PT_SLAB_QMAX, pt_abs(bb)) * PT_SLAB_EPS;   // ray-dependent slack, same rule as the device (pt_layout.h)
            float cc = pt_fma(-PT_QBIAS, a, bb);                                     // the byte enters as QBIAS + q, the bias is folded in here
            float qn = PT_QBIAS + (float)(r.neg[k] ? qh[k][slot] : ql[k][slot]), qf = PT_QBIAS + (float)(r.neg[k] ? ql[k][slot] : qh[k][slot]);
            tnk[k] = pt_fma(qn, a, cc - err); tfk[k] = pt_fma(qf, a, cc + err);
        }
        tn = fmaxf(fmaxf(tnk[0], tnk[1]), fmaxf(tnk[2], tmin));
        tf = fminf(fminf(tfk[0], tfk[1]), fminf(tfk[2], best->t));
        return tn <= tf;
    }

    void visit(const PtNode8* nodes, uint32_t node_base, uint32_t idx, const RayC& r, bool blas, uint32_t tri_base, uint32_t inst) {
        const PtNode8& n = nodes[node_base + idx];
        c->nodes++;
        bool hit[8];
        for (int sl = 0; sl < 8; ++sl) hit[sl] = n.meta[sl] != 0 && child_hit(n, sl, r);
        for (int sl = 0; sl < 8 && !done; ++sl) {       // leaves first, ascending triangle offset
            if (!hit[sl] || (n.imask >> sl & 1)) continue;
            uint32_t off = n.meta[sl] & 31u, bits = n.meta[sl] >> 5;
            uint32_t cnt = bits == 1 ? 1 : bits == 3 ? 2 : 3;
            for (uint32_t q = 0; q < cnt && !done; ++q) {
                uint32_t k = n.tri_base + off + q;
                if (blas) {
                    test_tri(s->ftris[tri_base + k], r, tmin, inst, best, c);
                    if (any && best->prim != PT_NONE) done = true;
                } else {
                    const PtInstance& in = s->tinst[k];
                    c->insts++;
                    RayC lr = make_ray(pt_xform_point(in.w2o, world.o), pt_xform_vec(in.w2o, world.d));
                    visit(s->fnodes.data(), in.node_base, 0, lr, true, in.tri_base, in.inst_id);
                }
            }
        }
        for (int pos = 7; pos >= 0 && !done; --pos) {    // internal children, front to back by octant
            int sl = pos ^ (int)r.oct_inv;
            if (!hit[sl] || !(n.imask >> sl & 1)) continue;
            uint32_t rel = (uint32_t)pt_popc(n.imask & ((1u << sl) - 1u));
            visit(nodes, node_base, n.child_base + rel, r, blas, tri_base, inst);
        }
    }
};


Hit trace_bvh(const Scene& s, pt_v3 o, pt_v3 d, float tmin, float tmax, bool any, Counters* c) {
    Hit best; best.t = tmax; best.U = best.V = 0; best.ad = 1; best.prim = PT_NONE; best.inst = PT_NONE;
    Trav tv; tv.s = &s; tv.any = any; tv.tmin = tmin; tv.best = &best; tv.c = c;
    tv.world = make_ray(o, d);
    if (s.fnodes.empty()) return best;
    if (s.has_insts) tv.visit(s.fnodes.data(), 0, 0, tv.world, false, 0, PT_NONE);
    else tv.visit(s.fnodes.data(), 0, 0, tv.world, true, 0, 0);
    return best;
}

// exhaustive ground truth: every triangle of every instance, input order, no acceleration structure
Hit trace_brute(const Scene& s, pt_v3 o, pt_v3 d, float tmin, float tmax, bool any, Counters* c) {
    Hit best; best.t = tmax; best.U = best.V = 0; best.ad = 1; best.prim = PT_NONE; best.inst = PT_NONE;
    auto run = [&](const Mesh& m, const RayC& r, uint32_t inst) {
        for (uint32_t i = 0; i < m.ntris; ++i) {
            const float* t = &m.v[9 * (size_t)i];
            PtTri tr; tr.v0x = t[0]; tr.v0y = t[1]; tr.v0z = t[2]; tr.prim = i;
            tr.e1x = t[3] - t[0]; tr.e1y = t[4] - t[1]; tr.e1z = t[5] - t[2];
            tr.e2x = t[6] - t[0]; tr.e2y = t[7] - t[1]; tr.e2z = t[8] - t[2];
            test_tri(tr, r, tmin, inst, &best, c);
            if (any && best.prim != PT_NONE) return;
        }
    };
    if (s.has_insts) {
        for (uint32_t i = 0; i < s.insts.size(); ++i) {
            float w2o[12]; pt_invert_affine(s.insts[i].o2w, w2o);
            run(s.meshes[s.insts[i].mesh], make_ray(pt_xform_point(w2o, o), pt_xform_vec(w2o, d)), i);
            if (any && best.prim != PT_NONE) break;
        }
    } else run(s.meshes[0], make_ray(o, d), 0);
    return best;
}

template <class F>
void parallel_for(uint64_t n, int nthreads, uint64_t chunk, F f) {
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next{0};
    auto worker = [&](int tid) {
        for (;;) {
            uint64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            f(b, std::min(n, b + chunk), tid);
        }
    };
    if (nthreads == 1) { worker(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
}

struct HitOut { float t, u, v; uint32_t prim; };
struct RayIn { float o[3], tmin, d[3], tmax; };

// world-space edges + material of a hit
void hit_surface(const Scene& s, const Hit& h, pt_v3* e1, pt_v3* e2, PtMaterial* mat) {
    const Mesh& m = s.has_insts ? s.meshes[s.insts[h.inst].mesh] : s.meshes[0];
    const float* t = &m.v[9 * (size_t)h.prim];
    pt_v3 a = pt_mk(t[3] - t[0], t[4] - t[1], t[5] - t[2]), b = pt_mk(t[6] - t[0], t[7] - t[1], t[8] - t[2]);
    if (s.has_insts) { a = pt_xform_vec(s.insts[h.inst].o2w, a); b = pt_xform_vec(s.insts[h.inst].o2w, b); }
    *e1 = a; *e2 = b;
    uint32_t mi = m.mat[h.prim];
    mi = mi < s.mats.size() ? mi : 0;
    *mat = s.mats[mi];
    if (!m.uv.empty() || !m.col.empty()) {   // base colour x vertex colour x albedo texel: the shared arithmetic, fed with per-corner (unindexed) streams
        PtMeshAttr a; a.uv = m.uv.empty() ? nullptr : (const uint8_t*)m.uv.data(); a.col = m.col.empty() ? nullptr : (const uint8_t*)m.col.data();
        a.idx = nullptr; a.uv_stride = 8; a.col_stride = 12; a.idx_fmt = 0; a.pad = 0;
        std::vector<PtTexture> td(s.textures.size() ? s.textures.size() : 1);
        for (size_t k = 0; k < s.textures.size(); ++k) { td[k].texels = s.textures[k].texels.data(); td[k].width = s.textures[k].w; td[k].height = s.textures[k].h; td[k].pad = 0; }
        uint32_t tid = mi < s.mat_tex.size() && s.mat_tex[mi] < s.textures.size() ? s.mat_tex[mi] : PT_NONE;
        pt_material_apply_attributes(mat, a, td.data(), tid, h.prim, pt_div(h.U, h.ad), pt_div(h.V, h.ad));
    }
}

bool owns_pixel(uint32_t x, uint32_t y, uint32_t rank, uint32_t count, uint32_t tile) {
    if (count <= 1) return true;
    return ((x / tile) + (y / tile)) % count == rank;
}

}  // namespace

// =====================================================================================================
// C interface (ctypes).  Mirrors the product's C ABI in shape so tests read the same on both sides.
// =====================================================================================================
extern "C" {

void* orc_create(uint32_t max_leaf) { Scene* s = new Scene(); s->max_leaf = max_leaf ? max_leaf : PT_MAX_LEAF; s->mats.push_back(PtMaterial{0.8f, 0.8f, 0.8f, 0.5f, 0, 0, 0, 0}); return s; }
void orc_destroy(void* p) { delete (Scene*)p; }

void orc_materials_set(void* p, const float* m, uint32_t n) {
    Scene* s = (Scene*)p; s->mats.resize(n);
    for (uint32_t i = 0; i < n; ++i) memcpy(&s->mats[i], m + 8 * (size_t)i, 32);
}
// positions: tightly packed float3; indices: uint32 triples or NULL
uint32_t orc_mesh_add(void* p, const float* pos, uint32_t nverts, const uint32_t* idx, uint32_t ntris, const uint32_t* mat) {
    Scene* s = (Scene*)p; s->meshes.emplace_back(); Mesh& m = s->meshes.back();
    (void)nverts;
    m.ntris = ntris; m.v.resize(9 * (size_t)ntris); m.mat.resize(ntris);
    for (uint32_t i = 0; i < ntris; ++i) {
        for (int k = 0; k < 3; ++k) {
            uint32_t vi = idx ? idx[3 * (size_t)i + k] : 3 * i + k;
            memcpy(&m.v[9 * (size_t)i + 3 * k], pos + 3 * (size_t)vi, 12);
        }
        m.mat[i] = mat ? mat[i] : 0;
    }
    return (uint32_t)s->meshes.size() - 1;
}
// per-vertex uv (float2, tight) / colour (float3, tight), either may be NULL; idx as in orc_mesh_add
void orc_mesh_attributes_set(void* p, uint32_t mesh, const float* uv, const float* col, const uint32_t* idx) {
    Mesh& m = ((Scene*)p)->meshes[mesh];
    m.uv.clear(); m.col.clear();
    if (uv) m.uv.resize(6 * (size_t)m.ntris);
    if (col) m.col.resize(9 * (size_t)m.ntris);
    for (uint32_t i = 0; i < m.ntris; ++i)
        for (int k = 0; k < 3; ++k) {
            uint32_t vi = idx ? idx[3 * (size_t)i + k] : 3 * i + k;
            if (uv) memcpy(&m.uv[6 * (size_t)i + 2 * k], uv + 2 * (size_t)vi, 8);
            if (col) memcpy(&m.col[9 * (size_t)i + 3 * k], col + 3 * (size_t)vi, 12);
        }
}
uint32_t orc_texture_add(void* p, const uint32_t* rgba8, uint32_t w, uint32_t h) {
    Scene* s = (Scene*)p; s->textures.emplace_back();
    s->textures.back().texels.assign(rgba8, rgba8 + (size_t)w * h); s->textures.back().w = w; s->textures.back().h = h;
    return (uint32_t)s->textures.size() - 1;
}
void orc_material_textures_set(void* p, const uint32_t* ids, uint32_t n) { ((Scene*)p)->mat_tex.assign(ids, ids + n); }
void orc_texture_sample(const uint32_t* rgba8, uint32_t w, uint32_t h, float u, float v, float* rgb) {
    PtTexture t; t.texels = rgba8; t.width = w; t.height = h; t.pad = 0;
    pt_v3 c = pt_texture_sample(t, u, v); rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
void orc_instances_set(void* p, const uint32_t* mesh_ids, const float* xf, uint32_t n) {
    Scene* s = (Scene*)p; s->insts.resize(n); s->has_insts = true;
    for (uint32_t i = 0; i < n; ++i) { s->insts[i].mesh = mesh_ids[i]; memcpy(s->insts[i].o2w, xf + 12 * (size_t)i, 48); }
}
void orc_commit(void* p) { commit(*(Scene*)p); }

void orc_blas_counts(void* p, uint32_t mesh, uint64_t* nn, uint64_t* nt) { Scene* s = (Scene*)p; *nn = s->meshes[mesh].nodes.size(); *nt = s->meshes[mesh].tris.size(); }
void orc_blas_get(void* p, uint32_t mesh, void* nodes, void* tris, uint32_t* order) {
    Mesh& m = ((Scene*)p)->meshes[mesh];
    if (nodes) memcpy(nodes, m.nodes.data(), m.nodes.size() * 80);
    if (tris) memcpy(tris, m.tris.data(), m.tris.size() * 48);
    if (order) memcpy(order, m.order.data(), m.order.size() * 4);
}
void orc_tlas_counts(void* p, uint64_t* nn, uint64_t* ni) { Scene* s = (Scene*)p; *nn = s->tnodes.size(); *ni = s->tinst.size(); }
void orc_tlas_get(void* p, void* nodes, uint32_t* order, void* inst_records) {
    Scene* s = (Scene*)p;
    if (nodes) memcpy(nodes, s->tnodes.data(), s->tnodes.size() * 80);
    if (order) memcpy(order, s->torder.data(), s->torder.size() * 4);
    if (inst_records) memcpy(inst_records, s->tinst.data(), s->tinst.size() * sizeof(PtInstance));
}
void orc_scene_info(void* p, float* lo, float* hi, float* ray_eps, uint32_t* nlights, float* light_area) {
    Scene* s = (Scene*)p;
    for (int k = 0; k < 3; ++k) { lo[k] = s->wbounds.lo[k]; hi[k] = s->wbounds.hi[k]; }
    *ray_eps = s->ray_eps; *nlights = (uint32_t)s->lights.size(); *light_area = s->light_area;
}

// mode 0: BVH traversal, 1: brute force.  counters (may be NULL): [nodes visited, triangles tested, instances entered]
void orc_trace_closest(void* p, const void* rays_, uint64_t n, void* hits_, uint32_t* inst_out, int mode, int nthreads, uint64_t* counters) {
    const Scene& s = *(Scene*)p; const RayIn* rays = (const RayIn*)rays_; HitOut* hits = (HitOut*)hits_;
    std::atomic<uint64_t> cn{0}, ct{0}, ci{0};
    parallel_for(n, nthreads, mode ? 4 : 4096, [&](uint64_t b, uint64_t e, int) {
        Counters c;
        for (uint64_t i = b; i < e; ++i) {
            const RayIn& r = rays[i];
            pt_v3 o = pt_mk(r.o[0], r.o[1], r.o[2]), d = pt_mk(r.d[0], r.d[1], r.d[2]);
            Hit h = mode ? trace_brute(s, o, d, r.tmin, r.tmax, false, &c) : trace_bvh(s, o, d, r.tmin, r.tmax, false, &c);
            if (h.prim == PT_NONE) { hits[i].t = INFINITY; hits[i].u = hits[i].v = 0; hits[i].prim = PT_NONE; if (inst_out) inst_out[i] = PT_NONE; }
            else { hits[i].t = h.t; hits[i].u = pt_div(h.U, h.ad); hits[i].v = pt_div(h.V, h.ad); hits[i].prim = h.prim; if (inst_out) inst_out[i] = h.inst; }
        }
        cn += c.nodes; ct += c.tris; ci += c.insts;
    });
    if (counters) { counters[0] = cn; counters[1] = ct; counters[2] = ci; }
}
void orc_trace_any(void* p, const void* rays_, uint64_t n, uint8_t* occ, int mode, int nthreads, uint64_t* counters) {
    const Scene& s = *(Scene*)p; const RayIn* rays = (const RayIn*)rays_;
    std::atomic<uint64_t> cn{0}, ct{0}, ci{0};
    parallel_for(n, nthreads, mode ? 4 : 4096, [&](uint64_t b, uint64_t e, int) {
        Counters c;
        for (uint64_t i = b; i < e; ++i) {
            const RayIn& r = rays[i];
            pt_v3 o = pt_mk(r.o[0], r.o[1], r.o[2]), d = pt_mk(r.d[0], r.d[1], r.d[2]);
            Hit h = mode ? trace_brute(s, o, d, r.tmin, r.tmax, true, &c) : trace_bvh(s, o, d, r.tmin, r.tmax, true, &c);
            occ[i] = h.prim != PT_NONE;
        }
        cn += c.nodes; ct += c.tris; ci += c.insts;
    });
    if (counters) { counters[0] = cn; counters[1] = ct; counters[2] = ci; }
}

int orc_camera_set(void* p, const float* view, const float* proj) { return pt_camera_derive(view, proj, &((Scene*)p)->cam) ? 0 : -1; }
void orc_camera_get(void* p, float* out12) { memcpy(out12, &((Scene*)p)->cam, 48); }

// Adds samples [s0, s0+ns) of every owned pixel to accum (float4 per pixel: rgb sum, w = sample count).
// ray_counts (may be NULL): [extension rays, shadow rays].  brute != 0: exhaustive intersection (tiny scenes).
void orc_render(void* p, uint32_t width, uint32_t height, uint64_t seed, uint32_t s0, uint32_t ns, uint32_t max_bounces, uint32_t flags,
                const float* bg, uint32_t rank, uint32_t count, uint32_t tile, float* accum, int nthreads, int brute, uint64_t* ray_counts) {
    const Scene& s = *(Scene*)p;
    if (!tile) tile = 32;
    PtShadeConsts sc; sc.lights = s.lights.data(); sc.num_lights = (uint32_t)s.lights.size(); sc.light_area = s.light_area;
    sc.ray_eps = s.ray_eps; sc.flags = flags; sc.seed = seed; sc.max_bounces = max_bounces; sc.bg[0] = bg[0]; sc.bg[1] = bg[1]; sc.bg[2] = bg[2];
    std::atomic<uint64_t> n_ext{0}, n_sh{0};
    parallel_for((uint64_t)width * height, nthreads, 256, [&](uint64_t b, uint64_t e, int) {
        Counters c; uint64_t ext = 0, shd = 0;
        for (uint64_t pix = b; pix < e; ++pix) {
            uint32_t x = (uint32_t)(pix % width), y = (uint32_t)(pix / width);
            if (!owns_pixel(x, y, rank, count, tile)) continue;
            for (uint32_t smp = s0; smp < s0 + ns; ++smp) {
                PtPath path;
                pt_path_init(&path, s.cam, seed, (uint32_t)pix, smp, width, height, flags);
                for (;;) {
                    ++ext;
                    Hit h = brute ? trace_brute(s, path.o, path.d, 0.0f, INFINITY, false, &c) : trace_bvh(s, path.o, path.d, 0.0f, INFINITY, false, &c);
                    if (h.prim == PT_NONE) { pt_shade_miss(&path, sc); break; }
                    pt_v3 e1, e2; PtMaterial mat;
                    hit_surface(s, h, &e1, &e2, &mat);
                    PtShadowRay sh;
                    bool alive = pt_shade_vertex(&path, sc, h.t, e1, e2, mat, &sh);
                    if (sh.valid) {
                        ++shd;
                        Hit o = brute ? trace_brute(s, sh.o, sh.d, 0.0f, sh.tmax, true, &c) : trace_bvh(s, sh.o, sh.d, 0.0f, sh.tmax, true, &c);
                        if (o.prim == PT_NONE) path.L = pt_add(path.L, sh.contrib);
                    }
                    if (!alive) break;
                }
                float* a = accum + 4 * pix;
                a[0] += path.L.x; a[1] += path.L.y; a[2] += path.L.z; a[3] += 1.0f;
            }
        }
        n_ext += ext; n_sh += shd;
    });
    if (ray_counts) { ray_counts[0] = n_ext; ray_counts[1] = n_sh; }
}

// direct access to the shading arithmetic for unit tests (energy conservation, reciprocity, sampling pdf)
void orc_bsdf_eval(const float* mat8, const float* wo, const float* wi, float* f3, float* pdf) {
    PtMaterial m; memcpy(&m, mat8, 32);
    PtBsdf b = pt_bsdf_make(m); pt_v3 f;
    pt_bsdf_eval(b, pt_mk(wo[0], wo[1], wo[2]), pt_mk(wi[0], wi[1], wi[2]), &f, pdf);
    f3[0] = f.x; f3[1] = f.y; f3[2] = f.z;
}
int orc_bsdf_sample(const float* mat8, const float* wo, float ul, float u1, float u2, float* wi3) {
    PtMaterial m; memcpy(&m, mat8, 32);
    PtBsdf b = pt_bsdf_make(m); pt_v3 wi;
    bool ok = pt_bsdf_sample(b, pt_mk(wo[0], wo[1], wo[2]), ul, u1, u2, &wi);
    wi3[0] = wi.x; wi3[1] = wi.y; wi3[2] = wi.z; return ok;
}
void orc_sincos2pi(float u, float* s, float* c) { pt_sincos2pi(u, s, c); }
uint32_t orc_pcg(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, uint32_t* out) {
    pt_rng r = pt_rng_for(seed, pixel, sample);
    for (uint32_t i = 0; i < n; ++i) out[i] = pt_rng_next(&r);
    return n;
}
uint32_t orc_pcg_raw(uint64_t initstate, uint64_t initseq, uint32_t n, uint32_t* out) {
    pt_rng r = pt_rng_seed(initstate, initseq);
    for (uint32_t i = 0; i < n; ++i) out[i] = pt_rng_next(&r);
    return n;
}
uint64_t orc_morton(const float* c, const float* lo, const float* inv) { return pt_morton63(pt_mk(c[0], c[1], c[2]), pt_mk(lo[0], lo[1], lo[2]), pt_mk(inv[0], inv[1], inv[2])); }
void orc_sobol02(uint32_t sample, uint32_t k0, uint32_t k1, float* x, float* y) { pt_sobol02(sample, k0, k1, x, y); }
void orc_sobol02_padded(uint32_t sample, uint64_t key, float* x, float* y) { pt_sobol02_padded(sample, key, x, y); }
uint64_t orc_path_dim_key(uint64_t seed, uint32_t pixel, uint32_t bounce, uint32_t which) { return pt_path_dim_key(seed, pixel, bounce, which); }
int orc_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

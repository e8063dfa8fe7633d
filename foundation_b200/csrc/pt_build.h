// pt_build.h — per-work-item bodies of the acceleration-structure build (stages A1, A3, A5, A6 of
// SURVEY.md §8a2): Morton keys, Karras-2012 LBVH emit, BVH2 -> BVH8 collapse with outward quantisation.
// __host__ __device__ so the same bodies run inside the sm_100a kernels (pt_kernels.cuh) and inside the
// CPU emulation harness used by the non-GPU tests; the cooperative parts (radix sort, scans, reductions,
// atomics-ordered refit) live in pt_kernels.cuh only.
// No reference counterpart: the reference builds no acceleration structure.
#pragma once
#include "pt_host_shared.h"
#include "pt_layout.h"

// Build-time BVH2 (n leaves in Morton order; n-1 internal nodes, root = 0).  A "ref" is an internal node
// index (< n-1) or n-1 + sorted leaf position.  box[] is indexed by ref.
struct PtBvh2 {
    uint32_t n;
    uint32_t *left, *right, *first, *last;  // n-1 each
    uint32_t* parent;                       // 2n-1
    PtBox* box;                             // 2n-1
    float* cost;                            // 8 per internal node: the collapse plan's cost[0..6] (see pt_plan_node), [7] unused
    uint64_t* plan;                         // 1 per internal node: byte j-2 = decision for "at most j slots", j = 2..8
};

PT_HD uint32_t pt_b2_count(const PtBvh2& b, uint32_t ref) { return ref < b.n - 1 ? b.last[ref] - b.first[ref] + 1 : 1u; }
PT_HD uint32_t pt_b2_lopos(const PtBvh2& b, uint32_t ref) { return ref < b.n - 1 ? b.first[ref] : ref - (b.n - 1); }

// ---- A1: Morton key of one primitive ----------------------------------------------------------------
PT_HD pt_v3 pt_inv_extent(const float* lo, const float* hi) {
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return pt_mk(ex > 0.0f ? pt_div(2097152.0f, ex) : 0.0f, ey > 0.0f ? pt_div(2097152.0f, ey) : 0.0f, ez > 0.0f ? pt_div(2097152.0f, ez) : 0.0f);
}

// ---- A3: one internal node of the radix tree (Karras 2012, index tie-break makes all keys distinct) ---
PT_HD int pt_kdelta(const uint64_t* keys, uint32_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= (int64_t)n) return -1;
    return pt_delta(keys[i], keys[j], (uint32_t)i, (uint32_t)j);
}
PT_HD void pt_karras_node(uint32_t idx, const uint64_t* keys, const PtBvh2& b) {
    const uint32_t n = b.n;
    const int64_t i = idx;
    int64_t d = (pt_kdelta(keys, n, i, i + 1) - pt_kdelta(keys, n, i, i - 1)) > 0 ? 1 : -1;
    int dmin = pt_kdelta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (pt_kdelta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (pt_kdelta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int64_t j = i + l * d;
    int dnode = pt_kdelta(keys, n, i, j);
    int64_t s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (pt_kdelta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int64_t g = i + s * d + (d < 0 ? d : 0);
    int64_t a = i < j ? i : j, e = i < j ? j : i;
    uint32_t L = (a == g) ? (uint32_t)(n - 1 + g) : (uint32_t)g;
    uint32_t R = (e == g + 1) ? (uint32_t)(n - 1 + g + 1) : (uint32_t)(g + 1);
    b.left[idx] = L; b.right[idx] = R; b.first[idx] = (uint32_t)a; b.last[idx] = (uint32_t)e;
    b.parent[L] = idx; b.parent[R] = idx;
}

// ---- A3, bottom-up form (Apetrei 2014): where does a finished subtree go? -------------------------------
// A subtree over sorted leaves [l, r] joins the neighbour it shares the longer key prefix with: delta(r, r+1) > delta(l-1, l) -> it
// is the LEFT child of node r, else the RIGHT child of node l-1 (node id = position of its split).  k[i - koff] is the key of sorted
// position i (koff lets a kernel pass a shared-memory window of the key array).  This yields exactly the tree of pt_karras_node.
struct PtJoin { uint32_t p; bool left; int q; };   // parent id, which child we are, the parent's common-prefix length
PT_HD PtJoin pt_join(const uint64_t* k, uint32_t koff, uint32_t n, uint32_t l, uint32_t r) {
    const int dl = l > 0 ? pt_delta(k[l - 1 - koff], k[l - koff], l - 1, l) : -1;
    const int dr = r + 1 < n ? pt_delta(k[r - koff], k[r + 1 - koff], r, r + 1) : -1;
    PtJoin j; j.left = dr > dl; j.p = j.left ? r : l - 1; j.q = j.left ? dr : dl;
    return j;
}
// Does the parent's whole leaf range lie inside [tile_lo, tile_hi]?  Decided from this child's side only: the sibling of a left child
// starts at r+1 and ends before the first key that does not share more than q bits with key r+1 (keys are sorted, so one delta against
// the key just past the tile answers it); mirrored for a right child.  Both children of a node compute the same answer.
PT_HD bool pt_join_is_local(const uint64_t* k, uint32_t koff, uint32_t n, uint32_t l, uint32_t r, const PtJoin& jn, uint32_t tile_lo, uint32_t tile_hi) {
    if (jn.left) return r < tile_hi && (tile_hi + 1 >= n || pt_delta(k[r + 1 - koff], k[tile_hi + 1 - koff], r + 1, tile_hi + 1) <= jn.q);
    return l > tile_lo && (tile_lo == 0 || pt_delta(k[tile_lo - 1 - koff], k[l - 1 - koff], tile_lo - 1, l - 1) <= jn.q);
}

// ---- A5 phase 0: the collapse plan --------------------------------------------------------------------
// Which descendants of a BVH2 node become the (at most 8) children of its wide node is decided by the surface-area-cost dynamic
// programme of Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017), section 3.1:
//   cost(n, i), i = 1..7 = cheapest representation of the subtree of n as at most i children of some wide node
//   D(n, j)   = min over k = 1..j-1 of cost(left, k) + cost(right, j - k)            (j slots shared between the two subtrees)
//   cost(n, 1) = area(n) * count * PT_COST_TRI   when the subtree fits one leaf slot (count <= max_leaf)
//              = area(n) * PT_COST_NODE + D(n, 8) otherwise (n becomes a wide node of its own)
//   cost(n, i) = min(D(n, i), cost(n, i - 1))
// A greedy "open the largest child" collapse leaves the bottom of the tree half empty (4.3 children per node on the 10 M-triangle
// terrain); the plan fills 5.7 of 8 slots with one triangle per leaf and 7.5 with three: 30 % fewer nodes, 3-9 % fewer node visits.
// plan byte j-2: low nibble = k of the best split of j slots, bit 7 = "j-1 slots are as cheap" (never set for j = 8).
// The float operations and their order are part of the build's arithmetic contract (the oracle restates them).
PT_HD float pt_plan_leaf_cost(const PtBox& x) { return pt_box_area(x.lox, x.loy, x.loz, x.hix, x.hiy, x.hiz) * PT_COST_TRI; }
// cl[0..6], cr[0..6]: cost(child, 1..7) of the left / right child.  Writes cost(n, 1..7) to c[0..6] and returns the plan word.
PT_HD uint64_t pt_plan_node(const float* cl, const float* cr, float area, uint32_t count, uint32_t max_leaf, float* c) {
    float D[9];
    uint64_t plan = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 2; j <= 8; ++j) {
        int bk = 1; float bc = cl[0] + cr[j - 2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 2; k < j; ++k) { float v = cl[k - 1] + cr[j - k - 1]; if (v < bc) { bc = v; bk = k; } }
        D[j] = bc; plan |= (uint64_t)bk << (8 * (j - 2));
    }
    c[0] = count <= max_leaf ? area * ((float)count * PT_COST_TRI) : area * PT_COST_NODE + D[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 2; i <= 7; ++i) {
        if (c[i - 2] <= D[i]) { c[i - 1] = c[i - 2]; plan |= (uint64_t)0x80u << (8 * (i - 2)); } else c[i - 1] = D[i];
    }
    return plan;
}
// children of the wide node that stands for internal BVH2 node `ref`, left to right; returns their number (2..8)
PT_HD int pt_plan_children(const PtBvh2& b, uint32_t ref, uint32_t* C) {
    uint32_t st_ref[8]; uint32_t st_j[8];
    int sp = 0, nc = 0;
    uint32_t k = (uint32_t)(b.plan[ref] >> 48) & 15u;
    st_ref[sp] = b.right[ref]; st_j[sp++] = 8u - k; st_ref[sp] = b.left[ref]; st_j[sp++] = k;
    while (sp) {
        uint32_t r = st_ref[--sp], j = st_j[sp];
        if (r >= b.n - 1 || j == 1u) { C[nc++] = r; continue; }
        // the plan word and both links of r are fetched TOGETHER (the links are only needed when the plan says "split", but waiting for
        // the plan word first makes every level of the walk two dependent round trips instead of one)
        const uint64_t pw = b.plan[r];
        const uint32_t rl = b.left[r], rr = b.right[r];
        uint32_t p = (uint32_t)(pw >> (8 * (j - 2))) & 0xffu;
        while ((p & 0x80u) && j > 2u) { --j; p = (uint32_t)(pw >> (8 * (j - 2))) & 0xffu; }   // "j-1 slots are as cheap": same node, fewer slots — no reload
        if (p & 0x80u) { C[nc++] = r; continue; }                                            // j == 2 and one slot is as cheap: r stays one child
        uint32_t kk = p & 15u;
        st_ref[sp] = rr; st_j[sp++] = j - kk; st_ref[sp] = rl; st_j[sp++] = kk;
    }
    return nc;
}

// ---- A5 phase 1: the (at most 8) children of one wide node, by the plan, and their octant slots -----------
// ref: BVH2 ref the wide node stands for.  slot_ref[s] = child ref or PT_NONE.
PT_HD void pt_collapse_select(const PtBvh2& b, uint32_t ref, uint32_t max_leaf, uint32_t* slot_ref, uint32_t* n_internal, uint32_t* n_prims) {
    uint32_t C[8];
    int nc = 0;
    if (pt_b2_count(b, ref) <= max_leaf) C[nc++] = ref;
    else nc = pt_plan_children(b, ref, C);
    const PtBox nb = b.box[ref];
    float cx = (nb.lox + nb.hix) * 0.5f, cy = (nb.loy + nb.hiy) * 0.5f, cz = (nb.loz + nb.hiz) * 0.5f;
    float dx[8], dy[8], dz[8];
    for (int k = 0; k < nc; ++k) {
        const PtBox cb = b.box[C[k]];
        dx[k] = (cb.lox + cb.hix) * 0.5f - cx; dy[k] = (cb.loy + cb.hiy) * 0.5f - cy; dz[k] = (cb.loz + cb.hiz) * 0.5f - cz;
    }
    uint32_t child_done = 0, slot_used = 0;
    for (int s = 0; s < 8; ++s) slot_ref[s] = PT_NONE;
    for (int it = 0; it < nc; ++it) {
        int bk = -1, bs = -1; float bc = 0.0f;
        for (int k = 0; k < nc; ++k) {
            if (child_done >> k & 1u) continue;
            for (int s = 0; s < 8; ++s) {
                if (slot_used >> s & 1u) continue;
                float c = (((s & 4) ? dx[k] : -dx[k]) + ((s & 2) ? dy[k] : -dy[k])) + ((s & 1) ? dz[k] : -dz[k]);
                if (bk < 0 || c > bc) { bc = c; bk = k; bs = s; }
            }
        }
        slot_ref[bs] = C[bk]; child_done |= 1u << bk; slot_used |= 1u << bs;
    }
    uint32_t ni = 0, np = 0;
    for (int s = 0; s < 8; ++s) {
        if (slot_ref[s] == PT_NONE) continue;
        uint32_t cnt = pt_b2_count(b, slot_ref[s]);
        if (cnt <= max_leaf) np += cnt; else ni += 1;
    }
    *n_internal = ni; *n_prims = np;
}

// ---- A5 phase 2: write the 80-byte node, the next level's refs and the leaf sequence ---------------------
PT_HD void pt_collapse_emit(const PtBvh2& b, uint32_t ref, const uint32_t* slot_ref, uint32_t max_leaf, float pad, uint32_t child_base,
                            uint32_t prim_base, PtNode8* out, uint32_t* next_refs /* at this node's first child */, uint32_t* leaf_seq /* global */) {
    const PtBox nb = b.box[ref];
    PtNode8 nd;
    float p[3] = {nb.lox - pad, nb.loy - pad, nb.loz - pad};
    float hi[3] = {nb.hix + pad, nb.hiy + pad, nb.hiz + pad};
    float inv[3]; uint32_t e[3];
    for (int k = 0; k < 3; ++k) { e[k] = pt_quant_exp(hi[k] - p[k]); inv[k] = pt_u2f((254u - e[k]) << 23); }
    nd.px = p[0]; nd.py = p[1]; nd.pz = p[2];
    nd.ex = (uint8_t)e[0]; nd.ey = (uint8_t)e[1]; nd.ez = (uint8_t)e[2]; nd.imask = 0;
    nd.child_base = child_base; nd.tri_base = prim_base;
    uint32_t off = 0, nint = 0;
    for (int s = 0; s < 8; ++s) {
        uint32_t r = slot_ref[s];
        if (r == PT_NONE) {
            nd.meta[s] = 0;
            nd.qlox[s] = nd.qloy[s] = nd.qloz[s] = 255; nd.qhix[s] = nd.qhiy[s] = nd.qhiz[s] = 0;
            continue;
        }
        const PtBox cb = b.box[r];
        nd.qlox[s] = (uint8_t)pt_quant_lo(cb.lox - pad, p[0], inv[0]); nd.qhix[s] = (uint8_t)pt_quant_hi(cb.hix + pad, p[0], inv[0]);
        nd.qloy[s] = (uint8_t)pt_quant_lo(cb.loy - pad, p[1], inv[1]); nd.qhiy[s] = (uint8_t)pt_quant_hi(cb.hiy + pad, p[1], inv[1]);
        nd.qloz[s] = (uint8_t)pt_quant_lo(cb.loz - pad, p[2], inv[2]); nd.qhiz[s] = (uint8_t)pt_quant_hi(cb.hiz + pad, p[2], inv[2]);
        uint32_t cnt = pt_b2_count(b, r);
        if (cnt <= max_leaf) {
            uint32_t fp = pt_b2_lopos(b, r);
            nd.meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | off);
            for (uint32_t q = 0; q < cnt; ++q) leaf_seq[prim_base + off + q] = fp + q;
            off += cnt;
        } else {
            nd.imask |= (uint8_t)(1u << s);
            nd.meta[s] = (uint8_t)(0x20u | (24u + (uint32_t)s));
            next_refs[nint++] = r;
        }
    }
    *out = nd;
}

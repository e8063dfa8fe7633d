// pt_emu.cpp — CPU emulation harness for the product's per-work-item device bodies (pt_build.h,
// pt_traverse.h).  TEST-ONLY: lets the non-GPU test tier execute the exact control flow the sm_100a
// kernels run (Karras emit, BVH8 collapse, group-stack traversal) and compare it with the independent
// oracle before any GPU time is spent.  It is never built or loaded by the product.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../foundation_b200/csrc/pt_build.h"
#include "../../foundation_b200/csrc/pt_traverse.h"

// Generic LBVH -> BVH8 build over n primitives given boxes (6 floats) and centroids (3 floats).
// Outputs: nodes (capacity n * 80 B), leaf_seq (n), order (n).  Returns the number of nodes.
// tile == 0: Karras emit (pt_karras_node) + sequential second-arriver refit.  tile > 0: the product's default, bottom-up agglomerative
// build (pt_join / pt_join_is_local) emulated with the kernels' structure — per tile of `tile` sorted leaves the local nodes are
// finished first (k_refit_agg), the roots of the tile-local subtrees then climb through the "global" records (k_refit_agg_up).
static uint32_t emu_build_impl(const float* pbox, const float* cent, uint32_t n, uint32_t max_leaf, float* out_lo, float* out_hi, void* nodes_out,
                               uint32_t* leaf_seq, uint32_t* order_out, uint32_t tile) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) { lo[k] = pt_min(lo[k], pbox[6 * i + k]); hi[k] = pt_max(hi[k], pbox[6 * i + 3 + k]); }
    memcpy(out_lo, lo, 12); memcpy(out_hi, hi, 12);
    pt_v3 inv = pt_inv_extent(lo, hi);
    std::vector<std::pair<uint64_t, uint32_t>> kv(n);
    for (uint32_t i = 0; i < n; ++i) kv[i] = {pt_morton63(pt_mk(cent[3 * i], cent[3 * i + 1], cent[3 * i + 2]), pt_mk(lo[0], lo[1], lo[2]), inv), i};
    std::stable_sort(kv.begin(), kv.end());
    std::vector<uint64_t> keys(n);
    for (uint32_t i = 0; i < n; ++i) { keys[i] = kv[i].first; order_out[i] = kv[i].second; }
    std::vector<uint32_t> left(n), right(n), first(n), last(n), parent(2 * (size_t)n), flags(n, 0);
    std::vector<PtBox> box(2 * (size_t)n);
    std::vector<float> cost(8 * (size_t)n, 0.0f);
    std::vector<uint64_t> plan(n, 0);
    PtBvh2 b{n, left.data(), right.data(), first.data(), last.data(), parent.data(), box.data(), cost.data(), plan.data()};
    uint32_t root = 0;
    auto finish = [&](uint32_t cur) {      // both children of `cur` are linked and finished: box, costs, plan
        PtBox l = box[left[cur]], r = box[right[cur]];
        box[cur] = PtBox{pt_min(l.lox, r.lox), pt_min(l.loy, r.loy), pt_min(l.loz, r.loz), pt_max(l.hix, r.hix), pt_max(l.hiy, r.hiy), pt_max(l.hiz, r.hiz)};
        float cl[7], cr[7];
        for (int i = 0; i < 7; ++i) {
            cl[i] = left[cur] >= n - 1 ? pt_plan_leaf_cost(l) : cost[8 * (size_t)left[cur] + i];
            cr[i] = right[cur] >= n - 1 ? pt_plan_leaf_cost(r) : cost[8 * (size_t)right[cur] + i];
        }
        const PtBox u = box[cur];
        plan[cur] = pt_plan_node(cl, cr, pt_box_area(u.lox, u.loy, u.loz, u.hix, u.hiy, u.hiz), last[cur] - first[cur] + 1u, max_leaf, &cost[8 * (size_t)cur]);
    };
    for (uint32_t j = 0; j < n; ++j) {
        const float* p = pbox + 6 * (size_t)order_out[j];
        box[n - 1 + j] = PtBox{p[0], p[1], p[2], p[3], p[4], p[5]};
    }
    if (tile == 0) {
        for (uint32_t i = 0; i + 1 < n; ++i) pt_karras_node(i, keys.data(), b);
        // refit: sequential emulation of the second-arriver rule
        for (uint32_t j = 0; j < n && n > 1; ++j) {
            uint32_t cur = parent[n - 1 + j];
            for (;;) {
                if (flags[cur]++ == 0) break;
                finish(cur);
                if (cur == 0) break;
                cur = parent[cur];
            }
        }
    } else if (n > 1) {
        struct Sub { uint32_t ref, l, r; };
        std::vector<Sub> up;                               // roots of the tile-local subtrees (the kernels' up_list)
        bool have_root = false;
        for (uint32_t tile_lo = 0; tile_lo < n; tile_lo += tile) {
            const uint32_t tile_hi = std::min(n, tile_lo + tile) - 1u;
            std::vector<Sub> work;
            for (uint32_t j = tile_lo; j <= tile_hi; ++j) work.push_back({n - 1 + j, j, j});
            std::vector<uint32_t> lflag(tile, 0);          // the tile's shared-memory arrival counters
            while (!work.empty()) {
                Sub s = work.back(); work.pop_back();
                if (s.l == 0 && s.r == n - 1) { root = s.ref; have_root = true; continue; }
                const PtJoin jn = pt_join(keys.data(), 0u, n, s.l, s.r);
                if (!pt_join_is_local(keys.data(), 0u, n, s.l, s.r, jn, tile_lo, tile_hi)) { up.push_back(s); continue; }
                if (jn.p < tile_lo || jn.p > tile_hi) return 0xffffffffu;            // a local parent must have a slot in the tile
                if (jn.left) { left[jn.p] = s.ref; first[jn.p] = s.l; } else { right[jn.p] = s.ref; last[jn.p] = s.r; }
                if (lflag[jn.p - tile_lo]++ == 1) { finish(jn.p); work.push_back({jn.p, first[jn.p], last[jn.p]}); }
            }
        }
        for (const Sub& s0 : up) {                          // upper levels: global arrival flags
            Sub s = s0;
            for (;;) {
                if (s.l == 0 && s.r == n - 1) { root = s.ref; have_root = true; break; }
                const PtJoin jn = pt_join(keys.data(), 0u, n, s.l, s.r);
                if (jn.left) { left[jn.p] = s.ref; first[jn.p] = s.l; } else { right[jn.p] = s.ref; last[jn.p] = s.r; }
                if (flags[jn.p]++ == 0) break;
                finish(jn.p);
                s = {jn.p, first[jn.p], last[jn.p]};
            }
        }
        if (!have_root) return 0xfffffffeu;
    }
    float pad = pt_pad_for(lo, hi);
    PtNode8* nodes = (PtNode8*)nodes_out;
    std::vector<uint32_t> level{root}, next, slots, ni, np;
    uint32_t level_start = 0, prim_total = 0;
    while (!level.empty()) {
        size_t m = level.size();
        slots.assign(8 * m, PT_NONE); ni.assign(m, 0); np.assign(m, 0);
        for (size_t w = 0; w < m; ++w) pt_collapse_select(b, level[w], max_leaf, &slots[8 * w], &ni[w], &np[w]);
        std::vector<uint32_t> si(m), sp(m); uint32_t ti = 0, tp = 0;
        for (size_t w = 0; w < m; ++w) { si[w] = ti; ti += ni[w]; sp[w] = tp; tp += np[w]; }
        next.assign(ti, 0);
        uint32_t next_start = level_start + (uint32_t)m;
        for (size_t w = 0; w < m; ++w)
            pt_collapse_emit(b, level[w], &slots[8 * w], max_leaf, pad, next_start + si[w], prim_total + sp[w], &nodes[level_start + w], next.data() + si[w], leaf_seq);
        prim_total += tp; level_start = next_start; level.swap(next);
    }
    return level_start;
}

extern "C" {
uint32_t emu_build(const float* pbox, const float* cent, uint32_t n, uint32_t max_leaf, float* out_lo, float* out_hi, void* nodes_out, uint32_t* leaf_seq,
                   uint32_t* order_out) {
    return emu_build_impl(pbox, cent, n, max_leaf, out_lo, out_hi, nodes_out, leaf_seq, order_out, 0u);
}
uint32_t emu_build_agglomerative(const float* pbox, const float* cent, uint32_t n, uint32_t max_leaf, float* out_lo, float* out_hi, void* nodes_out,
                                 uint32_t* leaf_seq, uint32_t* order_out, uint32_t tile) {
    return emu_build_impl(pbox, cent, n, max_leaf, out_lo, out_hi, nodes_out, leaf_seq, order_out, tile ? tile : 256u);
}
}

struct EmuRay { float o[3], tmin, d[3], tmax; };
struct EmuHit { float t, u, v; uint32_t prim; };

extern "C" {
// mode bit 0: any-hit; two_level != 0: instanced scene.  counters[3] accumulates nodes / tris / instances.
int emu_trace(const void* nodes, const void* tris, const void* instances, int two_level, const void* rays_, uint64_t n, void* hits_, uint32_t* inst_out,
              uint8_t* occ, int any, uint64_t* counters) {
    PtSceneView sc{(const PtU4*)nodes, (const PtU4*)tris, (const PtU4*)instances, 0u, 0u};   // oracle layout: TLAS first
    const EmuRay* rays = (const EmuRay*)rays_; EmuHit* hits = (EmuHit*)hits_;
    PtCount c{0, 0, 0};
    int ok = 1;
    for (uint64_t i = 0; i < n; ++i) {
        PtHitRec h;
        pt_v3 o = pt_mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d = pt_mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        bool fine;
        if (two_level) fine = any ? pt_traverse<true, true>(sc, o, d, rays[i].tmin, rays[i].tmax, &h, c) : pt_traverse<false, true>(sc, o, d, rays[i].tmin, rays[i].tmax, &h, c);
        else fine = any ? pt_traverse<true, false>(sc, o, d, rays[i].tmin, rays[i].tmax, &h, c) : pt_traverse<false, false>(sc, o, d, rays[i].tmin, rays[i].tmax, &h, c);
        if (!fine) ok = 0;
        if (any) { occ[i] = h.prim != PT_NONE; continue; }
        if (h.prim == PT_NONE) { hits[i] = EmuHit{INFINITY, 0, 0, PT_NONE}; if (inst_out) inst_out[i] = PT_NONE; }
        else {
            float u, v; uint32_t mat;
            if (two_level) pt_hit_bary<true>(sc, h, o, d, &u, &v, &mat); else pt_hit_bary<false>(sc, h, o, d, &u, &v, &mat);
            hits[i] = EmuHit{h.t, u, v, h.prim}; if (inst_out) inst_out[i] = h.inst;
        }
    }
    if (counters) { counters[0] = c.nodes; counters[1] = c.tris; counters[2] = c.insts; }
    return ok;
}
}

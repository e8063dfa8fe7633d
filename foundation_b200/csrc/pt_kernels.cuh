// pt_kernels.cuh — every sm_100a kernel of the path-tracing hot path.
// Hand-written CUDA; no CUB/Thrust, no OptiX, no tensor cores (nothing here is a dense contraction).
// Compiled with --fmad=false: fused multiply-adds happen exactly where pt_math.h says pt_fma.
// The reference has no counterpart for any kernel in this file (it launches no compute work at all:
// src/Platform/RHI/Command.hpp:38-115 has no Dispatch); the per-item arithmetic is in the shared headers.
#pragma once
#include <cuda_runtime.h>

#include "pt_build.h"
#include "pt_shading.h"
#include "pt_traverse.h"

#define PT_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pt_f2ord(float f) {  // order-preserving float -> uint map for atomicMin/Max
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float pt_ord2f(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}
__device__ __forceinline__ PtU4 pt_ldg4(const PtU4* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    PtU4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
__device__ __forceinline__ uint32_t pt_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t pt_gtid() { return blockIdx.x * blockDim.x + threadIdx.x; }
__device__ __forceinline__ uint32_t pt_gsize() { return gridDim.x * blockDim.x; }

// ---------------------------------------------------------------------------------------------------
// raw mesh access: positions with byte stride, u16 / u32 / no indices (formats of RHIResourceFormat,
// mos9527/Foundation src/Platform/RHI/Common.hpp:18-27)
// ---------------------------------------------------------------------------------------------------
struct PtMeshRaw {
    const uint8_t* pos; uint32_t stride; const void* idx; uint32_t idx_fmt; uint32_t ntris; const uint32_t* mat;
};
__device__ __forceinline__ pt_v3 pt_load_vertex(const PtMeshRaw& m, uint32_t v) {
    const float* p = reinterpret_cast<const float*>(m.pos + (size_t)v * m.stride);
    return pt_mk(p[0], p[1], p[2]);
}
__device__ __forceinline__ void pt_load_tri(const PtMeshRaw& m, uint32_t i, pt_v3* a, pt_v3* b, pt_v3* c) {
    uint32_t i0, i1, i2;
    if (m.idx_fmt == 32) { const uint32_t* q = (const uint32_t*)m.idx + 3 * (size_t)i; i0 = q[0]; i1 = q[1]; i2 = q[2]; }
    else if (m.idx_fmt == 16) { const uint16_t* q = (const uint16_t*)m.idx + 3 * (size_t)i; i0 = q[0]; i1 = q[1]; i2 = q[2]; }
    else { i0 = 3 * i; i1 = i0 + 1; i2 = i0 + 2; }
    *a = pt_load_vertex(m, i0); *b = pt_load_vertex(m, i1); *c = pt_load_vertex(m, i2);
}

// ---------------------------------------------------------------------------------------------------
// A1: primitive boxes + scene bounds (warp-shuffle reduce, one ordered-uint atomic per warp and plane)
// bounds[0..2] = lo (init 0xffffffff), bounds[3..5] = hi (init 0)
// ---------------------------------------------------------------------------------------------------
// Per-thread running bounds -> warp shuffle reduce -> shared-memory block reduce -> 6 atomics per BLOCK
// (the first version issued 6 same-address atomics per warp and was atomics-bound: 1.25 ms for 10 M triangles).
struct PtOrdBounds { uint32_t lo[3], hi[3]; };
__device__ __forceinline__ void pt_ord_init(PtOrdBounds& b) { for (int k = 0; k < 3; ++k) { b.lo[k] = 0xffffffffu; b.hi[k] = 0u; } }
__device__ __forceinline__ void pt_ord_grow(PtOrdBounds& b, const PtBox& bx) {
    b.lo[0] = min(b.lo[0], pt_f2ord(bx.lox)); b.lo[1] = min(b.lo[1], pt_f2ord(bx.loy)); b.lo[2] = min(b.lo[2], pt_f2ord(bx.loz));
    b.hi[0] = max(b.hi[0], pt_f2ord(bx.hix)); b.hi[1] = max(b.hi[1], pt_f2ord(bx.hiy)); b.hi[2] = max(b.hi[2], pt_f2ord(bx.hiz));
}
__device__ __forceinline__ void pt_block_reduce_bounds(PtOrdBounds b, uint32_t* bounds) {   // all threads of the block must call
    __shared__ uint32_t s_lo[3], s_hi[3];
    if (threadIdx.x == 0) { for (int k = 0; k < 3; ++k) { s_lo[k] = 0xffffffffu; s_hi[k] = 0u; } }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) { b.lo[k] = __reduce_min_sync(PT_FULL, b.lo[k]); b.hi[k] = __reduce_max_sync(PT_FULL, b.hi[k]); }
    if (pt_lane() == 0) { for (int k = 0; k < 3; ++k) { atomicMin(&s_lo[k], b.lo[k]); atomicMax(&s_hi[k], b.hi[k]); } }
    __syncthreads();
    if (threadIdx.x == 0) { for (int k = 0; k < 3; ++k) { atomicMin(&bounds[k], s_lo[k]); atomicMax(&bounds[3 + k], s_hi[k]); } }
}

__global__ void __launch_bounds__(256) k_tri_boxes(PtMeshRaw m, PtBox* prim_box, uint32_t* bounds) {
    PtOrdBounds ob; pt_ord_init(ob);
    for (uint32_t i = pt_gtid(); i < m.ntris; i += pt_gsize()) {
        pt_v3 a, b, c;
        pt_load_tri(m, i, &a, &b, &c);
        PtBox bx;
        bx.lox = pt_min(pt_min(a.x, b.x), c.x); bx.loy = pt_min(pt_min(a.y, b.y), c.y); bx.loz = pt_min(pt_min(a.z, b.z), c.z);
        bx.hix = pt_max(pt_max(a.x, b.x), c.x); bx.hiy = pt_max(pt_max(a.y, b.y), c.y); bx.hiz = pt_max(pt_max(a.z, b.z), c.z);
        prim_box[i] = bx;
        pt_ord_grow(ob, bx);
    }
    pt_block_reduce_bounds(ob, bounds);
}

// device-resident build parameters derived from the reduced bounds (no host round trip)
struct PtBuildParams { float lo[3], hi[3]; float pad; float inv[3]; };
__global__ void k_build_params(const uint32_t* bounds, PtBuildParams* bp) {
    if (pt_gtid() != 0) return;
    for (int k = 0; k < 3; ++k) { bp->lo[k] = pt_ord2f(bounds[k]); bp->hi[k] = pt_ord2f(bounds[3 + k]); }
    bp->pad = pt_pad_for(bp->lo, bp->hi);
    pt_v3 inv = pt_inv_extent(bp->lo, bp->hi);
    bp->inv[0] = inv.x; bp->inv[1] = inv.y; bp->inv[2] = inv.z;
}

__global__ void __launch_bounds__(256) k_morton_tris(PtMeshRaw m, const PtBuildParams* bp, uint64_t* keys, uint32_t* vals) {
    pt_v3 lo = pt_mk(bp->lo[0], bp->lo[1], bp->lo[2]), inv = pt_mk(bp->inv[0], bp->inv[1], bp->inv[2]);
    for (uint32_t i = pt_gtid(); i < m.ntris; i += pt_gsize()) {
        pt_v3 a, b, c;
        pt_load_tri(m, i, &a, &b, &c);
        keys[i] = pt_morton63(pt_tri_centroid(a, b, c), lo, inv);
        vals[i] = i;
    }
}
__global__ void __launch_bounds__(256) k_morton_boxes(const PtBox* prim_box, uint32_t n, const PtBuildParams* bp, uint64_t* keys, uint32_t* vals) {
    pt_v3 lo = pt_mk(bp->lo[0], bp->lo[1], bp->lo[2]), inv = pt_mk(bp->inv[0], bp->inv[1], bp->inv[2]);
    for (uint32_t i = pt_gtid(); i < n; i += pt_gsize()) {
        PtBox b = prim_box[i];
        pt_v3 c = pt_mk((b.lox + b.hix) * 0.5f, (b.loy + b.hiy) * 0.5f, (b.loz + b.hiz) * 0.5f);
        keys[i] = pt_morton63(c, lo, inv);
        vals[i] = i;
    }
}

// ---------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (single pass over 4096-element tiles, decoupled look-back)
// ---------------------------------------------------------------------------------------------------
#define PT_SCAN_THREADS 256
#define PT_SCAN_ITEMS 16
#define PT_SCAN_CHUNK (PT_SCAN_THREADS * PT_SCAN_ITEMS)

__device__ __forceinline__ uint32_t pt_block_excl_scan(uint32_t v, uint32_t* total) {  // blockDim.x multiple of 32, <= 1024
    __shared__ uint32_t warp_sums[33];
    const uint32_t lane = pt_lane(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(PT_FULL, inc, o); if (lane >= (uint32_t)o) inc += t; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarps ? warp_sums[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(PT_FULL, wi, o); if (lane >= (uint32_t)o) wi += t; }
        warp_sums[lane] = wi - w;             // exclusive prefix of the warp sums
        if (lane == 31) warp_sums[32] = wi;   // block total
    }
    __syncthreads();
    uint32_t res = warp_sums[warp] + inc - v;
    *total = warp_sums[32];
    __syncthreads();                          // warp_sums is reused by the next call
    return res;
}

// Single-pass exclusive scan (decoupled look-back, Merrill & Garland 2016): one launch instead of chunks + sums + add, every element read
// and written once.  A tile takes a ticket (so all its predecessors are running or done: the spin below always makes progress), scans
// its 4096 elements, publishes its aggregate, then warp 0 walks back over the predecessors' 64-bit status words, 32 at a time — flag in the top two
// bits (1 = aggregate, 2 = inclusive prefix), value in the low 32 — until it meets an inclusive prefix, and publishes its own.
// state[0 .. tiles) and the ticket at state[tiles] must be zero at launch.
__global__ void __launch_bounds__(PT_SCAN_THREADS) k_scan_lookback(const uint32_t* in, uint32_t* out, uint32_t n, unsigned long long* state, uint32_t tiles,
                                                                    uint32_t* grand_total) {
    __shared__ uint32_t s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(reinterpret_cast<uint32_t*>(state + tiles), 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * PT_SCAN_CHUNK + threadIdx.x * PT_SCAN_ITEMS;
    uint32_t v[PT_SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < PT_SCAN_ITEMS; ++k) { uint32_t i = base + k; v[k] = i < n ? in[i] : 0; sum += v[k]; }
    uint32_t total;
    uint32_t excl = pt_block_excl_scan(sum, &total);
    if (threadIdx.x < 32) {     // warp 0: publish the aggregate, then look back 32 predecessors at a time
        const uint32_t lane = threadIdx.x;
        if (lane == 0) atomicExch(&state[tile], ((tile == 0 ? 2ull : 1ull) << 62) | total);
        uint32_t prefix = 0;
        if (tile > 0) {
            for (long long p = (long long)tile - 1 - lane;; p -= 32) {
                unsigned long long w = 2ull << 62;                       // before tile 0: an inclusive prefix of 0
                if (p >= 0) do { w = *reinterpret_cast<volatile unsigned long long*>(&state[p]); } while ((w >> 62) == 0ull);
                const uint32_t done = __ballot_sync(PT_FULL, (w >> 62) == 2ull);     // lanes that found an inclusive prefix (lane 0 = nearest tile)
                const uint32_t upto = done ? (uint32_t)__ffs(done) - 1u : 31u;
                prefix += __reduce_add_sync(PT_FULL, lane <= upto ? (uint32_t)w : 0u);
                if (done) break;
            }
            if (lane == 0) atomicExch(&state[tile], (2ull << 62) | (unsigned long long)(prefix + total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == tiles - 1 && grand_total) *grand_total = prefix + total;
        }
    }
    __syncthreads();
    excl += s_prefix;
#pragma unroll
    for (int k = 0; k < PT_SCAN_ITEMS; ++k) { uint32_t i = base + k; if (i < n) out[i] = excl; excl += v[k]; }
}

// ---------------------------------------------------------------------------------------------------
// A2: stable LSD radix sort of (uint64 key, uint32 value), 8 bits per pass.
// Per pass: per-tile digit histogram -> exclusive scan over [digit][tile] -> stable scatter with a tile-local sort.
// ---------------------------------------------------------------------------------------------------
// Lanes of the warp whose 8-bit digit equals this lane's (only used when PT_RANK_ATOMIC_OR = 0).  PT_MATCH_BALLOT = 1: eight ballots, one per digit
// bit, intersected in registers (the form CUB's radix ranking uses); 0: one MATCH.ANY.  Measured on the 10 M-pair three-kernel sort, same box
// (profiles/r02_ab_sort_ranking.log): MATCH.ANY 1.066 ms, ballots 1.056 ms, shared-memory atomicOr (pt_rank_round) 1.043 ms.
#ifndef PT_MATCH_BALLOT
#define PT_MATCH_BALLOT 1
#endif
__device__ __forceinline__ uint32_t pt_match_digit8(uint32_t d) {
#if PT_MATCH_BALLOT
    uint32_t peers = PT_FULL;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(PT_FULL, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
#else
    return __match_any_sync(PT_FULL, d);
#endif
}
// One ranking round of a warp: every lane holds a key with digit d; returns the key's rank among the warp's keys of that digit seen so far
// (cnt_w: the warp's 256 running counts in shared memory).  PT_RANK_ATOMIC_OR = 1: the peer set is built in shared memory — every lane ORs its
// lane bit into the digit's word of mask_w (all zero between rounds), then reads the word back — instead of by warp votes: OR is order-independent,
// so the result is as deterministic as a vote (CUB's WARP_MATCH_ATOMIC_OR).
#ifndef PT_RANK_ATOMIC_OR
#define PT_RANK_ATOMIC_OR 1
#endif
__device__ __forceinline__ uint32_t pt_rank_round(uint32_t d, uint32_t* cnt_w, uint32_t* mask_w, uint32_t lane, uint32_t lt) {
#if PT_RANK_ATOMIC_OR
    atomicOr(&mask_w[d], 1u << lane);
    __syncwarp();
    const uint32_t peers = mask_w[d];
#else
    (void)mask_w;
    const uint32_t peers = pt_match_digit8(d);
#endif
    const uint32_t old = cnt_w[d];
    __syncwarp();
    if (lane == (uint32_t)__ffs(peers) - 1u) {
        cnt_w[d] = old + (uint32_t)__popc(peers);
#if PT_RANK_ATOMIC_OR
        mask_w[d] = 0u;
#endif
    }
    __syncwarp();
    return old + (uint32_t)__popc(peers & lt);       // rank of this key among the warp's keys with the same digit
}
#define PT_RS_THREADS 256
#ifndef PT_RS_ROUNDS
#define PT_RS_ROUNDS 5      // keys per thread: 1280-key tiles (measured on the 10 M-key sort: 4 -> 1.23 ms, 5 -> 1.13, 6 -> 1.17, 8 -> 1.31)
#endif
#define PT_RS_TILE (PT_RS_THREADS * PT_RS_ROUNDS)

__global__ void __launch_bounds__(PT_RS_THREADS) k_rs_hist(const uint64_t* keys, uint32_t n, int shift, uint32_t* tile_hist, uint32_t num_tiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * PT_RS_TILE;
    uint64_t k[PT_RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < PT_RS_ROUNDS; ++r) {       // all loads of the thread in flight before the first shared-memory atomic
        uint32_t i = base + r * PT_RS_THREADS + threadIdx.x;
        k[r] = i < n ? __ldg(keys + i) : 0ull;
    }
#pragma unroll
    for (int r = 0; r < PT_RS_ROUNDS; ++r) {
        uint32_t i = base + r * PT_RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(k[r] >> shift) & 255u], 1u);
    }
    __syncthreads();
    tile_hist[threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

// Scatter of one pass.  Each CTA sorts its tile (PT_RS_TILE keys) locally first, so global writes are runs of consecutive keys per digit
// (the first version wrote every 8-byte key to its own 32-byte sector and reached 15 % of HBM peak):
//   1. warp w loads its 32 * PT_RS_KEYS consecutive keys of the tile, PT_RS_KEYS per lane, every load coalesced (order in the tile = w, i, lane);
//   2. rank inside the warp: match_any groups equal digits, a per-warp shared counter carries the running count;
//   3. one pass over the 8 warp counters per digit + a block scan over the 256 digits give every key its tile-local position;
//   4. keys go to shared memory at that position, are read back in order and written to global_offset[digit] + run index;
//   5. values take the same route through the same shared buffer.
#define PT_RS_KEYS (PT_RS_TILE / PT_RS_THREADS)
__global__ void __launch_bounds__(PT_RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout,
                                                              uint32_t* __restrict__ vout, uint32_t n, int shift, const uint32_t* __restrict__ tile_off,
                                                              uint32_t num_tiles) {
    __shared__ uint32_t wh[PT_RS_THREADS / 32][256];
#if PT_RANK_ATOMIC_OR
    __shared__ uint32_t wm[PT_RS_THREADS / 32][256];
#else
    uint32_t (*wm)[256] = wh;
#endif
    __shared__ uint32_t dstart[256], goff[256];
    __shared__ uint64_t sbuf[PT_RS_TILE];
    uint32_t* svals = reinterpret_cast<uint32_t*>(sbuf);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t base = blockIdx.x * PT_RS_TILE;
    const uint32_t n_valid = min((uint32_t)PT_RS_TILE, n - base);
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int w = 0; w < PT_RS_THREADS / 32; ++w) { wh[w][tid] = 0; if (PT_RANK_ATOMIC_OR) wm[w][tid] = 0; }
    uint64_t key[PT_RS_KEYS];
    uint32_t lp[PT_RS_KEYS];
#pragma unroll
    for (int i = 0; i < PT_RS_KEYS; ++i) {
        uint32_t j = warp * (32 * PT_RS_KEYS) + i * 32 + lane;
        key[i] = j < n_valid ? kin[base + j] : ~0ull;     // padding sorts to the very end of the tile and is never written
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT_RS_KEYS; ++i) {
        uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        lp[i] = pt_rank_round(d, wh[warp], wm[warp], lane, lt);
    }
    __syncthreads();
    {   // thread `tid` owns digit `tid`: exclusive prefix over the warps, then over the digits
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < PT_RS_THREADS / 32; ++w) { uint32_t c = wh[w][tid]; wh[w][tid] = acc; acc += c; }
        uint32_t total;
        dstart[tid] = pt_block_excl_scan(acc, &total);
        goff[tid] = tile_off[tid * num_tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT_RS_KEYS; ++i) {
        uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        lp[i] += dstart[d] + wh[warp][d];
        sbuf[lp[i]] = key[i];
    }
    __syncthreads();
    uint32_t outpos[PT_RS_KEYS];
#pragma unroll
    for (int k = 0; k < PT_RS_KEYS; ++k) {
        uint32_t j = k * PT_RS_THREADS + tid;
        outpos[k] = 0xffffffffu;
        if (j < n_valid) {
            uint64_t kk = sbuf[j];
            uint32_t d = (uint32_t)(kk >> shift) & 255u;
            outpos[k] = goff[d] + (j - dstart[d]);
            kout[outpos[k]] = kk;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT_RS_KEYS; ++i) {
        uint32_t j = warp * (32 * PT_RS_KEYS) + i * 32 + lane;
        if (j < n_valid) svals[lp[i]] = vin[base + j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PT_RS_KEYS; ++k) {
        uint32_t j = k * PT_RS_THREADS + tid;
        if (j < n_valid) vout[outpos[k]] = svals[j];
    }
}

// ---------------------------------------------------------------------------------------------------
// A2, one-sweep form (Adinets & Merrill, "Onesweep: a faster least significant digit radix sort for GPUs", 2022): the digit histograms of
// ALL eight passes are taken in one up-front read of the keys, and every pass is ONE kernel that reads each (key, value) once and writes
// it once — the tile's global offsets come from a per-digit decoupled look-back over the preceding tiles' published counts instead of a
// separate histogram kernel + scan kernel per pass (round 1: 3 kernels and 2 extra key reads per pass, 1.13 ms for 10 M pairs).
// ---------------------------------------------------------------------------------------------------
#ifndef PT_OS_KPT
#define PT_OS_KPT 10           // keys per thread: 2560-key tiles at 64 registers / 4 CTAs per SM (12 keys at 80 registers / 3 CTAs is level: 0.775 vs 0.778 ms for 10 M pairs)
#endif
#ifndef PT_OS_LB
#define PT_OS_LB 8             // predecessors inspected per look-back round trip
#endif
#ifndef PT_OS_EARLY_VALS
#define PT_OS_EARLY_VALS 1
#endif
#ifndef PT_OS_MIN_BLOCKS
#define PT_OS_MIN_BLOCKS 4
#endif
#define PT_OS_THREADS 256
#define PT_OS_TILE (PT_OS_THREADS * PT_OS_KPT)
#define PT_OS_FLAG_AGG 0x40000000u
#define PT_OS_FLAG_INC 0x80000000u
#define PT_OS_VALUE 0x3fffffffu

// hist[p][d] += number of keys whose digit p equals d.  Sorted-ish input (triangles in mesh order) makes the high digits uniform across a
// warp: those go through one aggregated atomic per warp, the rest through per-lane shared-memory atomics.
__global__ void __launch_bounds__(256) k_rs_hist_all(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[8][256];
    for (int p = 0; p < 8; ++p) h[p][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t rounds = (n + pt_gsize() - 1) / pt_gsize();
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t i = r * pt_gsize() + pt_gtid();
        const bool valid = i < n;
        const uint64_t k = valid ? keys[i] : 0ull;
        const uint32_t live = __ballot_sync(PT_FULL, valid);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const uint32_t d = (uint32_t)(k >> (8 * p)) & 255u;
            const uint32_t d0 = __shfl_sync(PT_FULL, d, __ffs(live | 0x80000000u) - 1);
            if (live == PT_FULL && __all_sync(PT_FULL, d == d0)) { if (pt_lane() == 0) atomicAdd(&h[p][d0], 32u); }
            else if (valid) atomicAdd(&h[p][d], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < 8; ++p) { const uint32_t c = h[p][threadIdx.x]; if (c) atomicAdd(&hist[p * 256 + threadIdx.x], c); }
}
// one block per pass: counts -> exclusive digit offsets, in place
__global__ void __launch_bounds__(256) k_rs_hist_scan(uint32_t* hist) {
    uint32_t total;
    const uint32_t c = hist[blockIdx.x * 256 + threadIdx.x];
    hist[blockIdx.x * 256 + threadIdx.x] = pt_block_excl_scan(c, &total);
}

// One pass.  Tiles take tickets (tile t only ever waits for tiles < t, which are running or done), rank their keys exactly like
// k_rs_scatter (match_any against per-warp counters), then thread d publishes the tile's count of digit d — flag AGG — walks back over the
// predecessors' words of digit d until it meets an inclusive prefix, publishes its own inclusive prefix — flag INC — and so knows where
// the tile's run of digit d starts in the output.  status: tiles x 256 words, zero at launch; *ticket zero at launch.
__global__ void __launch_bounds__(PT_OS_THREADS, PT_OS_MIN_BLOCKS) k_rs_onesweep(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout,
                                                               uint32_t* __restrict__ vout, uint32_t n, int shift, const uint32_t* __restrict__ digit_base,
                                                               uint32_t* status, uint32_t* ticket) {
    __shared__ uint32_t wh[PT_OS_THREADS / 32][256];
#if PT_RANK_ATOMIC_OR
    __shared__ uint32_t wm[PT_OS_THREADS / 32][256];
#else
    uint32_t (*wm)[256] = wh;
#endif
    __shared__ uint32_t dstart[256], goff[256];
    __shared__ uint64_t sbuf[PT_OS_TILE];
    __shared__ uint32_t s_tile;
    uint32_t* svals = reinterpret_cast<uint32_t*>(sbuf);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < PT_OS_THREADS / 32; ++w) { wh[w][tid] = 0; if (PT_RANK_ATOMIC_OR) wm[w][tid] = 0; }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * PT_OS_TILE;
    const uint32_t n_valid = min((uint32_t)PT_OS_TILE, n - base);
    const uint32_t lt = (1u << lane) - 1u;
    uint64_t key[PT_OS_KPT];
    uint32_t lp[PT_OS_KPT];
#pragma unroll
    for (int i = 0; i < PT_OS_KPT; ++i) {
        const uint32_t j = warp * (32 * PT_OS_KPT) + i * 32 + lane;
        key[i] = j < n_valid ? __ldcs(kin + base + j) : ~0ull;     // padding sorts to the very end of the tile and is never written
    }
#if PT_OS_EARLY_VALS == 2
#pragma unroll                    // variant: only pull the tile's values into L2 now (no registers held)
    for (int i = 0; i < PT_OS_KPT; ++i) {
        const uint32_t j = warp * (32 * PT_OS_KPT) + i * 32 + lane;
        if (j < n_valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(vin + base + j));
    }
#elif PT_OS_EARLY_VALS
    uint32_t val[PT_OS_KPT];      // the values are requested with the keys: loaded after the keys had left (as the three-kernel scatter does) every one of the
#pragma unroll                    // twelve shared-memory stores below waited for its own DRAM round trip (ncu: 47 % of the stall samples long-scoreboard, all on them)
    for (int i = 0; i < PT_OS_KPT; ++i) {
        const uint32_t j = warp * (32 * PT_OS_KPT) + i * 32 + lane;
        val[i] = j < n_valid ? __ldcs(vin + base + j) : 0u;
    }
#endif
#pragma unroll
    for (int i = 0; i < PT_OS_KPT; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        lp[i] = pt_rank_round(d, wh[warp], wm[warp], lane, lt);
    }
    __syncthreads();
    {   // thread `tid` owns digit `tid`
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < PT_OS_THREADS / 32; ++w) { const uint32_t c = wh[w][tid]; wh[w][tid] = acc; acc += c; }
        // padding keys (~0ull) were counted under their digit 255 (any shift): take them out of the published count
        const uint32_t pad = (tid == 255u) ? (uint32_t)PT_OS_TILE - n_valid : 0u;
        const uint32_t cnt = acc - pad;
        volatile uint32_t* st = status;
        uint32_t excl = 0;
        if (tile == 0) st[tid] = PT_OS_FLAG_INC | cnt;
        else {
            st[(size_t)tile * 256 + tid] = PT_OS_FLAG_AGG | cnt;
            // Look-back, PT_OS_LB predecessors per round trip: all tiles of a wave start together, so a tile typically has to walk over
            // every tile that is resident with it before it meets an inclusive prefix — one dependent L2 read per predecessor made the
            // first version slower than the three-kernel passes it replaces.  The loads of a batch are independent and issued together.
            bool found = false;
            for (uint32_t t = tile; t > 0 && !found;) {
                uint32_t w[PT_OS_LB];
#pragma unroll
                for (int i = 0; i < PT_OS_LB; ++i) w[i] = (uint32_t)i < t ? st[(size_t)(t - 1u - i) * 256 + tid] : PT_OS_FLAG_INC;   // before tile 0: an inclusive prefix of 0
#pragma unroll
                for (int i = 0; i < PT_OS_LB; ++i) {
                    if (found) break;
                    while ((w[i] & ~PT_OS_VALUE) == 0u) w[i] = st[(size_t)(t - 1u - i) * 256 + tid];
                    excl += w[i] & PT_OS_VALUE;
                    found = (w[i] & PT_OS_FLAG_INC) != 0u;
                }
                t = t > (uint32_t)PT_OS_LB ? t - PT_OS_LB : 0u;
            }
            st[(size_t)tile * 256 + tid] = PT_OS_FLAG_INC | (excl + cnt);
        }
        goff[tid] = digit_base[tid] + excl;
        uint32_t total;
        dstart[tid] = pt_block_excl_scan(acc, &total);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT_OS_KPT; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        lp[i] += dstart[d] + wh[warp][d];
        sbuf[lp[i]] = key[i];
    }
    __syncthreads();
    uint32_t outpos[PT_OS_KPT];
#pragma unroll
    for (int k = 0; k < PT_OS_KPT; ++k) {
        const uint32_t j = k * PT_OS_THREADS + tid;
        outpos[k] = 0xffffffffu;
        if (j < n_valid) {
            const uint64_t kk = sbuf[j];
            const uint32_t d = (uint32_t)(kk >> shift) & 255u;
            outpos[k] = goff[d] + (j - dstart[d]);
            kout[outpos[k]] = kk;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT_OS_KPT; ++i) {
        const uint32_t j = warp * (32 * PT_OS_KPT) + i * 32 + lane;
#if PT_OS_EARLY_VALS == 1
        if (j < n_valid) svals[lp[i]] = val[i];
#else
        if (j < n_valid) svals[lp[i]] = __ldcs(vin + base + j);
#endif
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PT_OS_KPT; ++k) {
        const uint32_t j = k * PT_OS_THREADS + tid;
        if (j < n_valid) vout[outpos[k]] = svals[j];
    }
}

// ---------------------------------------------------------------------------------------------------
// A3 / A4: Karras emit and bottom-up refit (second arriver proceeds)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_karras(const uint64_t* keys, PtBvh2 b) {
    for (uint32_t i = pt_gtid(); i + 1 < b.n; i += pt_gsize()) pt_karras_node(i, keys, b);
}
__device__ __forceinline__ PtBox pt_ldcg_box(const PtBox* p) {
    const float* f = reinterpret_cast<const float*>(p);
    PtBox b; b.lox = __ldcg(f); b.loy = __ldcg(f + 1); b.loz = __ldcg(f + 2); b.hix = __ldcg(f + 3); b.hiy = __ldcg(f + 4); b.hiz = __ldcg(f + 5);
    return b;
}
// A4 also runs the collapse plan (pt_plan_node) for every internal node: a node needs the boxes and the seven costs of both children.
//
// Tiled and round-synchronous.  A block owns PT_REFIT_TILE consecutive sorted leaves.  A node whose leaf range [first, last] lies
// inside the tile ("local": all but ~1/TILE of the nodes) has both subtrees inside the tile, so everything about it lives in SHARED
// memory: topology (loaded once per tile, coalesced), arrival counter, boxes, costs.  Local nodes are processed in rounds: whoever
// delivers the second child of a node appends the node to the next round's queue, and the next round hands the queued nodes to
// consecutive threads.  The dynamic programme (~300 instructions per node) therefore runs with full warps; in the classic "second
// arriver climbs on" formulation it runs with 1-4 live lanes per warp, which is what made the first version issue-bound at 3.4 ms
// for 10 M triangles (5.8 % of DRAM peak).  Results of the tile are written back coalesced at the end.
// The roots of the tile's maximal local subtrees are appended to a global list; k_refit_up continues from them with the global
// protocol (device fence, global arrival counter, sibling read back from L2): a few threads per tile, log-many levels.
#ifndef PT_REFIT_TILE
#define PT_REFIT_TILE 256
#endif
#ifndef PT_REFIT_THREADS
#define PT_REFIT_THREADS 128     // threads of a k_refit_agg block (a tile's rounds hold at most PT_REFIT_TILE / 2 nodes)
#endif
__device__ __forceinline__ PtBox pt_lds_box(const PtBox* p) {   // volatile: written by another thread of the block in an earlier round
    const volatile float* f = reinterpret_cast<const volatile float*>(p);
    PtBox b; b.lox = f[0]; b.loy = f[1]; b.loz = f[2]; b.hix = f[3]; b.hiy = f[4]; b.hiz = f[5];
    return b;
}
__device__ __forceinline__ void pt_box_union(PtBox& a, const PtBox& s) {
    a.lox = pt_min(a.lox, s.lox); a.loy = pt_min(a.loy, s.loy); a.loz = pt_min(a.loz, s.loz);
    a.hix = pt_max(a.hix, s.hix); a.hiy = pt_max(a.hiy, s.hiy); a.hiz = pt_max(a.hiz, s.hiz);
}
__global__ void __launch_bounds__(PT_REFIT_TILE) k_refit(PtBvh2 b, const PtBox* prim_box, const uint32_t* order, uint32_t* up_list, uint32_t* up_count, uint32_t max_leaf) {
    constexpr uint32_t T = PT_REFIT_TILE;
    __shared__ PtBox s_box[2 * T];                       // [0, T): internal node tile_lo + i;  [T, 2T): leaf tile_lo + i
    __shared__ float4 s_cost[T][2];                      // cost(node, 1..7) of internal node tile_lo + i
    __shared__ unsigned long long s_plan[T];
    __shared__ uint32_t s_left[T], s_right[T], s_first[T], s_last[T], s_par[2 * T];   // s_par: [0, T) internal, [T, 2T) leaves
    __shared__ uint32_t s_flag[T];                       // children delivered (2 = node processed in this tile)
    __shared__ uint32_t s_q[3][T / 2];                   // round queues (read / fill / being reset): a round never holds more than T / 2 nodes
    __shared__ uint32_t s_up[T];                         // roots of the maximal local subtrees (refs), to be continued globally
    __shared__ uint32_t s_qn[3], s_upn;
    const uint32_t n = b.n, tid = threadIdx.x;
    if (n == 1) { if (pt_gtid() == 0) b.box[0] = prim_box[order[0]]; return; }
    for (uint32_t tile_lo = blockIdx.x * T; tile_lo < n; tile_lo += gridDim.x * T) {
        const uint32_t tile_hi = min(n, tile_lo + T) - 1u;   // last leaf position of the tile
        const uint32_t j = tile_lo + tid;
        // ---- load the tile: leaf boxes (gather), topology of the internal nodes tile_lo .. tile_hi
        if (j < n) {
            const PtBox lb = prim_box[order[j]];
            s_box[T + tid] = lb;
            b.box[n - 1 + j] = lb;
            s_par[T + tid] = b.parent[n - 1 + j];
        }
        if (j < n - 1) {
            s_left[tid] = b.left[j]; s_right[tid] = b.right[j]; s_first[tid] = b.first[j]; s_last[tid] = b.last[j];
            s_par[tid] = j ? b.parent[j] : 0u;
        }
        s_flag[tid] = 0;
        if (tid == 0) { s_qn[0] = 0; s_qn[1] = 0; s_qn[2] = 0; s_upn = 0; }
        __syncthreads();
        // child `ref` is finished: deliver it to its parent p.  Second delivery queues p for the next round; a parent that is not local
        // ends the tile's part of this subtree.
        auto deliver = [&](uint32_t ref, uint32_t p, uint32_t q) {
            const uint32_t lp = p - tile_lo;     // a node's index is an end point of its leaf range, so a local node has tile_lo <= p <= tile_hi
            const bool local = p >= tile_lo && p <= tile_hi && s_first[lp] >= tile_lo && s_last[lp] <= tile_hi;
            if (!local) { s_up[atomicAdd(&s_upn, 1u)] = ref; return; }
            if (atomicAdd(&s_flag[lp], 1u) == 1u) s_q[q][atomicAdd(&s_qn[q], 1u)] = p;
        };
        if (j < n) deliver(n - 1 + j, s_par[T + tid], 0u);
        __syncthreads();
        // ---- rounds over the local nodes
        for (uint32_t cur = 0;; cur = cur == 2u ? 0u : cur + 1u) {
            const uint32_t cnt = s_qn[cur], nxt = cur == 2u ? 0u : cur + 1u;
            if (cnt == 0) break;                          // block-uniform: read after a barrier, not written before the next one
            if (tid == 0) s_qn[nxt == 2u ? 0u : nxt + 1u] = 0;   // the queue after next: read in the previous round, filled in the next one
            if (tid < cnt) {
                const uint32_t p = s_q[cur][tid], lp = p - tile_lo;
                const uint32_t L = s_left[lp], R = s_right[lp];
                float cl[7], cr[7], mc[7];
                PtBox bl, br;
                if (L >= n - 1) {
                    bl = pt_lds_box(&s_box[T + (L - (n - 1)) - tile_lo]);
                    const float lc = pt_plan_leaf_cost(bl);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cl[i] = lc;
                } else {
                    bl = pt_lds_box(&s_box[L - tile_lo]);
                    const volatile float* c = reinterpret_cast<const volatile float*>(&s_cost[L - tile_lo][0]);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cl[i] = c[i];
                }
                if (R >= n - 1) {
                    br = pt_lds_box(&s_box[T + (R - (n - 1)) - tile_lo]);
                    const float lc = pt_plan_leaf_cost(br);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cr[i] = lc;
                } else {
                    br = pt_lds_box(&s_box[R - tile_lo]);
                    const volatile float* c = reinterpret_cast<const volatile float*>(&s_cost[R - tile_lo][0]);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cr[i] = c[i];
                }
                pt_box_union(bl, br);
                s_box[lp] = bl;
                s_plan[lp] = pt_plan_node(cl, cr, pt_box_area(bl.lox, bl.loy, bl.loz, bl.hix, bl.hiy, bl.hiz), s_last[lp] - s_first[lp] + 1u, max_leaf, mc);
                s_cost[lp][0] = make_float4(mc[0], mc[1], mc[2], mc[3]); s_cost[lp][1] = make_float4(mc[4], mc[5], mc[6], 0.0f);
                if (p != 0) deliver(p, s_par[lp], nxt);
            }
            __syncthreads();
        }
        // ---- write the tile's finished nodes back (coalesced), make them visible device-wide
        if (j < n - 1 && s_flag[tid] == 2u) {
            b.box[j] = s_box[tid];
            float4* dst = reinterpret_cast<float4*>(b.cost + 8 * (size_t)j);
            dst[0] = s_cost[tid][0]; dst[1] = s_cost[tid][1];
            b.plan[j] = s_plan[tid];
        }
        // ---- the roots of the tile's maximal local subtrees are handed to k_refit_up (one global list, one atomic per tile).  Climbing
        // on right here made the whole block wait at the tile's last barrier for a handful of fence + atomic + L2 round-trip chains:
        // 73 % of all stall samples, 3.4 ms for 10 M triangles.
        __syncthreads();
        if (tid == 0) s_qn[0] = atomicAdd(up_count, s_upn);
        __syncthreads();
        if (tid < s_upn) up_list[s_qn[0] + tid] = s_up[tid];
        __syncthreads();                                   // the tile's shared state is reused by the block's next tile
    }
}
// Upper part of the refit: every root of a tile-local subtree climbs with the classic protocol (publish, device fence, global arrival
// counter; the second arriver reads the sibling back from L2 and owns the parent).  ~4 roots per tile, log-many levels.
__global__ void __launch_bounds__(128) k_refit_up(PtBvh2 b, const uint32_t* up_list, const uint32_t* up_count, uint32_t* flags, uint32_t max_leaf) {
    const uint32_t n = b.n, m = *up_count;
    for (uint32_t k = pt_gtid(); k < m; k += pt_gsize()) {
        uint32_t me = up_list[k];
        PtBox mine = b.box[me];
        float mc[7];
        if (me >= n - 1) {
            const float lc = pt_plan_leaf_cost(mine);
#pragma unroll
            for (int i = 0; i < 7; ++i) mc[i] = lc;
        } else {
            const float4 c0 = *reinterpret_cast<const float4*>(b.cost + 8 * (size_t)me), c1 = *reinterpret_cast<const float4*>(b.cost + 8 * (size_t)me + 4);
            mc[0] = c0.x; mc[1] = c0.y; mc[2] = c0.z; mc[3] = c0.w; mc[4] = c1.x; mc[5] = c1.y; mc[6] = c1.z;
        }
        for (;;) {
            const uint32_t cur = b.parent[me];
            __threadfence();                       // publish box[me] and cost[me] device-wide before announcing arrival
            if (atomicAdd(&flags[cur], 1u) == 0u) break;   // first arriver leaves; the second one owns the parent
            const uint32_t l = b.left[cur];
            const bool me_left = l == me;
            const uint32_t sib = me_left ? b.right[cur] : l;
            const PtBox s = pt_ldcg_box(&b.box[sib]);       // written by another thread of this kernel (or by k_refit): read it from L2, not L1
            float sc[7];
            if (sib >= n - 1) {
                const float lc = pt_plan_leaf_cost(s);
#pragma unroll
                for (int i = 0; i < 7; ++i) sc[i] = lc;
            } else {
                const float4 c0 = __ldcg(reinterpret_cast<const float4*>(b.cost + 8 * (size_t)sib)), c1 = __ldcg(reinterpret_cast<const float4*>(b.cost + 8 * (size_t)sib + 4));
                sc[0] = c0.x; sc[1] = c0.y; sc[2] = c0.z; sc[3] = c0.w; sc[4] = c1.x; sc[5] = c1.y; sc[6] = c1.z;
            }
            pt_box_union(mine, s);
            b.box[cur] = mine;
            float cl[7], cr[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) { cl[i] = me_left ? mc[i] : sc[i]; cr[i] = me_left ? sc[i] : mc[i]; }
            const uint32_t cnt = b.last[cur] - b.first[cur] + 1u;
            b.plan[cur] = pt_plan_node(cl, cr, pt_box_area(mine.lox, mine.loy, mine.loz, mine.hix, mine.hiy, mine.hiz), cnt, max_leaf, mc);
            float4* dst = reinterpret_cast<float4*>(b.cost + 8 * (size_t)cur);
            dst[0] = make_float4(mc[0], mc[1], mc[2], mc[3]); dst[1] = make_float4(mc[4], mc[5], mc[6], 0.0f);
            if (cur == 0) break;
            me = cur;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// A3 + A4 fused (PT_AGGLOMERATIVE): the radix tree is built bottom-up WHILE the refit climbs (Apetrei, "Fast and Simple Agglomerative
// LBVH Construction", 2014).  A finished subtree over sorted leaves [l, r] joins the neighbour it shares the longer key prefix with:
// delta(r, r+1) > delta(l-1, l)  ->  it is the LEFT child of node r, else the RIGHT child of node l-1 (a node's id is the position of
// its split).  Same tree as Karras' top-down emit (the deltas are all distinct thanks to the index tie-break), different node ids —
// nothing downstream depends on the ids — and no binary searches: k_karras (divergence-bound, 0.43 ms for 10 M keys) and the
// parent / left / right / first / last round trip through HBM disappear.  Tile-local test without knowing the sibling: the parent of a
// left child [l, r] is local iff the right sibling, which starts at r+1 and ends before the first key that does NOT share more than
// delta(r, r+1) bits with key r+1, ends inside the tile: one more delta against the key just past the tile (mirrored for right children).
// Both children evaluate the same predicate, so they meet either in shared memory or through the global protocol, never one in each.
__global__ void __launch_bounds__(PT_REFIT_THREADS) k_refit_agg(PtBvh2 b, const uint64_t* keys, const PtBox* prim_box, const uint32_t* order, uint32_t* up_list,
                                                            uint32_t* up_count, uint32_t* root_out, uint32_t max_leaf) {
    constexpr uint32_t T = PT_REFIT_TILE;
    __shared__ PtBox s_box[2 * T];
    __shared__ float4 s_cost[T][2];
    __shared__ unsigned long long s_plan[T];
    __shared__ uint64_t s_key[T + 2];                    // keys of positions tile_lo - 1 .. tile_hi + 1
    __shared__ uint32_t s_left[T], s_right[T], s_first[T], s_last[T];
    __shared__ uint32_t s_flag[T];
    __shared__ uint32_t s_q[3][T / 2];
    __shared__ uint32_t s_up[T];
    __shared__ uint32_t s_qn[3], s_upn;
    const uint32_t n = b.n, tid = threadIdx.x;
    if (n == 1) { if (pt_gtid() == 0) b.box[0] = prim_box[order[0]]; return; }
    for (uint32_t tile_lo = blockIdx.x * T; tile_lo < n; tile_lo += gridDim.x * T) {
        const uint32_t tile_hi = min(n, tile_lo + T) - 1u;
        const uint32_t koff = tile_lo - 1u;              // s_key[i - koff]; position tile_lo - 1 only exists when tile_lo > 0 (never read otherwise)
        // PT_REFIT_THREADS threads serve a tile of T leaves: a round never holds more than T / 2 nodes, so T / 2 threads are enough for the rounds and
        // half as many warps wait at every round barrier; the per-leaf phases (load, write back) take T / PT_REFIT_THREADS passes
        for (uint32_t t = tid; t < T; t += PT_REFIT_THREADS) {
            const uint32_t j = tile_lo + t;
            if (j < n) {
                const PtBox lb = prim_box[order[j]];
                s_box[T + t] = lb;
                b.box[n - 1 + j] = lb;
                s_key[t + 1] = keys[j];
            }
            s_flag[t] = 0;
        }
        if (tid == 0 && tile_lo > 0) s_key[0] = keys[tile_lo - 1];
        if (tid == 0 && tile_hi + 1 < n) s_key[tile_hi + 2 - tile_lo] = keys[tile_hi + 1];
        if (tid == 0) { s_qn[0] = 0; s_qn[1] = 0; s_qn[2] = 0; s_upn = 0; }
        __syncthreads();
        // subtree `ref` over [l, r] (inside the tile) is finished: find its parent, hand it over
        auto deliver = [&](uint32_t ref, uint32_t l, uint32_t r, uint32_t q) {
            if (l == 0 && r == n - 1) { *root_out = ref; return; }                      // the whole tree fits one tile
            const PtJoin jn = pt_join(s_key, koff, n, l, r);
            const bool local = pt_join_is_local(s_key, koff, n, l, r, jn, tile_lo, tile_hi);
            if (!local) { s_up[atomicAdd(&s_upn, 1u)] = ref; return; }
            const uint32_t lp = jn.p - tile_lo;
            if (jn.left) { s_left[lp] = ref; s_first[lp] = l; } else { s_right[lp] = ref; s_last[lp] = r; }
            __threadfence_block();
            if (atomicAdd(&s_flag[lp], 1u) == 1u) s_q[q][atomicAdd(&s_qn[q], 1u)] = jn.p;
        };
        for (uint32_t t = tid; t < T; t += PT_REFIT_THREADS) { const uint32_t j = tile_lo + t; if (j < n) deliver(n - 1 + j, j, j, 0u); }
        __syncthreads();
        for (uint32_t cur = 0;; cur = cur == 2u ? 0u : cur + 1u) {
            const uint32_t cnt = s_qn[cur], nxt = cur == 2u ? 0u : cur + 1u;
            if (cnt == 0) break;
            if (tid == 0) s_qn[nxt == 2u ? 0u : nxt + 1u] = 0;
            for (uint32_t qi = tid; qi < cnt; qi += PT_REFIT_THREADS) {
                const uint32_t p = s_q[cur][qi], lp = p - tile_lo;
                const uint32_t L = *const_cast<volatile uint32_t*>(&s_left[lp]), R = *const_cast<volatile uint32_t*>(&s_right[lp]);
                const uint32_t f = *const_cast<volatile uint32_t*>(&s_first[lp]), la = *const_cast<volatile uint32_t*>(&s_last[lp]);
                float cl[7], cr[7], mc[7];
                PtBox bl, br;
                if (L >= n - 1) {
                    bl = pt_lds_box(&s_box[T + (L - (n - 1)) - tile_lo]);
                    const float lc = pt_plan_leaf_cost(bl);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cl[i] = lc;
                } else {
                    bl = pt_lds_box(&s_box[L - tile_lo]);
                    const volatile float* c = reinterpret_cast<const volatile float*>(&s_cost[L - tile_lo][0]);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cl[i] = c[i];
                }
                if (R >= n - 1) {
                    br = pt_lds_box(&s_box[T + (R - (n - 1)) - tile_lo]);
                    const float lc = pt_plan_leaf_cost(br);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cr[i] = lc;
                } else {
                    br = pt_lds_box(&s_box[R - tile_lo]);
                    const volatile float* c = reinterpret_cast<const volatile float*>(&s_cost[R - tile_lo][0]);
#pragma unroll
                    for (int i = 0; i < 7; ++i) cr[i] = c[i];
                }
                pt_box_union(bl, br);
                s_box[lp] = bl;
                s_plan[lp] = pt_plan_node(cl, cr, pt_box_area(bl.lox, bl.loy, bl.loz, bl.hix, bl.hiy, bl.hiz), la - f + 1u, max_leaf, mc);
                s_cost[lp][0] = make_float4(mc[0], mc[1], mc[2], mc[3]); s_cost[lp][1] = make_float4(mc[4], mc[5], mc[6], 0.0f);
                deliver(p, f, la, nxt);
            }
            __syncthreads();
        }
        // write back the nodes finished in this tile: boxes, costs, plan AND their topology (the collapse reads left / right / first / last)
        for (uint32_t t = tid; t < T; t += PT_REFIT_THREADS) {
            const uint32_t j = tile_lo + t;
            if (j < n - 1 && s_flag[t] == 2u) {
                b.box[j] = s_box[t];
                float4* dst = reinterpret_cast<float4*>(b.cost + 8 * (size_t)j);
                dst[0] = s_cost[t][0]; dst[1] = s_cost[t][1];
                b.plan[j] = s_plan[t];
                b.left[j] = s_left[t]; b.right[j] = s_right[t]; b.first[j] = s_first[t]; b.last[j] = s_last[t];
            }
        }
        __syncthreads();
        if (tid == 0) s_qn[0] = atomicAdd(up_count, s_upn);
        __syncthreads();
        for (uint32_t t = tid; t < s_upn; t += PT_REFIT_THREADS) up_list[s_qn[0] + t] = s_up[t];
        __syncthreads();
    }
}
// Upper levels: the roots of the tile-local subtrees keep joining neighbours through global memory.  The first child to arrive leaves
// its link and its end of the range in the parent's record; the second one reads them back (L2), owns the parent and climbs on.
__global__ void __launch_bounds__(128) k_refit_agg_up(PtBvh2 b, const uint64_t* keys, const uint32_t* up_list, const uint32_t* up_count, uint32_t* flags,
                                                      uint32_t* root_out, uint32_t max_leaf) {
    const uint32_t n = b.n, m = *up_count;
    for (uint32_t k = pt_gtid(); k < m; k += pt_gsize()) {
        uint32_t me = up_list[k];
        uint32_t l, r;
        PtBox mine = b.box[me];
        float mc[7];
        if (me >= n - 1) {
            l = r = me - (n - 1);
            const float lc = pt_plan_leaf_cost(mine);
#pragma unroll
            for (int i = 0; i < 7; ++i) mc[i] = lc;
        } else {
            l = b.first[me]; r = b.last[me];
            const float4 c0 = *reinterpret_cast<const float4*>(b.cost + 8 * (size_t)me), c1 = *reinterpret_cast<const float4*>(b.cost + 8 * (size_t)me + 4);
            mc[0] = c0.x; mc[1] = c0.y; mc[2] = c0.z; mc[3] = c0.w; mc[4] = c1.x; mc[5] = c1.y; mc[6] = c1.z;
        }
        for (;;) {
            if (l == 0 && r == n - 1) { *root_out = me; break; }
            const PtJoin jn = pt_join(keys, 0u, n, l, r);
            const uint32_t cur = jn.p;
            if (jn.left) { b.left[cur] = me; b.first[cur] = l; } else { b.right[cur] = me; b.last[cur] = r; }
            __threadfence();                       // publish box / cost of `me` and its link in `cur` before announcing arrival
            if (atomicAdd(&flags[cur], 1u) == 0u) break;
            uint32_t sib;
            if (jn.left) { sib = __ldcg(&b.right[cur]); r = __ldcg(&b.last[cur]); } else { sib = __ldcg(&b.left[cur]); l = __ldcg(&b.first[cur]); }
            const PtBox s = pt_ldcg_box(&b.box[sib]);
            float sc[7];
            if (sib >= n - 1) {
                const float lc = pt_plan_leaf_cost(s);
#pragma unroll
                for (int i = 0; i < 7; ++i) sc[i] = lc;
            } else {
                const float4 c0 = __ldcg(reinterpret_cast<const float4*>(b.cost + 8 * (size_t)sib)), c1 = __ldcg(reinterpret_cast<const float4*>(b.cost + 8 * (size_t)sib + 4));
                sc[0] = c0.x; sc[1] = c0.y; sc[2] = c0.z; sc[3] = c0.w; sc[4] = c1.x; sc[5] = c1.y; sc[6] = c1.z;
            }
            pt_box_union(mine, s);
            b.box[cur] = mine;
            float cl[7], cr[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) { cl[i] = jn.left ? mc[i] : sc[i]; cr[i] = jn.left ? sc[i] : mc[i]; }
            b.plan[cur] = pt_plan_node(cl, cr, pt_box_area(mine.lox, mine.loy, mine.loz, mine.hix, mine.hiy, mine.hiz), r - l + 1u, max_leaf, mc);
            float4* dst = reinterpret_cast<float4*>(b.cost + 8 * (size_t)cur);
            dst[0] = make_float4(mc[0], mc[1], mc[2], mc[3]); dst[1] = make_float4(mc[4], mc[5], mc[6], 0.0f);
            me = cur;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// A5: level-synchronous BVH2 -> BVH8 collapse
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_collapse_select(PtBvh2 b, const uint32_t* refs, uint32_t m, uint32_t max_leaf, uint32_t* slots, uint32_t* n_int,
                                                         uint32_t* n_prim) {
    for (uint32_t w = pt_gtid(); w < m; w += pt_gsize()) {
        uint32_t s[8], ni, np;
        pt_collapse_select(b, refs[w], max_leaf, s, &ni, &np);
#pragma unroll
        for (int k = 0; k < 8; ++k) slots[8 * (size_t)w + k] = s[k];
        n_int[w] = ni; n_prim[w] = np;
    }
}
__global__ void __launch_bounds__(128) k_collapse_emit(PtBvh2 b, const uint32_t* refs, uint32_t m, uint32_t max_leaf, const PtBuildParams* bp,
                                                       const uint32_t* slots, const uint32_t* off_int, const uint32_t* off_prim, uint32_t level_start,
                                                       uint32_t next_start, uint32_t prim_total, PtNode8* nodes, uint32_t* next_refs, uint32_t* leaf_seq) {
    const float pad = bp->pad;
    for (uint32_t w = pt_gtid(); w < m; w += pt_gsize()) {
        uint32_t s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = slots[8 * (size_t)w + k];
        pt_collapse_emit(b, refs[w], s, max_leaf, pad, next_start + off_int[w], prim_total + off_prim[w], &nodes[level_start + w], next_refs + off_int[w],
                         leaf_seq);
    }
}
// The first levels of the collapse hold 1, ~6, ~40, ... nodes: as separate launches each level is a select kernel, two 3-kernel scans, an
// emit kernel and a host round trip for the level's totals — ~100 us of pure latency per level.  One block walks all levels of at most
// PT_TOP_NODES nodes instead (one thread per wide node, block-wide scans, the same numbering as the level-synchronous path) and leaves
// {m, level_start, prim_total, which refs buffer is current} for the host to carry on from.
#define PT_TOP_NODES 256
__global__ void __launch_bounds__(PT_TOP_NODES) k_collapse_top(PtBvh2 b, uint32_t* refs_a, uint32_t* refs_b, uint32_t max_leaf, const PtBuildParams* bp, PtNode8* nodes,
                                                               uint32_t* leaf_seq, uint32_t* state) {
    const float pad = bp->pad;
    const uint32_t tid = threadIdx.x;
    uint32_t m = 1, level_start = 0, prim_total = 0, parity = 0;
    uint32_t *cur = refs_a, *nxt = refs_b;
    while (m > 0 && m <= PT_TOP_NODES && level_start + m <= b.n) {
        uint32_t s[8], ni = 0, np = 0;
        if (tid < m) pt_collapse_select(b, cur[tid], max_leaf, s, &ni, &np);
        uint32_t tot_i, tot_p;
        const uint32_t off_i = pt_block_excl_scan(ni, &tot_i);
        const uint32_t off_p = pt_block_excl_scan(np, &tot_p);
        if (tid < m) pt_collapse_emit(b, cur[tid], s, max_leaf, pad, level_start + m + off_i, prim_total + off_p, &nodes[level_start + tid], nxt + off_i, leaf_seq);
        __syncthreads();                                  // next level's refs (global memory) are visible to the whole block
        level_start += m; prim_total += tot_p; m = tot_i; parity ^= 1u;
        uint32_t* t = cur; cur = nxt; nxt = t;
    }
    if (tid == 0) { state[0] = m; state[1] = level_start; state[2] = prim_total; state[3] = parity; }
}
// ---------------------------------------------------------------------------------------------------
// A5, all remaining levels in ONE persistent kernel (round 1: per level a select kernel, two scan kernels, an emit kernel and a host
// round trip for the level's totals).  The grid is one resident wave (cooperative launch); levels are separated by a grid barrier.
//   phase 1  the collapse plan is walked one lane per node (32 nodes per warp), then every wide node is built by EIGHT lanes (one per child /
//            per slot), four nodes of the warp at a time: each lane loads one child box, the greedy octant assignment runs as 8 rounds of an 8-lane arg-max (the
//            thread-per-node form spent ~440 warp instructions per node on it: divergent loops over local-memory arrays), each lane
//            quantises the child of its slot, and the 80-byte node — everything but child_base / tri_base, which need the level's
//            prefix sums — leaves through shared memory as coalesced words.  Per node it also records the slot refs and the two counts.
//   barrier, then every block sums the per-block counts before it: no scan kernel, no host.
//   phase 2  one thread per node: block scan of the counts inside the block's contiguous segment, patch the two bases, write the next
//            level's refs and the leaf sequence.
// Same numbering as the level-synchronous path (nodes of a level in ref order, children contiguous in slot order).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pt_grid_barrier(uint32_t* counter, uint32_t& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();                                   // cumulative: the block's writes (ordered before by the barrier above) become visible device-wide
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile uint32_t*>(counter) < epoch) {}
        __threadfence();                                   // acquire side: also drops stale L1 lines of this SM
    }
    __syncthreads();
}
struct PtCollapseArgs {
    PtBvh2 b; uint32_t* refs_a; uint32_t* refs_b; uint32_t max_leaf; const PtBuildParams* bp; PtNode8* nodes; uint32_t* leaf_seq;
    uint32_t* state;        // in: {m, level_start, prim_total, parity} from k_collapse_top; out: {0, num_nodes, prim_total, error}
    uint32_t* slots; uint32_t* n_int; uint32_t* n_prim; uint32_t cap;   // per-level scratch, cap entries
    uint32_t* block_sums;   // 2 x gridDim
    uint32_t* barrier;      // zero at launch
    uint32_t node_cap;      // capacity of `nodes`
};
#ifndef PT_COLLAPSE_TIMING
#define PT_COLLAPSE_TIMING 0     // debug: block 0 prints the wall time of every level's phases (globaltimer)
#endif
#ifndef PT_COLLAPSE_REDUX
#define PT_COLLAPSE_REDUX 1
#endif
#define PT_CL_THREADS 128
__global__ void __launch_bounds__(PT_CL_THREADS) k_collapse_levels(PtCollapseArgs a) {
    __shared__ uint32_t s_C[PT_CL_THREADS][8];
    __shared__ uint32_t s_nc[PT_CL_THREADS], s_ref[PT_CL_THREADS];
    __shared__ __align__(16) uint32_t s_node[PT_CL_THREADS / 8][20];
    __shared__ uint32_t s_red[4][PT_CL_THREADS / 32];
    const PtBvh2& b = a.b;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, grp = tid >> 3, k = tid & 7u, gl = lane & ~7u /* first lane of my group */;
    const float pad = a.bp->pad;
    uint32_t m = a.state[0], level_start = a.state[1], prim_total = a.state[2];
    uint32_t* cur = a.state[3] ? a.refs_b : a.refs_a;
    uint32_t* nxt = a.state[3] ? a.refs_a : a.refs_b;
    uint32_t epoch = 0, error = 0;
#if PT_COLLAPSE_TIMING
    unsigned long long t_lvl[16][4]; uint32_t m_lvl[16]; int n_lvl = 0;
    auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
#endif
    while (m > 0) {
        if (m > a.cap || (uint64_t)level_start + m > a.node_cap) { error = 1; break; }      // grid-uniform
#if PT_COLLAPSE_TIMING
        if (n_lvl < 16) { t_lvl[n_lvl][0] = now(); m_lvl[n_lvl] = m; }
#endif
        const uint32_t seg = (m + gridDim.x - 1) / gridDim.x;
        const uint32_t lo = min(m, blockIdx.x * seg), hi = min(m, lo + seg);
        uint32_t sum_i = 0, sum_p = 0;
        // ---------------- phase 1
        for (uint32_t wb = lo; wb < hi; wb += PT_CL_THREADS) {          // block-uniform trip count; a warp owns 32 consecutive nodes of the step
            {   // (A) the plan walk, ONE LANE PER NODE: as "lane 0 of every 8-lane group" it ran with 4 lanes per warp and was 24 % of all issued
                // instructions of the kernel (ncu source page); the children lists go to shared memory for the cooperative part
                const uint32_t wn = wb + tid;
                uint32_t C[8]; int n_c = 0; uint32_t r0 = 0;
                if (wn < hi) {
                    r0 = __ldcg(cur + wn);
                    if (pt_b2_count(b, r0) <= a.max_leaf) { C[0] = r0; n_c = 1; } else n_c = pt_plan_children(b, r0, C);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) s_C[tid][i] = i < n_c ? C[i] : PT_NONE;
                s_nc[tid] = (uint32_t)n_c; s_ref[tid] = r0;
            }
            __syncwarp();
          for (uint32_t sub = 0; sub < 8u; ++sub) {                      // (B) 4 nodes of the warp at a time, 8 lanes each
            const uint32_t nidx = warp * 32u + sub * 4u + (lane >> 3);   // which of the block's 128 nodes my group builds now
            const uint32_t w0 = wb + warp * 32u + sub * 4u - warp * 4u;  // so that w0 + warp * 4 is the warp's first node of this sub-step (write-out below)
            const uint32_t w = wb + nidx;
            const bool node_ok = w < hi;
            uint32_t ref = 0, nc = 0;
            if (node_ok) { ref = s_ref[nidx]; nc = s_nc[nidx]; }
            const bool child_ok = node_ok && k < nc;
            const uint32_t cref = child_ok ? s_C[nidx][k] : PT_NONE;
            PtBox nb, cb;
            nb.lox = nb.loy = nb.loz = nb.hix = nb.hiy = nb.hiz = 0.0f; cb = nb;
            uint32_t ccnt = 0, cfirst = 0;
            if (node_ok) nb = b.box[ref];
            if (child_ok) {
                cb = b.box[cref];
                // an internal ref covers >= 2 primitives: with one triangle per leaf slot its count / first position (two more loads) are never needed
                const bool leaf_ref = cref >= b.n - 1;
                ccnt = leaf_ref ? 1u : (a.max_leaf > 1u ? pt_b2_count(b, cref) : 2u);
                cfirst = leaf_ref ? cref - (b.n - 1) : (a.max_leaf > 1u ? b.first[cref] : 0u);
            }
            // costs of putting this child into slot s: (+-dx +- dy) +- dz, the same operation order as pt_collapse_select
            const float cx = (nb.lox + nb.hix) * 0.5f, cy = (nb.loy + nb.hiy) * 0.5f, cz = (nb.loz + nb.hiz) * 0.5f;
            const float dx = (cb.lox + cb.hix) * 0.5f - cx, dy = (cb.loy + cb.hiy) * 0.5f - cy, dz = (cb.loz + cb.hiz) * 0.5f - cz;
            float c[8];
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) c[sl] = (((sl & 4) ? dx : -dx) + ((sl & 2) ? dy : -dy)) + ((sl & 1) ? dz : -dz);
            // greedy assignment: nc rounds; in each the best (child, free slot) pair of the group wins, ties to the smallest (child, slot)
            int kslot = -1;                                                      // kslot: child index that landed in slot `k` (this lane as a SLOT)
            const uint32_t rounds = __reduce_max_sync(PT_FULL, nc);
#if PT_COLLAPSE_REDUX
            // Same greedy rule, a third of the instructions (this loop was 49 % of everything the kernel issued, ncu source page): a lane keeps the
            // costs of its still-free slots (taken ones become -inf), its proposal is their maximum (FMNMX3 tree; ties -> smallest slot), and the
            // group's winner comes from two integer butterflies over the 8 lanes: the maximum of the order-preserving integer image of the proposals,
            // then the smallest (child, slot) key among the lanes that hold it.  -0 is folded into +0 first (x + 0.0f), so integer equality is float
            // equality.  (redux.sync with one member mask per 8-lane group is NOT the tool: with four different masks in a warp the compiler emits a
            // loop over the distinct masks, BRA.DIV + CREDUX per group, and nothing is gained — measured.)
            const float ninf = __uint_as_float(0xff800000u);
            bool done = !child_ok;
            float cm[8];
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) cm[sl] = c[sl];
            for (uint32_t it = 0; it < rounds; ++it) {
                const float m = fmaxf(fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])), fmaxf(fmaxf(cm[4], cm[5]), fmaxf(cm[6], cm[7])));
                uint32_t bs = 7u;
#pragma unroll
                for (int sl = 6; sl >= 0; --sl) bs = cm[sl] == m ? (uint32_t)sl : bs;            // smallest free slot that attains the maximum
                const uint32_t mb = __float_as_uint(m + 0.0f);
                const uint32_t u = done ? 0u : ((mb & 0x80000000u) ? ~mb : (mb | 0x80000000u));   // 0: no proposal (every finite float maps above it)
                uint32_t umax = u;
#pragma unroll
                for (int o = 4; o >= 1; o >>= 1) umax = max(umax, __shfl_xor_sync(PT_FULL, umax, o));
                uint32_t win = (!done && u == umax) ? (k << 3 | bs) : 0xffu;                // both butterflies run warp-wide: a group without proposals
#pragma unroll                                                                               // (umax = 0) must not leave the shuffles to the others
                for (int o = 4; o >= 1; o >>= 1) win = min(win, __shfl_xor_sync(PT_FULL, win, o));
                if (win != 0xffu) {                                                         // group-uniform
                    const uint32_t wk = win >> 3, ws = win & 7u;
                    if (wk == k) done = true;
                    if (ws == k) kslot = (int)wk;
#pragma unroll
                    for (int sl = 0; sl < 8; ++sl) cm[sl] = ws == (uint32_t)sl ? ninf : cm[sl];
                }
            }
#else
            uint32_t slot_used = 0; bool done = !child_ok;
            for (uint32_t it = 0; it < rounds; ++it) {
                int bs = -1; float bv = 0.0f;
#pragma unroll
                for (int sl = 0; sl < 8; ++sl)
                    if (!done && !((slot_used >> sl) & 1u) && (bs < 0 || c[sl] > bv)) { bv = c[sl]; bs = sl; }
                uint32_t key = bs >= 0 ? (k << 3 | (uint32_t)bs) : 0xffu;                // 0xff: no candidate
#pragma unroll
                for (int o = 4; o >= 1; o >>= 1) {
                    const float ov = __shfl_xor_sync(PT_FULL, bv, o);
                    const uint32_t ok = __shfl_xor_sync(PT_FULL, key, o);
                    const bool take = ok != 0xffu && (key == 0xffu || ov > bv || (ov == bv && ok < key));
                    if (take) { bv = ov; key = ok; }
                }
                if (key != 0xffu) {                                                     // group-uniform
                    const uint32_t wk = key >> 3, ws = key & 7u;
                    slot_used |= 1u << ws;
                    if (wk == k) done = true;
                    if (ws == k) kslot = (int)wk;
                }
            }
#endif
            // the child of my slot: fetch its data from the lane that owns it
            const int src = (int)gl + (kslot >= 0 ? kslot : 0);
            const uint32_t sref = __shfl_sync(PT_FULL, cref, src), scnt = __shfl_sync(PT_FULL, ccnt, src), sfirst = __shfl_sync(PT_FULL, cfirst, src);
            PtBox sb;
            sb.lox = __shfl_sync(PT_FULL, cb.lox, src); sb.loy = __shfl_sync(PT_FULL, cb.loy, src); sb.loz = __shfl_sync(PT_FULL, cb.loz, src);
            sb.hix = __shfl_sync(PT_FULL, cb.hix, src); sb.hiy = __shfl_sync(PT_FULL, cb.hiy, src); sb.hiz = __shfl_sync(PT_FULL, cb.hiz, src);
            const bool slot_ok = node_ok && kslot >= 0;
            const bool is_leaf = slot_ok && scnt <= a.max_leaf;
            const bool is_int = slot_ok && !is_leaf;
            const uint32_t imask = (__ballot_sync(PT_FULL, is_int) >> gl) & 0xffu;
            // triangle offset of my leaf inside the node's block: exclusive prefix of the leaf counts over the lower slots
            uint32_t incl = is_leaf ? scnt : 0u;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { const uint32_t t = __shfl_up_sync(PT_FULL, incl, o, 8); if (k >= (uint32_t)o) incl += t; }
            const uint32_t off = incl - (is_leaf ? scnt : 0u);
            const uint32_t nprim = __shfl_sync(PT_FULL, incl, (int)gl + 7);
            // quantisation frame of the node (pt_collapse_emit's arithmetic)
            float p[3] = {nb.lox - pad, nb.loy - pad, nb.loz - pad};
            float hi3[3] = {nb.hix + pad, nb.hiy + pad, nb.hiz + pad};
            float inv[3]; uint32_t e[3];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) { e[ax] = pt_quant_exp(hi3[ax] - p[ax]); inv[ax] = pt_u2f((254u - e[ax]) << 23); }
            uint32_t q[6] = {255u, 255u, 255u, 0u, 0u, 0u}, meta = 0u;
            if (slot_ok) {
                q[0] = pt_quant_lo(sb.lox - pad, p[0], inv[0]); q[3] = pt_quant_hi(sb.hix + pad, p[0], inv[0]);
                q[1] = pt_quant_lo(sb.loy - pad, p[1], inv[1]); q[4] = pt_quant_hi(sb.hiy + pad, p[1], inv[1]);
                q[2] = pt_quant_lo(sb.loz - pad, p[2], inv[2]); q[5] = pt_quant_hi(sb.hiz + pad, p[2], inv[2]);
                meta = is_leaf ? ((((1u << scnt) - 1u) << 5) | off) : (0x20u | (24u + k));
            }
            if (node_ok) {
                uint8_t* img = reinterpret_cast<uint8_t*>(&s_node[grp][0]);
                img[24 + k] = (uint8_t)meta;
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) img[32 + 8 * pl + k] = (uint8_t)q[pl];
                if (k == 0) {
                    s_node[grp][0] = pt_f2u(p[0]); s_node[grp][1] = pt_f2u(p[1]); s_node[grp][2] = pt_f2u(p[2]);
                    s_node[grp][3] = e[0] | (e[1] << 8) | (e[2] << 16) | (imask << 24);
                    s_node[grp][4] = 0u; s_node[grp][5] = 0u;                             // child_base, tri_base: phase 2
                    a.n_int[w] = (uint32_t)__popc(imask); a.n_prim[w] = nprim;
                    sum_i += (uint32_t)__popc(imask); sum_p += nprim;
                }
                a.slots[8 * (size_t)w + k] = slot_ok ? sref : PT_NONE;
                // first sorted position of my leaf's primitives rides along in the high part of the scratch: phase 2 needs it without touching first[] again
            }
            __syncwarp();
            {   // the warp's (up to) four nodes leave as consecutive words
                const uint32_t wfirst = w0 + warp * 4u;
                const uint32_t nn = wfirst < hi ? min(4u, hi - wfirst) : 0u;
                uint32_t* dst = reinterpret_cast<uint32_t*>(a.nodes + level_start + wfirst);
                const uint32_t* srcw = &s_node[warp * 4u][0];
                for (uint32_t i = lane; i < nn * 20u; i += 32u) dst[i] = srcw[i];
            }
            __syncwarp();
          }
        }
        // per-block counts of this level
        sum_i = __reduce_add_sync(PT_FULL, sum_i); sum_p = __reduce_add_sync(PT_FULL, sum_p);
        if (lane == 0) { s_red[0][warp] = sum_i; s_red[1][warp] = sum_p; }
        __syncthreads();
        if (tid == 0) {
            uint32_t si = 0, sp = 0;
            for (int i = 0; i < PT_CL_THREADS / 32; ++i) { si += s_red[0][i]; sp += s_red[1][i]; }
            a.block_sums[2 * blockIdx.x] = si; a.block_sums[2 * blockIdx.x + 1] = sp;
        }
#if PT_COLLAPSE_TIMING
        if (n_lvl < 16) t_lvl[n_lvl][1] = now();
#endif
        pt_grid_barrier(a.barrier, epoch);
#if PT_COLLAPSE_TIMING
        if (n_lvl < 16) t_lvl[n_lvl][2] = now();
#endif
        // ---------------- offsets of this block's segment + the level's totals
        uint32_t pre_i = 0, pre_p = 0, tot_i = 0, tot_p = 0;
        for (uint32_t bb = tid; bb < gridDim.x; bb += PT_CL_THREADS) {
            const uint32_t vi = __ldcg(a.block_sums + 2 * bb), vp = __ldcg(a.block_sums + 2 * bb + 1);
            tot_i += vi; tot_p += vp;
            if (bb < blockIdx.x) { pre_i += vi; pre_p += vp; }
        }
        pre_i = __reduce_add_sync(PT_FULL, pre_i); pre_p = __reduce_add_sync(PT_FULL, pre_p);
        tot_i = __reduce_add_sync(PT_FULL, tot_i); tot_p = __reduce_add_sync(PT_FULL, tot_p);
        if (lane == 0) { s_red[0][warp] = pre_i; s_red[1][warp] = pre_p; s_red[2][warp] = tot_i; s_red[3][warp] = tot_p; }
        __syncthreads();
        pre_i = pre_p = tot_i = tot_p = 0;
        for (int i = 0; i < PT_CL_THREADS / 32; ++i) { pre_i += s_red[0][i]; pre_p += s_red[1][i]; tot_i += s_red[2][i]; tot_p += s_red[3][i]; }
        __syncthreads();
        // ---------------- phase 2
        const uint32_t next_start = level_start + m;
        for (uint32_t w0 = lo; w0 < hi; w0 += PT_CL_THREADS) {
            const uint32_t w = w0 + tid;
            const bool ok = w < hi;
            const uint32_t ni = ok ? a.n_int[w] : 0u, np = ok ? a.n_prim[w] : 0u;
            uint32_t ti, tp;
            const uint32_t oi = pre_i + pt_block_excl_scan(ni, &ti);
            const uint32_t op = pre_p + pt_block_excl_scan(np, &tp);
            pre_i += ti; pre_p += tp;
            if (ok) {
                const uint32_t child_base = next_start + oi, prim_base = prim_total + op;
                *reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(a.nodes + level_start + w) + 4) = make_uint2(child_base, prim_base);
                uint32_t offp = 0, nint = 0;
#pragma unroll
                for (int sl = 0; sl < 8; ++sl) {
                    const uint32_t r = a.slots[8 * (size_t)w + sl];
                    if (r == PT_NONE) continue;
                    const uint32_t cnt = r >= b.n - 1 ? 1u : (a.max_leaf > 1u ? pt_b2_count(b, r) : 2u);   // an internal ref covers >= 2 primitives
                    if (cnt <= a.max_leaf) {
                        const uint32_t fp = pt_b2_lopos(b, r);
                        for (uint32_t qq = 0; qq < cnt; ++qq) a.leaf_seq[prim_base + offp + qq] = fp + qq;
                        offp += cnt;
                    } else nxt[oi + nint++] = r;
                }
            }
        }
#if PT_COLLAPSE_TIMING
        if (n_lvl < 16) t_lvl[n_lvl][3] = now();
#endif
        pt_grid_barrier(a.barrier, epoch);
#if PT_COLLAPSE_TIMING
        ++n_lvl;
#endif
        level_start = next_start; prim_total += tot_p; m = tot_i;
        uint32_t* t = cur; cur = nxt; nxt = t;
    }
#if PT_COLLAPSE_TIMING
    if ((blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && tid == 0 && n_lvl > 6) {
        const unsigned long long t_end = now();
        for (int i = 0; i < n_lvl && i < 16; ++i)
            printf("blk %u level %d m %u: phase1 %llu ns, barrier wait %llu ns, phase2 %llu ns, to next level %llu ns\n", blockIdx.x, i, m_lvl[i], t_lvl[i][1] - t_lvl[i][0],
                   t_lvl[i][2] - t_lvl[i][1], t_lvl[i][3] - t_lvl[i][2], (i + 1 < n_lvl ? t_lvl[i + 1][0] : t_end) - t_lvl[i][3]);
    }
#endif
    if (blockIdx.x == 0 && tid == 0) { a.state[0] = m; a.state[1] = level_start; a.state[2] = prim_total; a.state[3] = error; }
}

__global__ void __launch_bounds__(256) k_write_tris(PtMeshRaw m, const uint32_t* order, const uint32_t* leaf_seq, PtTri* tris) {
    for (uint32_t k = pt_gtid(); k < m.ntris; k += pt_gsize()) {
        uint32_t i = order[leaf_seq[k]];
        pt_v3 a, b, c;
        pt_load_tri(m, i, &a, &b, &c);
        PtTri t;
        t.v0x = a.x; t.v0y = a.y; t.v0z = a.z; t.prim = i;
        t.e1x = b.x - a.x; t.e1y = b.y - a.y; t.e1z = b.z - a.z; t.mat = m.mat ? m.mat[i] : 0u;
        t.e2x = c.x - a.x; t.e2y = c.y - a.y; t.e2z = c.z - a.z; t.pad = 0u;
        uint4* dst = reinterpret_cast<uint4*>(tris + k);
        const uint4* src = reinterpret_cast<const uint4*>(&t);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}

// Emissive-triangle list of a mesh (input for the light table): stream the material ids once on the device instead of
// walking 10 M triangles on the host.  Order is restored by a host sort of the (tiny) result.
__global__ void __launch_bounds__(256) k_emissive_list(const uint32_t* mat, uint32_t n, const uint8_t* mat_emissive, uint32_t num_mats, uint32_t* out, uint32_t* count) {
    const uint32_t rounds = (n + pt_gsize() - 1) / pt_gsize();
    for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t i = r * pt_gsize() + pt_gtid();
        bool em = false;
        if (i < n) { uint32_t m = mat[i]; em = mat_emissive[m < num_mats ? m : 0] != 0; }
        uint32_t mask = __ballot_sync(PT_FULL, em);
        if (mask) {
            uint32_t base = 0;
            if (pt_lane() == 0) base = atomicAdd(count, (uint32_t)__popc(mask));
            base = __shfl_sync(PT_FULL, base, 0);
            if (em) out[base + __popc(mask & ((1u << pt_lane()) - 1u))] = i;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// A6: TLAS inputs — world boxes of instances, and the leaf-ordered instance records
// ---------------------------------------------------------------------------------------------------
struct PtMeshInfo { float lo[3], hi[3]; float pad; uint32_t node_base, tri_base, ntris, nnodes; };
// caller's instance list (64-byte ABI records: mesh id, 3 pad words, 12 floats) -> device instance records with the inverse transform
struct PtInstanceIn { uint32_t mesh_id, r0, r1, r2; float o2w[12]; };
__global__ void __launch_bounds__(256) k_inst_prepare(const PtInstanceIn* in, uint32_t n, PtInstance* out, uint32_t* status) {
    for (uint32_t i = pt_gtid(); i < n; i += pt_gsize()) {
        PtInstance r;
        for (int k = 0; k < 12; ++k) r.o2w[k] = in[i].o2w[k];
        if (!pt_invert_affine(r.o2w, r.w2o)) atomicOr(status, 2u);
        r.node_base = 0; r.tri_base = 0; r.mesh_id = in[i].mesh_id; r.inst_id = i;
        out[i] = r;
    }
}
__global__ void __launch_bounds__(256) k_inst_boxes(const PtInstance* inst, uint32_t n, const PtMeshInfo* meshes, PtBox* prim_box, uint32_t* bounds) {
    PtOrdBounds ob; pt_ord_init(ob);
    for (uint32_t i = pt_gtid(); i < n; i += pt_gsize()) {
        const PtMeshInfo mi = meshes[inst[i].mesh_id];
        float lo[3], hi[3], wlo[3], whi[3], o2w[12];
        for (int k = 0; k < 3; ++k) { lo[k] = mi.lo[k] - mi.pad; hi[k] = mi.hi[k] + mi.pad; }
        for (int k = 0; k < 12; ++k) o2w[k] = inst[i].o2w[k];
        pt_world_box(o2w, lo, hi, wlo, whi);
        PtBox bx;
        bx.lox = wlo[0]; bx.loy = wlo[1]; bx.loz = wlo[2]; bx.hix = whi[0]; bx.hiy = whi[1]; bx.hiz = whi[2];
        prim_box[i] = bx;
        pt_ord_grow(ob, bx);
    }
    pt_block_reduce_bounds(ob, bounds);
}
__global__ void __launch_bounds__(256) k_write_instances(const PtInstance* in, uint32_t n, const uint32_t* order, const uint32_t* leaf_seq,
                                                         const PtMeshInfo* meshes, PtInstance* out) {
    for (uint32_t k = pt_gtid(); k < n; k += pt_gsize()) {
        PtInstance r = in[order[leaf_seq[k]]];
        r.node_base = meshes[r.mesh_id].node_base; r.tri_base = meshes[r.mesh_id].tri_base;
        out[k] = r;
    }
}

// ---------------------------------------------------------------------------------------------------
// B2 / B5 on explicit ray sets: closest-hit and any-hit kernels (one lane per ray, 128-bit coalesced
// loads of the 32-byte ray records, 128-bit stores of the 16-byte hit records)
// ---------------------------------------------------------------------------------------------------
struct PtDevCounters { unsigned long long nodes, tris, insts; };

// Warp-persistent traversal with dynamic ray fetch (Aila & Laine 2009 "persistent while-while"): every lane
// owns one ray at a time; when fewer than THRESH lanes of the warp are still traversing, the idle lanes claim
// new work items from a global counter with ONE atomic per warp (ballot + popc + shfl), so a few long rays no
// longer hold 31 finished lanes hostage.  Job supplies load(i) -> ray and store(i, hit).
// Resident CTAs per SM the traversal kernels are compiled for.  Flat scenes: 8 x 128 threads x 64 registers = the whole register
// file (measured: forcing 10 or 12 CTAs spills and is 25-40 % slower).  Two-level kernels carry the world ray as well and need 80
// registers: 6 CTAs (forcing 64 registers spills and costs 12 %).
#ifndef PT_TRACE_FLAT_BLOCKS
#define PT_TRACE_FLAT_BLOCKS 8
#endif
#define PT_TRACE_MIN_BLOCKS(two_level) ((two_level) ? 6 : PT_TRACE_FLAT_BLOCKS)
template <bool ANY, bool TWO_LEVEL, class Counter, class Job>
__device__ __forceinline__ void pt_warp_trace(const PtSceneView& sc, Job& job, unsigned long long n, unsigned long long* work_counter, uint32_t* status,
                                              Counter& cnt, int fetch_thresh) {
    PtTravState st;
    PtArrayStack stack;
    PtHitRec best;
    bool active = false, drained = false;   // drained: the global queue is empty (warp-uniform once set)
    unsigned long long idx = 0;
    const uint32_t lane = pt_lane();
    for (;;) {
        uint32_t need = __ballot_sync(PT_FULL, !active);
        if (need && !drained) {
            uint32_t leader = (uint32_t)__ffs(need) - 1u;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(work_counter, (unsigned long long)__popc(need));
            base = __shfl_sync(PT_FULL, base, leader);
            bool got_none = false;
            if (!active) {
                idx = base + (unsigned long long)__popc(need & ((1u << lane) - 1u));
                if (idx < n) {
                    pt_v3 o, d; float tmin, tmax;
                    job.load(idx, &o, &d, &tmin, &tmax);
                    pt_trav_init<TWO_LEVEL>(&st, o, d, tmin, tmax, &best, sc.tlas_base);
                    active = true;
                } else got_none = true;
            }
            drained = __any_sync(PT_FULL, got_none);
        }
        if (!__any_sync(PT_FULL, active)) break;
        while (active) {
            if (pt_trav_step<ANY, TWO_LEVEL>(sc, &st, stack, &best, cnt) == PT_STEP_DONE) {
                if (st.sp < 0) atomicOr(status, 1u);       // traversal-stack overflow (pt_trav_step left sp = -1)
                job.store(idx, best, sc, st.world.o, st.world.d);
                active = false;
                break;
            }
            if (!drained && __popc(__activemask()) < fetch_thresh) break;
        }
    }
}

// B2 / B5 on explicit ray sets: 32-byte ray records read with two 128-bit loads, 16-byte hit records written
// with one 128-bit store.
template <bool ANY, bool TWO_LEVEL>
struct PtRaySetJob {
    const float4* __restrict__ rays; float4* __restrict__ hits; uint32_t* __restrict__ inst_out; uint8_t* __restrict__ occ;
    __device__ __forceinline__ void load(unsigned long long i, pt_v3* o, pt_v3* d, float* tmin, float* tmax) const {
        float4 a = __ldcs(rays + 2 * i), b = __ldcs(rays + 2 * i + 1);   // streaming: read once, do not displace BVH nodes from L2
        *o = pt_mk(a.x, a.y, a.z); *d = pt_mk(b.x, b.y, b.z); *tmin = a.w; *tmax = b.w;
    }
    __device__ __forceinline__ void store(unsigned long long i, const PtHitRec& h, const PtSceneView& sc, pt_v3 wo, pt_v3 wd) const {
        if (ANY) { occ[i] = h.prim != PT_NONE ? 1 : 0; return; }
        float4 o;
        if (h.prim == PT_NONE) { o.x = __uint_as_float(PT_INF_BITS); o.y = 0.0f; o.z = 0.0f; }
        else { uint32_t mat; o.x = h.t; pt_hit_bary<TWO_LEVEL>(sc, h, wo, wd, &o.y, &o.z, &mat); }
        o.w = __uint_as_float(h.prim);
        __stcs(hits + i, o);                                                  // streaming store (evict-first)
        if (inst_out) __stcs(inst_out + i, h.inst);
    }
};

template <bool ANY, bool TWO_LEVEL, bool COUNT>
__global__ void __launch_bounds__(128, PT_TRACE_MIN_BLOCKS(TWO_LEVEL)) k_trace_rays(PtSceneView sc, const float4* __restrict__ rays, unsigned long long n, float4* __restrict__ hits,
                                                    uint32_t* __restrict__ inst_out, uint8_t* __restrict__ occ, uint32_t* status, PtDevCounters* counters,
                                                    unsigned long long* work_counter, int fetch_thresh) {
    PtRaySetJob<ANY, TWO_LEVEL> job; job.rays = rays; job.hits = hits; job.inst_out = inst_out; job.occ = occ;
    if (COUNT) {
        PtCount cnt; cnt.nodes = cnt.tris = cnt.insts = 0;
        pt_warp_trace<ANY, TWO_LEVEL>(sc, job, n, work_counter, status, cnt, fetch_thresh);
        atomicAdd(&counters->nodes, (unsigned long long)cnt.nodes); atomicAdd(&counters->tris, (unsigned long long)cnt.tris);
        atomicAdd(&counters->insts, (unsigned long long)cnt.insts);
    } else {
        PtNoCount nc;
        pt_warp_trace<ANY, TWO_LEVEL>(sc, job, n, work_counter, status, nc, fetch_thresh);
    }
}

// Exhaustive closest hit: every triangle of every instance, no BVH (device-side ground truth for tests).
// One lane per ray; all lanes of a block read the same triangle (broadcast), staged through shared memory.
template <bool TWO_LEVEL>
__global__ void __launch_bounds__(128) k_trace_brute(PtSceneView sc, const PtInstance* inst_in, uint32_t num_inst, const PtMeshInfo* meshes,
                                                     uint32_t flat_ntris, const float4* __restrict__ rays, unsigned long long n, float4* __restrict__ hits,
                                                     uint32_t* __restrict__ inst_out) {
    __shared__ PtU4 tile[3 * 128];
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    float4 a = valid ? rays[2 * i] : make_float4(0, 0, 0, 0), b = valid ? rays[2 * i + 1] : make_float4(0, 0, 0, 0);
    PtHitRec best;
    best.t = b.w; best.prim = PT_NONE; best.inst = PT_NONE; best.tidx = best.iidx = 0;
#if !PT_SLIM_HIT
    best.U = best.V = 0; best.ad = 1; best.mat = 0;
#endif
    PtNoCount nc;
    uint32_t ninst = TWO_LEVEL ? num_inst : 1u;
    for (uint32_t ii = 0; ii < ninst; ++ii) {
        PtRayCtx r;
        uint32_t tri_base = 0, ntris = flat_ntris, inst_id = 0;
        if (TWO_LEVEL) {
            float w2o[12];
            for (int k = 0; k < 12; ++k) w2o[k] = inst_in[ii].w2o[k];
            pt_ray_ctx(&r, pt_xform_point(w2o, pt_mk(a.x, a.y, a.z)), pt_xform_vec(w2o, pt_mk(b.x, b.y, b.z)));
            tri_base = meshes[inst_in[ii].mesh_id].tri_base; ntris = meshes[inst_in[ii].mesh_id].ntris; inst_id = inst_in[ii].inst_id;
        } else pt_ray_ctx(&r, pt_mk(a.x, a.y, a.z), pt_mk(b.x, b.y, b.z));
        for (uint32_t t0 = 0; t0 < ntris; t0 += 128) {
            __syncthreads();
            uint32_t cntt = min(128u, ntris - t0);
            for (uint32_t k = threadIdx.x; k < 3 * cntt; k += blockDim.x) tile[k] = pt_load4(sc.tris + 3 * (size_t)(tri_base + t0) + k);
            __syncthreads();
            if (valid)
                for (uint32_t k = 0; k < cntt; ++k) pt_test_tri_words(tile[3 * k], tile[3 * k + 1], tile[3 * k + 2], tri_base + t0 + k, r, a.w, inst_id, ii, &best, nc);   // shared-memory words
        }
    }
    if (!valid) return;
    float4 o;
    if (best.prim == PT_NONE) { o.x = __uint_as_float(PT_INF_BITS); o.y = 0.0f; o.z = 0.0f; }
    else {
        PtSceneView sv = sc;
        if (TWO_LEVEL) sv.instances = reinterpret_cast<const PtU4*>(inst_in);      // the exhaustive search walks the CALLER-order instance list: best.iidx indexes it
        uint32_t mat;
        o.x = best.t; pt_hit_bary<TWO_LEVEL>(sv, best, pt_mk(a.x, a.y, a.z), pt_mk(b.x, b.y, b.z), &o.y, &o.z, &mat);
    }
    o.w = __uint_as_float(best.prim);
    hits[i] = o;
    if (inst_out) inst_out[i] = best.inst;
}

// ---------------------------------------------------------------------------------------------------
// Wavefront path tracer state (SoA, one float4 / uint4 array per field; slot = one owned pixel)
// ---------------------------------------------------------------------------------------------------
struct PtWaveCounters {
    uint32_t n_active, n_next, n_shadow, pad;
    unsigned long long total_extend, total_shadow;
    unsigned long long work_extend, work_connect;   // dynamic-fetch cursors of the two traversal kernels
};
struct PtWave {
    float4* ray_o;     // o.xyz, pdf_prev
    float4* ray_d;     // d.xyz, bounce (uint bits)
    float4* beta;      // beta.rgb, pixel (uint bits)
    float4* L;         // L.rgb, unused
    uint4* rng;        // state lo/hi, inc lo/hi
    float4* hit;       // t, tidx (bits), iidx (bits), sort key (bits)
    float2* hit_uv;    // barycentric weights of vertex 1 and 2 of the hit; only allocated when a mesh carries uv / colour streams (else NULL)
    uint32_t* active;  // slots alive this bounce
    uint32_t* next;    // slots alive next bounce
    uint32_t* sorted;  // active, reordered by material key
    float4* sh_o;      // shadow ray o.xyz, tmax
    float4* sh_d;      // shadow ray d.xyz, slot (bits)
    float4* sh_c;      // contribution rgb
    const uint32_t* slot_pixel;   // slot -> pixel index (tile partition), or nullptr = identity
    PtWaveCounters* ctr;
    uint32_t* key_hist;           // PT_KEY_BUCKETS + 1 counters for the material sort
    uint32_t num_slots;           // slots in flight this wave = num_pixels * samples in the wave (sample-major)
    uint32_t num_pixels;          // owned pixels
};
#define PT_KEY_BUCKETS 1024u      // material ids >= 1023 share the last bucket; bucket 1023+1 = miss
#define PT_KEY_MISS PT_KEY_BUCKETS

struct PtFrame {
    PtCamera cam; uint64_t seed; uint32_t width, height, sample, flags;
};

// B1: ray generation (one lane per slot)
__global__ void __launch_bounds__(256) k_raygen(PtWave w, PtFrame f) {
    for (uint32_t s = pt_gtid(); s < w.num_slots; s += pt_gsize()) {
        uint32_t ps = s % w.num_pixels, k = s / w.num_pixels;     // several samples of every owned pixel share one wave
        uint32_t pixel = w.slot_pixel ? w.slot_pixel[ps] : ps;
        PtPath p;
        pt_path_init(&p, f.cam, f.seed, pixel, f.sample + k, f.width, f.height, f.flags);
        w.ray_o[s] = make_float4(p.o.x, p.o.y, p.o.z, 0.0f);
        w.ray_d[s] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(0u));
        w.beta[s] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pixel));
        w.L[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        w.rng[s] = make_uint4((uint32_t)p.rng.state, (uint32_t)(p.rng.state >> 32), (uint32_t)p.rng.inc, (uint32_t)(p.rng.inc >> 32));
        w.active[s] = s;
    }
    if (pt_gtid() == 0) { w.ctr->n_active = w.num_slots; w.ctr->n_next = 0; w.ctr->n_shadow = 0; w.ctr->work_extend = 0; w.ctr->work_connect = 0; }
}

// B2: extend — closest hit for every active path (warp-persistent, dynamic fetch over the active list);
// writes the hit record with the material sort key.
template <bool ATTR>
struct PtExtendJob {
    PtWave w; const PtU4* tris; bool two_level;
    __device__ __forceinline__ void load(unsigned long long j, pt_v3* o, pt_v3* d, float* tmin, float* tmax) const {
        uint32_t s = w.active[j];
        float4 a = w.ray_o[s], b = w.ray_d[s];
        *o = pt_mk(a.x, a.y, a.z); *d = pt_mk(b.x, b.y, b.z); *tmin = 0.0f; *tmax = __uint_as_float(PT_INF_BITS);
    }
#if PT_SLIM_HIT
    template <bool TWO_LEVEL>
    __device__ __forceinline__ void store_impl(unsigned long long j, const PtHitRec& h, const PtSceneView& sc, pt_v3 wo, pt_v3 wd) const {
        uint32_t s = w.active[j];
        uint32_t key = PT_KEY_MISS;
        if (h.prim != PT_NONE) {
            float u, v; uint32_t mat;
            if (w.hit_uv) { pt_hit_bary<TWO_LEVEL>(sc, h, wo, wd, &u, &v, &mat); w.hit_uv[s] = make_float2(u, v); }   // only scenes with uv / colour streams need them
            else mat = pt_ldg4(tris + 3 * (size_t)h.tidx + 1).w;                                                  // material id: word 1 .w of the hit triangle, once per ray
            key = min(mat, PT_KEY_BUCKETS - 1u);
        }
        w.hit[s] = make_float4(h.t, __uint_as_float(h.tidx), __uint_as_float(h.iidx), __uint_as_float(key));
    }
    __device__ __forceinline__ void store(unsigned long long j, const PtHitRec& h, const PtSceneView& sc, pt_v3 wo, pt_v3 wd) const {
        if (two_level) store_impl<true>(j, h, sc, wo, wd); else store_impl<false>(j, h, sc, wo, wd);
    }
#else
    __device__ __forceinline__ void store(unsigned long long j, const PtHitRec& h, const PtSceneView&, pt_v3, pt_v3) const {
        uint32_t s = w.active[j];
        uint32_t key = PT_KEY_MISS;
        if (h.prim != PT_NONE) key = min(h.mat, PT_KEY_BUCKETS - 1u);
        w.hit[s] = make_float4(h.t, __uint_as_float(h.tidx), __uint_as_float(h.iidx), __uint_as_float(key));
        if (ATTR && h.prim != PT_NONE) w.hit_uv[s] = make_float2(pt_div(h.U, h.ad), pt_div(h.V, h.ad));   // only scenes with uv / colour streams carry them (separate
                                                                                                           // instantiation: the two divisions cost the default kernel a spill)
    }
#endif
};
template <bool TWO_LEVEL, bool ATTR>
__global__ void __launch_bounds__(128, PT_TRACE_MIN_BLOCKS(TWO_LEVEL)) k_extend(PtSceneView sc, PtWave w, uint32_t* status, int fetch_thresh) {
    PtExtendJob<ATTR> job; job.w = w; job.tris = sc.tris; job.two_level = TWO_LEVEL;
    PtNoCount nc;
    pt_warp_trace<false, TWO_LEVEL>(sc, job, (unsigned long long)w.ctr->n_active, &w.ctr->work_extend, status, nc, fetch_thresh);
}
// histogram of the sort keys (warp-aggregated: one atomic per distinct key per warp)
__global__ void __launch_bounds__(256) k_key_hist(PtWave w) {
    const uint32_t n = w.ctr->n_active;
    const uint32_t rounds = (n + pt_gsize() - 1) / pt_gsize();
    for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t j = r * pt_gsize() + pt_gtid();
        bool valid = j < n;
        uint32_t key = valid ? __float_as_uint(w.hit[w.active[j]].w) : (0xffff0000u + pt_lane());
        uint32_t peers = __match_any_sync(PT_FULL, key);
        if (valid && (uint32_t)(__ffs(peers) - 1) == pt_lane()) atomicAdd(&w.key_hist[key], (uint32_t)__popc(peers));
    }
}

// B6: counting sort of the active list by material key.  key_hist holds counts -> exclusive offsets (one block).
__global__ void __launch_bounds__(1024) k_key_scan(uint32_t* key_hist) {
    uint32_t carry = 0;
    for (uint32_t b = 0; b < PT_KEY_BUCKETS + 1u; b += 1024) {
        uint32_t i = b + threadIdx.x;
        uint32_t v = i < PT_KEY_BUCKETS + 1u ? key_hist[i] : 0, total;
        uint32_t e = pt_block_excl_scan(v, &total);
        if (i < PT_KEY_BUCKETS + 1u) key_hist[i] = carry + e;
        carry += total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_key_scatter(PtWave w) {
    const uint32_t n = w.ctr->n_active;
    const uint32_t rounds = (n + pt_gsize() - 1) / pt_gsize();
    for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t j = r * pt_gsize() + pt_gtid();
        bool valid = j < n;
        uint32_t s = valid ? w.active[j] : 0u;
        uint32_t key = valid ? __float_as_uint(w.hit[s].w) : (0xffff0000u + pt_lane());
        // warp-aggregated slot claim: lanes with equal keys share one atomic
        uint32_t peers = __match_any_sync(PT_FULL, key);
        uint32_t leader = (uint32_t)__ffs(peers) - 1u, rank = __popc(peers & ((1u << pt_lane()) - 1u));
        uint32_t base = 0;
        if (valid && pt_lane() == leader) base = atomicAdd(&w.key_hist[key], (uint32_t)__popc(peers));
        base = __shfl_sync(PT_FULL, base, leader);
        if (valid) w.sorted[base + rank] = s;
    }
}
__global__ void k_key_clear(uint32_t* key_hist) {
    for (uint32_t i = pt_gtid(); i < PT_KEY_BUCKETS + 1u; i += pt_gsize()) key_hist[i] = 0;
}

struct PtShadeScene {
    PtSceneView sv;
    const PtMaterial* mats; uint32_t num_mats;
    const PtMeshAttr* mesh_attr;      // per mesh, or NULL when no mesh has uv / colour streams (the default path is then untouched)
    const PtTexture* textures; const uint32_t* mat_tex;
    PtShadeConsts sc;
};

// B3 + B4 (+ the compaction half of B6): shade every active path; ballot/popc compaction of survivors and
// of the emitted shadow rays (one atomic per warp each).
#ifndef PT_SHADE_MIN_BLOCKS
#define PT_SHADE_MIN_BLOCKS 6     // 80 registers, no spill: -3 % / -6 % shade time on configs 2 / 4, +-0 on config 3 vs 100 registers; 64 registers (8 blocks) spills and is 20 % slower (profiles/r02_ab_shade_occupancy.log)
#endif
template <bool TWO_LEVEL>
__global__ void __launch_bounds__(128, PT_SHADE_MIN_BLOCKS) k_shade(PtShadeScene ss, PtWave w, const uint32_t* list) {
    const uint32_t n = w.ctr->n_active;
    const uint32_t rounds = (n + pt_gsize() - 1) / pt_gsize();
    for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t j = r * pt_gsize() + pt_gtid();
        bool valid = j < n, alive = false, shadow = false;
        uint32_t s = 0;
        PtShadowRay sh; sh.valid = false;
        if (valid) {
            s = list[j];
            float4 o = w.ray_o[s], d = w.ray_d[s], be = w.beta[s], L = w.L[s], h = w.hit[s];
            uint4 g = w.rng[s];
            PtPath p;
            p.o = pt_mk(o.x, o.y, o.z); p.d = pt_mk(d.x, d.y, d.z); p.beta = pt_mk(be.x, be.y, be.z); p.L = pt_mk(L.x, L.y, L.z);
            p.rng.state = ((uint64_t)g.y << 32) | g.x; p.rng.inc = ((uint64_t)g.w << 32) | g.z;
            p.pdf_prev = o.w; p.pixel = __float_as_uint(be.w); p.bounce = __float_as_uint(d.w);
            uint32_t key = __float_as_uint(h.w);
            if (key == PT_KEY_MISS) pt_shade_miss(&p, ss.sc);
            else {
                uint32_t tidx = __float_as_uint(h.y);
                PtU4 t1 = pt_ldg4(ss.sv.tris + 3 * (size_t)tidx + 1), t2 = pt_ldg4(ss.sv.tris + 3 * (size_t)tidx + 2);
                pt_v3 e1 = pt_mk(__uint_as_float(t1.x), __uint_as_float(t1.y), __uint_as_float(t1.z));
                pt_v3 e2 = pt_mk(__uint_as_float(t2.x), __uint_as_float(t2.y), __uint_as_float(t2.z));
                if (TWO_LEVEL) {
                    const PtU4* ip = ss.sv.instances + 7 * (size_t)__float_as_uint(h.z);
                    PtU4 m3 = pt_ldg4(ip + 3), m4 = pt_ldg4(ip + 4), m5 = pt_ldg4(ip + 5);
                    float o2w[12] = {__uint_as_float(m3.x), __uint_as_float(m3.y), __uint_as_float(m3.z), __uint_as_float(m3.w),
                                     __uint_as_float(m4.x), __uint_as_float(m4.y), __uint_as_float(m4.z), __uint_as_float(m4.w),
                                     __uint_as_float(m5.x), __uint_as_float(m5.y), __uint_as_float(m5.z), __uint_as_float(m5.w)};
                    e1 = pt_xform_vec(o2w, e1); e2 = pt_xform_vec(o2w, e2);
                }
                uint32_t mi = t1.w < ss.num_mats ? t1.w : 0u;
                PtMaterial mat = ss.mats[mi];
                if (ss.mesh_attr) {   // base colour x interpolated vertex colour x albedo texel (the reference's fragment shader, Triangle.slang:34-37)
                    uint32_t mesh = 0;
                    if (TWO_LEVEL) mesh = pt_ldg4(ss.sv.instances + 7 * (size_t)__float_as_uint(h.z) + 6).z;
                    const uint32_t prim = pt_ldg4(ss.sv.tris + 3 * (size_t)tidx).w;
                    const float2 buv = w.hit_uv[s];
                    pt_material_apply_attributes(&mat, ss.mesh_attr[mesh], ss.textures, ss.mat_tex[mi], prim, buv.x, buv.y);
                }
                alive = pt_shade_vertex(&p, ss.sc, h.x, e1, e2, mat, &sh);
                shadow = sh.valid;
            }
            w.L[s] = make_float4(p.L.x, p.L.y, p.L.z, 0.0f);
            if (alive) {
                w.ray_o[s] = make_float4(p.o.x, p.o.y, p.o.z, p.pdf_prev);
                w.ray_d[s] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(p.bounce));
                w.beta[s] = make_float4(p.beta.x, p.beta.y, p.beta.z, be.w);
                w.rng[s] = make_uint4((uint32_t)p.rng.state, (uint32_t)(p.rng.state >> 32), g.z, g.w);
            }
        }
        // compaction: survivors -> next list
        uint32_t am = __ballot_sync(PT_FULL, alive);
        if (am) {
            uint32_t base = 0;
            if (pt_lane() == 0) base = atomicAdd(&w.ctr->n_next, (uint32_t)__popc(am));
            base = __shfl_sync(PT_FULL, base, 0);
            if (alive) w.next[base + __popc(am & ((1u << pt_lane()) - 1u))] = s;
        }
        // compaction: shadow rays -> shadow queue
        uint32_t sm = __ballot_sync(PT_FULL, shadow);
        if (sm) {
            uint32_t base = 0;
            if (pt_lane() == 0) base = atomicAdd(&w.ctr->n_shadow, (uint32_t)__popc(sm));
            base = __shfl_sync(PT_FULL, base, 0);
            if (shadow) {
                uint32_t q = base + __popc(sm & ((1u << pt_lane()) - 1u));
                w.sh_o[q] = make_float4(sh.o.x, sh.o.y, sh.o.z, sh.tmax);
                w.sh_d[q] = make_float4(sh.d.x, sh.d.y, sh.d.z, __uint_as_float(s));
                w.sh_c[q] = make_float4(sh.contrib.x, sh.contrib.y, sh.contrib.z, 0.0f);
            }
        }
    }
}

// B5: connect — any-hit traversal of the shadow queue; unoccluded contributions are added to the path's L
struct PtConnectJob {
    PtWave w;
    __device__ __forceinline__ void load(unsigned long long q, pt_v3* o, pt_v3* d, float* tmin, float* tmax) const {
        float4 a = w.sh_o[q], b = w.sh_d[q];
        *o = pt_mk(a.x, a.y, a.z); *d = pt_mk(b.x, b.y, b.z); *tmin = 0.0f; *tmax = a.w;
    }
    __device__ __forceinline__ void store(unsigned long long q, const PtHitRec& h, const PtSceneView&, pt_v3, pt_v3) const {
        if (h.prim != PT_NONE) return;
        uint32_t s = __float_as_uint(w.sh_d[q].w);
        float4 c = w.sh_c[q], L = w.L[s];
        w.L[s] = make_float4(L.x + c.x, L.y + c.y, L.z + c.z, 0.0f);   // one shadow ray per slot per bounce: no atomics needed
    }
};
template <bool TWO_LEVEL>
__global__ void __launch_bounds__(128, PT_TRACE_MIN_BLOCKS(TWO_LEVEL)) k_connect(PtSceneView sc, PtWave w, uint32_t* status, int fetch_thresh) {
    PtConnectJob job; job.w = w;
    PtNoCount nc;
    pt_warp_trace<true, TWO_LEVEL>(sc, job, (unsigned long long)w.ctr->n_shadow, &w.ctr->work_connect, status, nc, fetch_thresh);
}

// end of bounce: account rays, swap lists
__global__ void k_bounce_end(PtWave w) {
    if (pt_gtid() != 0) return;
    w.ctr->total_extend += w.ctr->n_active; w.ctr->total_shadow += w.ctr->n_shadow;
    w.ctr->n_active = w.ctr->n_next; w.ctr->n_next = 0; w.ctr->n_shadow = 0; w.ctr->work_extend = 0; w.ctr->work_connect = 0;
}

// B7: accumulate the wave's samples into the frame (one add per pixel per sample, in sample order: deterministic).
// C1, fused form: when this context is a non-root member of a multi-GPU frame in direct mode, `remote` is the ROOT's accumulation buffer
// mapped into this device's address space (peer access inside one process, CUDA IPC across processes).  The finished sum of every
// owned pixel is stored there too, so the gather over NVLink happens tile by tile inside the producing kernel and the collective
// that follows is only a barrier.  Remote stores only (no remote read): the owner keeps the master copy.
__global__ void __launch_bounds__(256) k_accumulate(PtWave w, float4* accum, float4* remote) {
    const uint32_t samples = w.num_slots / w.num_pixels;
    for (uint32_t ps = pt_gtid(); ps < w.num_pixels; ps += pt_gsize()) {
        uint32_t pixel = w.slot_pixel ? w.slot_pixel[ps] : ps;
        float4 a = accum[pixel];
        for (uint32_t k = 0; k < samples; ++k) {                    // ascending sample index: the same order as one sample per wave
            float4 L = w.L[(size_t)k * w.num_pixels + ps];
            a = make_float4(a.x + L.x, a.y + L.y, a.z + L.z, a.w + 1.0f);
        }
        accum[pixel] = a;
        if (remote) remote[pixel] = a;
    }
}
// zero the owned pixels only (direct mode: the other pixels of the root's frame belong to their owners' remote stores)
__global__ void __launch_bounds__(256) k_clear_owned(float4* accum, const uint32_t* slot_pixel, uint32_t n) {
    for (uint32_t i = pt_gtid(); i < n; i += pt_gsize()) accum[slot_pixel[i]] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ---------------------------------------------------------------------------------------------------
// C1: gather of the tile-partitioned frame (NCCL form).  A rank packs its owned pixels in slot order (= scanline order of the owned
// pixels), sends them to the root with ncclSend; the root receives every rank's block and scatters it into its frame.  The root
// needs no per-rank pixel list: the position of pixel (x, y) inside its owner's block follows from the tile rule — row_base[r][y]
// pixels of rank r lie in the rows above, and in row y the owner's tiles left of tile tx are all full width.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_owned(const float4* __restrict__ accum, const uint32_t* __restrict__ slot_pixel, uint32_t n, float4* __restrict__ out) {
    for (uint32_t i = pt_gtid(); i < n; i += pt_gsize()) out[i] = accum[slot_pixel[i]];
}
struct PtGatherPlan { uint32_t width, height, tile, count, root, only /* PT_NONE = every rank but the root, else this rank's pixels only */; };
__global__ void __launch_bounds__(256) k_unpack_gathered(PtGatherPlan g, const float4* __restrict__ stage, const uint32_t* __restrict__ stage_off /* [count] */,
                                                         const uint32_t* __restrict__ row_base /* [count][height] */, float4* __restrict__ accum) {
    const uint32_t n = g.width * g.height;
    for (uint32_t p = pt_gtid(); p < n; p += pt_gsize()) {
        const uint32_t x = p % g.width, y = p / g.width, tx = x / g.tile, ty = y / g.tile;
        const uint32_t r = (tx + ty) % g.count;
        if (r == g.root || (g.only != PT_NONE && r != g.only)) continue;
        const uint32_t a = (r + g.count - ty % g.count) % g.count;            // first tile column of rank r in this tile row
        const uint32_t before = tx > a ? (tx - a - 1u) / g.count + 1u : 0u;   // owned tiles left of tx (all full width)
        accum[p] = stage[stage_off[r] + row_base[r * g.height + y] + before * g.tile + (x - tx * g.tile)];
    }
}
__global__ void __launch_bounds__(256) k_resolve_rgba8(const float4* accum, uint32_t n, uint32_t* out) {
    for (uint32_t i = pt_gtid(); i < n; i += pt_gsize()) {
        float4 a = accum[i];
        float inv = a.w > 0.0f ? pt_div(1.0f, a.w) : 0.0f;
        uint32_t r = (uint32_t)(pt_clamp(a.x * inv, 0.0f, 1.0f) * 255.0f + 0.5f), g = (uint32_t)(pt_clamp(a.y * inv, 0.0f, 1.0f) * 255.0f + 0.5f),
                 b = (uint32_t)(pt_clamp(a.z * inv, 0.0f, 1.0f) * 255.0f + 0.5f);
        out[i] = r | (g << 8) | (b << 16) | (a.w > 0.0f ? 0xff000000u : 0u);
    }
}

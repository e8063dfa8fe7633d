// pt_host_shared.h — host-side derivations whose RESULT is data consumed by both the device kernels and
// the CPU oracle (camera basis, instance inverse transforms and world boxes, light records).  They run in
// double on the host and are rounded once to float, so both sides see identical bits.
// Camera conventions: column-major view/proj exactly as `struct uniform_buffer`
// (mos9527/Foundation src/Renderer/Renderer.cpp:28-33), right-handed lookAt, Vulkan clip space with the
// Y flip applied by the caller (Renderer.cpp:373-380).
#pragma once
#include "pt_layout.h"

static inline bool pt_invert4(const double* m, double* out) {  // column-major 4x4, Gauss-Jordan, partial pivoting
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { a[r][c] = m[c * 4 + r]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
        int piv = col; double best = a[col][col] < 0 ? -a[col][col] : a[col][col];
        for (int r = col + 1; r < 4; ++r) { double v = a[r][col] < 0 ? -a[r][col] : a[r][col]; if (v > best) { best = v; piv = r; } }
        if (best == 0.0) return false;
        if (piv != col) for (int c = 0; c < 8; ++c) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
        double inv = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] *= inv;
        for (int r = 0; r < 4; ++r) if (r != col) { double f = a[r][col]; if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c]; }
    }
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[c * 4 + r] = a[r][4 + c];
    return true;
}

static inline bool pt_camera_derive(const float* view, const float* proj, PtCamera* cam) {
    double v[16], p[16], m[16], mi[16], vi[16];
    for (int i = 0; i < 16; ++i) { v[i] = view[i]; p[i] = proj[i]; }
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) { double s = 0; for (int k = 0; k < 4; ++k) s += p[k * 4 + r] * v[c * 4 + k]; m[c * 4 + r] = s; }
    if (!pt_invert4(m, mi) || !pt_invert4(v, vi)) return false;
    double eye[3] = {vi[12] / vi[15], vi[13] / vi[15], vi[14] / vi[15]};
    double cx[3][3];
    const double ndc[3][2] = {{-1, -1}, {1, -1}, {-1, 1}};
    for (int k = 0; k < 3; ++k) {
        double x = ndc[k][0], y = ndc[k][1], z = 0.0, h[4];
        for (int r = 0; r < 4; ++r) h[r] = mi[0 * 4 + r] * x + mi[1 * 4 + r] * y + mi[2 * 4 + r] * z + mi[3 * 4 + r];
        if (h[3] == 0.0) return false;
        for (int r = 0; r < 3; ++r) cx[k][r] = h[r] / h[3];
    }
    for (int r = 0; r < 3; ++r) {
        double dx = (cx[1][r] - cx[0][r]) * 0.5, dy = (cx[2][r] - cx[0][r]) * 0.5;
        cam->eye[r] = (float)eye[r];
        cam->dx[r] = (float)dx; cam->dy[r] = (float)dy;
        cam->d0[r] = (float)(cx[0][r] + dx + dy - eye[r]);
    }
    return true;
}

// world->object rows from object->world rows (affine 3x4), in double.  __host__ __device__: IEEE double + - * / in a fixed order, no
// contraction on either side, so the device (k_inst_prepare) and the oracle (host) produce the same floats.
PT_HD bool pt_invert_affine(const float* o2w, float* w2o) {
    double a = o2w[0], b = o2w[1], c = o2w[2], d = o2w[4], e = o2w[5], f = o2w[6], g = o2w[8], h = o2w[9], i = o2w[10];
    double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    double det = a * A + b * B + c * C;
    if (det == 0.0) return false;
    double id = 1.0 / det;
    double r[9] = {A * id, -(b * i - c * h) * id, (b * f - c * e) * id, B * id, (a * i - c * g) * id, -(a * f - c * d) * id,
                   C * id, -(a * h - b * g) * id, (a * e - b * d) * id};
    double tx = o2w[3], ty = o2w[7], tz = o2w[11];
    for (int k = 0; k < 3; ++k) {
        w2o[k * 4 + 0] = (float)r[k * 3 + 0]; w2o[k * 4 + 1] = (float)r[k * 3 + 1]; w2o[k * 4 + 2] = (float)r[k * 3 + 2];
        w2o[k * 4 + 3] = (float)(-(r[k * 3 + 0] * tx + r[k * 3 + 1] * ty + r[k * 3 + 2] * tz));
    }
    return true;
}

// world box of an object-space box under o2w: min/max over the 8 transformed corners (float, pt_xform_point)
PT_HD void pt_world_box(const float* o2w, const float* lo, const float* hi, float* wlo, float* whi) {
    for (int k = 0; k < 8; ++k) {
        pt_v3 c = pt_xform_point(o2w, pt_mk((k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2]));
        if (k == 0) { wlo[0] = whi[0] = c.x; wlo[1] = whi[1] = c.y; wlo[2] = whi[2] = c.z; }
        else {
            wlo[0] = pt_min(wlo[0], c.x); whi[0] = pt_max(whi[0], c.x);
            wlo[1] = pt_min(wlo[1], c.y); whi[1] = pt_max(whi[1], c.y);
            wlo[2] = pt_min(wlo[2], c.z); whi[2] = pt_max(whi[2], c.z);
        }
    }
}

// padding applied to child boxes of a BVH whose coordinates are bounded by [lo, hi]
PT_HD float pt_pad_for(const float* lo, const float* hi) {
    float m = 0.0f;
    for (int k = 0; k < 3; ++k) { m = pt_max(m, pt_abs(lo[k])); m = pt_max(m, pt_abs(hi[k])); }
    return pt_max(m * PT_PAD_REL, 1e-30f);
}

// light record from a world-space triangle; returns its area (0.5 |e1 x e2|)
PT_HD float pt_light_make(PtLight* l, pt_v3 v0, pt_v3 e1, pt_v3 e2, float er, float eg, float eb) {
    pt_v3 c = pt_cross(e1, e2);
    float area = 0.5f * pt_sqrt(pt_dot(c, c));
    l->v0x = v0.x; l->v0y = v0.y; l->v0z = v0.z; l->cdf = 0.0f;
    l->e1x = e1.x; l->e1y = e1.y; l->e1z = e1.z; l->area = area;
    l->e2x = e2.x; l->e2y = e2.y; l->e2z = e2.z; l->pad0 = 0.0f;
    l->emr = er; l->emg = eg; l->emb = eb; l->pad1 = 0.0f;
    return area;
}
// Drops degenerate emitters (area not > 0: a zero-area or NaN triangle would put NaN into every cdf and, with a total area of 0,
// inf / NaN into the NEE pdf), then writes the cumulative area fractions (sequential double prefix, rounded once).  *n is updated to
// the number of lights kept; returns the total area as float (0 when no light is left: callers then run without NEE).
static inline float pt_lights_finalize(PtLight* l, uint32_t* n) {
    uint32_t m = 0;
    for (uint32_t i = 0; i < *n; ++i) if (l[i].area > 0.0f && l[i].area < 3.0e38f) l[m++] = l[i];
    *n = m;
    double total = 0.0;
    for (uint32_t i = 0; i < m; ++i) total += (double)l[i].area;
    double run = 0.0;
    for (uint32_t i = 0; i < m; ++i) { run += (double)l[i].area; l[i].cdf = (float)(run / total); }
    if (m) l[m - 1].cdf = 1.0f;
    return m ? (float)total : 0.0f;
}
#define PT_RAY_EPS_REL 1.52587890625e-05f  // 2^-16 x largest world extent: secondary-ray origin offset

"""The reference's own material inputs — per-vertex colour + texture coordinate (`vertex_input`, mos9527/Foundation
src/Renderer/Renderer.cpp:23-27, quad data :153-157) and one RGBA8 texture sampled with the RHI sampler's defaults (linear, REPEAT;
src/Platform/RHI/Device.hpp:71-99) and multiplied with the colour (src/Renderer/Triangle.slang:34-37).  CPU tier: the sampler's
arithmetic against a float64 restatement; the quad scene through the oracle.  GPU tier: device == oracle, bit for bit, incl. an interleaved
over-aligned vertex buffer and the two-level (instanced) path."""
import numpy as np
import pytest

from foundation_b200 import scenes
from oracle.pt_oracle import OracleScene, texture_sample


def bilinear_repeat_f64(img, u, v):
    h, w = img.shape[:2]
    x = (u - np.floor(u)) * w - 0.5; y = (v - np.floor(v)) * h - 0.5
    x0 = int(np.floor(x)); y0 = int(np.floor(y)); fx = x - x0; fy = y - y0
    c = lambda xx, yy: img[yy % h, xx % w, :3].astype(np.float64) / 255.0
    a = c(x0, y0) * (1 - fx) + c(x0 + 1, y0) * fx; b = c(x0, y0 + 1) * (1 - fx) + c(x0 + 1, y0 + 1) * fx
    return a * (1 - fy) + b * fy


def test_sampler_is_bilinear_with_repeat_addressing():
    rng = np.random.default_rng(3)
    for size in ((64, 64), (5, 9), (1, 1), (17, 2)):
        img = rng.integers(0, 256, (size[1], size[0], 4), dtype=np.uint8)
        uvs = np.concatenate([rng.uniform(-3, 3, (400, 2)), [[0, 0], [1, 1], [0.999999, 0.5], [-1e-9, 0.25], [0.5 / size[0], 0.5 / size[1]], [1e9, -1e9]]])
        for u, v in uvs:
            got = texture_sample(img, float(np.float32(u)), float(np.float32(v)))
            if abs(u) >= 2 ** 23:
                want = bilinear_repeat_f64(img, 0.0, 0.0)          # no fractional bits left: defined as the fraction 0
            else:
                want = bilinear_repeat_f64(img, float(np.float32(u)), float(np.float32(v)))
            assert np.abs(got - want).max() < 2e-5, (size, u, v, got, want)
    # texel centres reproduce the texel exactly; the seam wraps
    img = rng.integers(0, 256, (4, 4, 4), dtype=np.uint8)
    assert np.allclose(texture_sample(img, 2.5 / 4, 1.5 / 4), img[1, 2, :3] / 255.0, atol=1e-7)
    assert np.allclose(texture_sample(img, 0.0, 0.0), (img[0, 0, :3].astype(float) + img[0, 3, :3] + img[3, 0, :3] + img[3, 3, :3]) / 4 / 255.0, atol=1e-6)


def test_reference_quad_shows_texture_times_vertex_colour():
    """One bounce under a white environment: a pixel's radiance is (to the coat's few per cent) the hit's base colour, so the image must
    correlate with texel x interpolated colour evaluated independently at the pixel centres that hit the quad."""
    sc = scenes.reference_quad(160, 90)
    o = OracleScene(sc)
    spp = 64
    img = o.render(sc.width, sc.height, 5, 0, spp, 1, background=sc.background)[..., :3] / spp
    rays = scenes.camera_rays(sc, 1, 1)          # only to get the camera maths: recompute per pixel centre below
    V = sc.view.astype(np.float64).reshape(4, 4).T; P = sc.proj.astype(np.float64).reshape(4, 4).T
    inv = np.linalg.inv(P @ V); eye = np.linalg.inv(V)[:3, 3]
    ys, xs = np.mgrid[0:sc.height, 0:sc.width]
    ndc = np.stack([2 * (xs + 0.5) / sc.width - 1, 2 * (ys + 0.5) / sc.height - 1, np.zeros_like(xs, float), np.ones_like(xs, float)], -1)
    wp = ndc @ inv.T; wp = wp[..., :3] / wp[..., 3:]
    d = wp - eye; t = -eye[2] / d[..., 2]; hit = eye + d * t[..., None]
    inside = (np.abs(hit[..., 0]) < 0.45) & (np.abs(hit[..., 1]) < 0.45) & (t > 0)
    assert inside.sum() > 300
    # bilinear interpolation of the quad's corner attributes over the unit square [-0.5, 0.5]^2
    m = sc.meshes[0]
    # the quad is two triangles: colour / uv are interpolated per triangle (barycentrics), not over a bilinear patch
    def tri_interp(a, i0, i1, i2, p):
        A, B, C = m.positions[i0, :2].astype(np.float64), m.positions[i1, :2].astype(np.float64), m.positions[i2, :2].astype(np.float64)
        T = np.linalg.inv(np.stack([B - A, C - A], 1)); bc = (p - A) @ T.T
        return a[i0] * (1 - bc[..., :1] - bc[..., 1:2]) + a[i1] * bc[..., :1] + a[i2] * bc[..., 1:2], bc
    p2 = hit[..., :2]
    c1, b1 = tri_interp(m.colors.astype(np.float64), 0, 1, 2, p2); u1, _ = tri_interp(m.uv.astype(np.float64), 0, 1, 2, p2)
    c2, b2 = tri_interp(m.colors.astype(np.float64), 2, 3, 0, p2); u2, _ = tri_interp(m.uv.astype(np.float64), 2, 3, 0, p2)
    in1 = (b1 >= 0).all(-1) & (b1.sum(-1) <= 1)
    col = np.where(in1[..., None], c1, c2); uv = np.where(in1[..., None], u1, u2)
    tex = np.stack([[bilinear_repeat_f64(sc.textures[0], uv[y, x, 0], uv[y, x, 1]) for x in range(sc.width)] for y in range(sc.height)])
    want = tex * col
    got = img[inside]; exp = want[inside]
    corr = np.corrcoef(got.reshape(-1), exp.reshape(-1))[0, 1]
    assert corr > 0.93, corr          # 64 jittered samples per pixel against a point evaluation at the pixel centre of a noisy texture
    ratio = got.sum() / exp.sum()
    assert 0.85 < ratio < 1.1, ratio          # albedo, minus what the dielectric coat reflects specularly, plus its highlight


@pytest.mark.gpu
def test_textured_quad_matches_oracle_on_device(gpu):
    from foundation_b200 import pt
    sc = scenes.reference_quad(320, 180)
    o = OracleScene(sc)
    want = o.render(sc.width, sc.height, 7, 0, 4, 3, background=sc.background)
    with pt.PathTracer(sc.width, sc.height, seed=7, background=sc.background) as tr:
        tr.load(sc)
        tr.render(0, 2, 3); tr.render(2, 2, 3)
        assert np.array_equal(tr.read_accum(), want)
    # the same data as ONE interleaved, over-aligned vertex buffer (pos 16 B | colour 16 B | uv 8 B | pad 8 B = 48-byte stride), like the
    # reference's glm-aligned vertex_input: three pointers into it, one stride
    m = sc.meshes[0]
    buf = np.zeros((4, 12), np.float32); buf[:, 0:3] = m.positions; buf[:, 4:7] = m.colors; buf[:, 8:10] = m.uv; buf[:, 3] = 7; buf[:, 7] = 9; buf[:, 10:] = 11
    raw = buf.view(np.uint8).reshape(-1)
    with pt.PathTracer(sc.width, sc.height, seed=7, background=sc.background) as tr:
        tr.materials_set(sc.materials)
        mid = tr.mesh_create(raw, m.indices.astype(np.uint16), m.material_ids, stride=48)
        tr.mesh_attributes_set(mid, uv=raw[32:], colors=raw[16:], uv_stride=48, color_stride=48)
        tr.texture_create(sc.textures[0]); tr.material_textures_set(sc.material_textures)
        tr.scene_commit(); tr.camera_set(sc.view, sc.proj)
        tr.render(0, 4, 3)
        assert np.array_equal(tr.read_accum(), want)
        # attributes can be dropped again: the frame becomes the untextured one
        tr.mesh_attributes_set(mid, None, None); tr.scene_commit(); tr.render(0, 4, 3)
        plain = scenes.Scene(sc.name, [scenes.Mesh(m.positions, m.indices, m.material_ids)], sc.materials, None, sc.view, sc.proj, sc.width, sc.height, sc.background)
        assert np.array_equal(tr.read_accum(), OracleScene(plain).render(sc.width, sc.height, 7, 0, 4, 3, background=sc.background))


@pytest.mark.gpu
def test_textured_instances_match_oracle_on_device(gpu):
    """Two-level path: the hit's mesh comes from the instance record; one mesh has attributes, the other (the light) has none."""
    from foundation_b200 import pt
    q = scenes.reference_quad(160, 90)
    light = scenes.Mesh(np.asarray([(-1, 1, 3), (1, 1, 3), (1, -1, 3), (-1, -1, 3)], np.float32), np.asarray([(0, 1, 2), (0, 2, 3)], np.uint32), np.ones(2, np.uint32))
    inst = np.zeros(6, scenes.INSTANCE_DTYPE)
    rng = np.random.default_rng(2)
    for i in range(5):
        a = rng.uniform(0, 2 * np.pi); s = rng.uniform(0.5, 1.2); c, sn = np.cos(a) * s, np.sin(a) * s
        inst["transform"][i] = (c, -sn, 0, rng.uniform(-1, 1), sn, c, 0, rng.uniform(-1, 1), 0, 0, s, rng.uniform(-0.5, 0.5))
    inst["mesh_id"][5] = 1; inst["transform"][5] = (1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0)
    mats = np.asarray([scenes._mat((1, 1, 1), 0.6), scenes._mat((0, 0, 0), 1.0, (12, 12, 12))], np.float32)
    sc = scenes.Scene("textured_instances", [q.meshes[0], light], mats, inst, q.view, q.proj, 160, 90, (0.2, 0.25, 0.3), q.textures, np.asarray([0, 0xFFFFFFFF], np.uint32))
    want = OracleScene(sc).render(sc.width, sc.height, 9, 0, 3, 4, background=sc.background)
    with pt.PathTracer(sc.width, sc.height, seed=9, background=sc.background) as tr:
        tr.load(sc)
        tr.render(0, 3, 4)
        got = tr.read_accum()
        assert np.array_equal(got, want), int((got != want).any(-1).sum())

"""Procedural scenes, cameras and ray sets for BASELINE.json's five configs.

Everything is generated from integer seeds with numpy so nothing opaque has to be committed or shipped to
the GPU box.  The output is plain *data* (float32 positions, uint32 indices, material tables, column-major
matrices) that is handed unchanged to the C ABI and to the CPU oracle, so it does not matter that numpy's
transcendental functions may differ between machines in the last bit.

Conventions follow the reference: world is Z-up and right-handed, the camera is glm::lookAt +
glm::infinitePerspective with the Vulkan Y flip (mos9527/Foundation src/Renderer/Renderer.cpp:373-380),
matrices are column-major float[16] like `struct uniform_buffer` (Renderer.cpp:28-33).
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional

import numpy as np


@dataclasses.dataclass
class Mesh:
    positions: np.ndarray      # (V,3) float32
    indices: np.ndarray        # (T,3) uint32
    material_ids: np.ndarray   # (T,)  uint32
    uv: Optional[np.ndarray] = None       # (V,2) float32 — the reference's vertex_input.texCoord
    colors: Optional[np.ndarray] = None   # (V,3) float32 — the reference's vertex_input.color


@dataclasses.dataclass
class Scene:
    name: str
    meshes: List[Mesh]
    materials: np.ndarray                      # (M,8) float32: base rgb, roughness, emission rgb, metallic
    instances: Optional[np.ndarray] = None     # (I,) structured: mesh_id u32, pad 3xu32, transform 12xf32 ; None = flat scene
    view: Optional[np.ndarray] = None          # (16,) float32 column-major
    proj: Optional[np.ndarray] = None
    width: int = 1920
    height: int = 1080
    background: tuple = (0.0, 0.0, 0.0)
    textures: Optional[list] = None             # list of (H,W,4) uint8 RGBA images
    material_textures: Optional[np.ndarray] = None   # (M,) uint32: albedo texture id per material, 0xFFFFFFFF = none

    @property
    def num_triangles(self) -> int:
        return int(sum(m.indices.shape[0] for m in self.meshes))

    @property
    def effective_triangles(self) -> int:
        if self.instances is None:
            return self.num_triangles
        return int(sum(self.meshes[int(i)].indices.shape[0] for i in self.instances["mesh_id"]))


INSTANCE_DTYPE = np.dtype([("mesh_id", "<u4"), ("reserved", "<u4", (3,)), ("transform", "<f4", (12,))])
RAY_DTYPE = np.dtype([("origin", "<f4", (3,)), ("tmin", "<f4"), ("direction", "<f4", (3,)), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])


# ----------------------------------------------------------------------------------------------
# camera maths — restated from glm 1.0.1 (not vendored in the reference): RH lookAt and
# infinitePerspective with the default [-1,1] clip depth, then the reference's proj[1][1] *= -1.
# ----------------------------------------------------------------------------------------------
def look_at(eye, center, up) -> np.ndarray:
    eye = np.asarray(eye, np.float64); center = np.asarray(center, np.float64); up = np.asarray(up, np.float64)
    f = center - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3] = s; m[1, :3] = u; m[2, :3] = -f
    m[0, 3] = -s @ eye; m[1, 3] = -u @ eye; m[2, 3] = f @ eye
    return m.T.reshape(16).astype(np.float32)          # column-major


def infinite_perspective(fovy_rad: float, aspect: float, near: float, flip_y: bool = True) -> np.ndarray:
    r = math.tan(fovy_rad / 2.0) * near
    left, right, bottom, top = -r * aspect, r * aspect, -r, r
    m = np.zeros((4, 4))
    m[0, 0] = 2 * near / (right - left)
    m[1, 1] = 2 * near / (top - bottom)
    m[2, 2] = -1.0
    m[3, 2] = -1.0          # glm result[2][3] = -1 (column 2, row 3)
    m[2, 3] = -2.0 * near   # glm result[3][2] = -2 near (column 3, row 2)
    if flip_y:
        m[1, 1] *= -1.0
    return m.T.reshape(16).astype(np.float32)


def reference_camera(aspect: float = 1920 / 1080):
    """The camera Renderer::Draw sets every frame (Renderer.cpp:373-380)."""
    return look_at((2, 2, 2), (0, 0, 0), (0, 0, 1)), infinite_perspective(math.radians(45.0), aspect, 0.1)


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _quad(a, b, c, d):
    """two triangles a-b-c, a-c-d (normal = (b-a) x (c-a))"""
    return [a, b, c, d], [[0, 1, 2], [0, 2, 3]]


def _mat(base=(0.8, 0.8, 0.8), rough=0.5, emission=(0, 0, 0), metallic=0.0):
    return [base[0], base[1], base[2], rough, emission[0], emission[1], emission[2], metallic]


class _Builder:
    def __init__(self):
        self.pos = []; self.idx = []; self.mat = []; self.nv = 0

    def add(self, verts, tris, mat):
        verts = np.asarray(verts, np.float32).reshape(-1, 3); tris = np.asarray(tris, np.uint32).reshape(-1, 3)
        self.pos.append(verts); self.idx.append(tris + np.uint32(self.nv)); self.mat.append(np.full(len(tris), mat, np.uint32))
        self.nv += len(verts)

    def mesh(self) -> Mesh:
        return Mesh(np.ascontiguousarray(np.concatenate(self.pos), np.float32), np.ascontiguousarray(np.concatenate(self.idx), np.uint32),
                    np.ascontiguousarray(np.concatenate(self.mat), np.uint32))


def _box_faces(lo, hi, rot_z=0.0, skip_bottom=True):
    """5 (or 6) outward-facing quads of an axis-aligned box rotated about Z around its centre."""
    lo = np.asarray(lo, np.float64); hi = np.asarray(hi, np.float64)
    c = (lo + hi) / 2
    cs, sn = math.cos(rot_z), math.sin(rot_z)

    def R(p):
        p = np.asarray(p, np.float64) - c
        return [c[0] + cs * p[0] - sn * p[1], c[1] + sn * p[0] + cs * p[1], c[2] + p[2]]

    x0, y0, z0 = lo; x1, y1, z1 = hi
    quads = [
        [(x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)],   # top    +z
        [(x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)],   # front  -y
        [(x1, y1, z0), (x0, y1, z0), (x0, y1, z1), (x1, y1, z1)],   # back   +y
        [(x0, y1, z0), (x0, y0, z0), (x0, y0, z1), (x0, y1, z1)],   # left   -x
        [(x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)],   # right  +x
    ]
    if not skip_bottom:
        quads.append([(x0, y1, z0), (x1, y1, z0), (x1, y0, z0), (x0, y0, z0)])
    return [[R(p) for p in q] for q in quads]


# ----------------------------------------------------------------------------------------------
# config 1: Cornell box, 32 triangles, one area light
# ----------------------------------------------------------------------------------------------
def cornell_box(width: int = 512, height: int = 512) -> Scene:
    b = _Builder()
    W, RED, GREEN, LIGHT, METAL = 0, 1, 2, 3, 4
    # room [-1,1] x [-1,1] x [0,2], open towards -y; normals face inward
    b.add(*_quad((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)), W)        # floor   (+z)
    b.add(*_quad((-1, 1, 2), (1, 1, 2), (1, -1, 2), (-1, -1, 2)), W)        # ceiling (-z)
    b.add(*_quad((-1, 1, 0), (1, 1, 0), (1, 1, 2), (-1, 1, 2)), W)          # back    (-y)
    b.add(*_quad((-1, -1, 0), (-1, 1, 0), (-1, 1, 2), (-1, -1, 2)), RED)    # left    (+x)
    b.add(*_quad((1, 1, 0), (1, -1, 0), (1, -1, 2), (1, 1, 2)), GREEN)      # right   (-x)
    for q in _box_faces((-0.65, -0.05, 0.0), (-0.05, 0.55, 1.2), rot_z=math.radians(18)):   # tall box, rough metal
        b.add(q, [[0, 1, 2], [0, 2, 3]], METAL)
    for q in _box_faces((0.1, -0.65, 0.0), (0.7, -0.05, 0.6), rot_z=math.radians(-17)):     # short box, white
        b.add(q, [[0, 1, 2], [0, 2, 3]], W)
    b.add(*_quad((-0.25, 0.25, 1.995), (0.25, 0.25, 1.995), (0.25, -0.25, 1.995), (-0.25, -0.25, 1.995)), LIGHT)  # faces -z
    mats = np.asarray([_mat((0.73, 0.73, 0.73), 1.0), _mat((0.65, 0.05, 0.05), 1.0), _mat((0.12, 0.45, 0.15), 1.0),
                       _mat((0, 0, 0), 1.0, (17.0, 12.0, 4.0)), _mat((0.9, 0.8, 0.6), 0.35, metallic=1.0)], np.float32)
    mesh = b.mesh()
    assert mesh.indices.shape[0] == 32
    view = look_at((0, -3.9, 1.0), (0, 0, 1.0), (0, 0, 1))
    proj = infinite_perspective(math.radians(39.0), width / height, 0.1)
    return Scene("cornell_box", [mesh], mats, None, view, proj, width, height)


# ----------------------------------------------------------------------------------------------
# config 2: sphere field, 800 x icosphere(3) = 1,024,000 triangles (+ ground + light)
# ----------------------------------------------------------------------------------------------
def icosphere(subdiv: int):
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.asarray([(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
                    (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.asarray([(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
                    (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)], np.int64)
    for _ in range(subdiv):
        edges = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        uniq, inv = np.unique(edges, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        nf = len(f)
        a = base + inv[:nf]; b = base + inv[nf:2 * nf]; c = base + inv[2 * nf:]
        f = np.concatenate([np.stack([f[:, 0], a, c], 1), np.stack([f[:, 1], b, a], 1), np.stack([f[:, 2], c, b], 1), np.stack([a, b, c], 1)])
    return v, f


def sphere_field(num_spheres: int = 800, subdiv: int = 3, seed: int = 2, width: int = 1920, height: int = 1080) -> Scene:
    rng = np.random.Generator(np.random.PCG64(seed))
    sv, sf = icosphere(subdiv)
    b = _Builder()
    nmat = 16
    mats = [_mat((0.55, 0.55, 0.5), 0.9), _mat((0, 0, 0), 1.0, (30.0, 28.0, 24.0))]     # 0 ground, 1 light
    for k in range(nmat):
        col = tuple(0.25 + 0.7 * rng.random(3))
        if k % 2 == 0:
            mats.append(_mat(col, 1.0))                                                 # Lambert
        else:
            mats.append(_mat(col, 0.05 + 0.55 * rng.random(), metallic=1.0 if k % 4 == 1 else 0.0))   # GGX conductor / coated
    ext = 40.0
    b.add(*_quad((-ext, -ext, 0), (ext, -ext, 0), (ext, ext, 0), (-ext, ext, 0)), 0)
    centres = np.stack([rng.uniform(-30, 30, num_spheres), rng.uniform(-30, 30, num_spheres), rng.uniform(1.0, 9.0, num_spheres)], 1)
    radii = rng.uniform(0.5, 1.6, num_spheres)
    which = rng.integers(0, nmat, num_spheres)
    pos = (sv[None, :, :] * radii[:, None, None] + centres[:, None, :]).reshape(-1, 3)
    idx = (sf[None, :, :] + (np.arange(num_spheres) * len(sv))[:, None, None]).reshape(-1, 3)
    b.add(pos, idx, 0)
    b.mat[-1] = np.repeat(which.astype(np.uint32) + 2, len(sf))
    L = 12.0
    b.add(*_quad((-L, L, 30.0), (L, L, 30.0), (L, -L, 30.0), (-L, -L, 30.0)), 1)        # faces -z
    view = look_at((0, -62, 24), (0, 0, 3), (0, 0, 1))
    proj = infinite_perspective(math.radians(40.0), width / height, 0.1)
    return Scene("sphere_field", [b.mesh()], np.asarray(mats, np.float32), None, view, proj, width, height, (0.25, 0.3, 0.4))


# ----------------------------------------------------------------------------------------------
# config 3: fractal terrain, (n+1)^2 heightfield -> 2 n^2 triangles (n = 2236 -> 9,999,392)
# ----------------------------------------------------------------------------------------------
def _fbm_heightfield(n: int, seed: int, octaves: int = 11, size: float = 100.0, amp: float = 14.0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.linspace(0.0, 1.0, n + 1)
    h = np.zeros((n + 1, n + 1), np.float64)
    for o in range(octaves):
        res = 2 ** (o + 1)
        grid = rng.uniform(-1.0, 1.0, (res + 1, res + 1))
        x = g * res
        i0 = np.minimum(x.astype(np.int64), res - 1); fx = x - i0
        fx = fx * fx * (3 - 2 * fx)                                     # smoothstep -> value noise
        rows = grid[i0, :] * (1 - fx)[:, None] + grid[i0 + 1, :] * fx[:, None]          # (n+1, res+1)
        layer = rows[:, i0] * (1 - fx)[None, :] + rows[:, i0 + 1] * fx[None, :]
        h += layer * (amp * 0.5 ** o)
    return h


def heightfield_mesh(h: np.ndarray, size_x: float, size_y: float, material: int = 0, origin=(0.0, 0.0)) -> Mesh:
    ny, nx = h.shape
    xs = origin[0] + np.linspace(0.0, size_x, nx); ys = origin[1] + np.linspace(0.0, size_y, ny)
    X, Y = np.meshgrid(xs, ys)
    pos = np.stack([X, Y, h], -1).reshape(-1, 3).astype(np.float32)
    j, i = np.meshgrid(np.arange(ny - 1, dtype=np.uint32), np.arange(nx - 1, dtype=np.uint32), indexing="ij")
    v00 = (j * nx + i).reshape(-1); v10 = v00 + 1; v01 = v00 + nx; v11 = v01 + 1
    idx = np.empty((v00.size, 2, 3), np.uint32)
    idx[:, 0, 0] = v00; idx[:, 0, 1] = v10; idx[:, 0, 2] = v11          # normal +z
    idx[:, 1, 0] = v00; idx[:, 1, 1] = v11; idx[:, 1, 2] = v01
    idx = idx.reshape(-1, 3)
    return Mesh(np.ascontiguousarray(pos), np.ascontiguousarray(idx), np.full(idx.shape[0], material, np.uint32))


def fractal_terrain(n: int = 2236, seed: int = 3, width: int = 1920, height: int = 1080, with_light: bool = True) -> Scene:
    size = 100.0
    h = _fbm_heightfield(n, seed, size=size)
    terrain = heightfield_mesh(h, size, size, 0)
    pos, idx, mat = terrain.positions, terrain.indices, terrain.material_ids
    if with_light:
        # an emissive panel above the terrain (2 triangles, faces -z) so NEE has a target; indices appended.  Round 1 had it 40 units
        # above the highest peak, which stretched the scene AABB (the box the incoherent ray origins are drawn from) to
        # mostly empty air; 14 units keeps it above the camera below and the box tight.
        z = float(h.max() + 14.0)
        lp = np.asarray([(20, 80, z), (80, 80, z), (80, 20, z), (20, 20, z)], np.float32)
        base = np.uint32(pos.shape[0])
        pos = np.concatenate([pos, lp]); idx = np.concatenate([idx, np.asarray([[0, 1, 2], [0, 2, 3]], np.uint32) + base])
        mat = np.concatenate([mat, np.asarray([1, 1], np.uint32)])
    mats = np.asarray([_mat((0.45, 0.40, 0.32), 0.8), _mat((0, 0, 0), 1.0, (40.0, 38.0, 34.0))], np.float32)
    zmid = float(h.mean())
    # camera below the light panel, looking down the valley: ~96 % of the primary rays hit the terrain (round 1's camera saw 69 % sky)
    view = look_at((22.0, 12.0, zmid + 27.0), (58.0, 56.0, zmid - 26.0), (0, 0, 1))
    proj = infinite_perspective(math.radians(45.0), width / height, 0.1)
    return Scene(f"fractal_terrain_{n}", [Mesh(np.ascontiguousarray(pos), np.ascontiguousarray(idx), np.ascontiguousarray(mat))], mats, None,
                 view, proj, width, height, (0.35, 0.45, 0.65))


# ----------------------------------------------------------------------------------------------
# config 4: instanced scene — 10,000 instances of a (patch+1)^2 heightfield patch (10,082 tris) = 100.82 M
# ----------------------------------------------------------------------------------------------
def instanced_patches(num_instances: int = 10000, patch: int = 71, seed: int = 5, width: int = 1920, height: int = 1080) -> Scene:
    rng = np.random.Generator(np.random.PCG64(seed))
    h = _fbm_heightfield(patch, seed + 100, octaves=5, amp=0.25)
    g = np.linspace(-1.0, 1.0, patch + 1)
    h = h + 0.5 * np.exp(-3.0 * (g[None, :] ** 2 + g[:, None] ** 2))     # a hill in the middle of the patch
    patch_mesh = heightfield_mesh(h, 1.0, 1.0, 0, origin=(-0.5, -0.5))
    side = int(math.ceil(math.sqrt(num_instances)))
    inst = np.zeros(num_instances + 1, INSTANCE_DTYPE)
    k = np.arange(num_instances)
    gx = (k % side).astype(np.float64); gy = (k // side).astype(np.float64)
    ang = rng.uniform(0, 2 * math.pi, num_instances); sc = rng.uniform(0.8, 1.3, num_instances)
    tx = gx + rng.uniform(-0.2, 0.2, num_instances); ty = gy + rng.uniform(-0.2, 0.2, num_instances); tz = rng.uniform(0.0, 0.4, num_instances)
    T = np.zeros((num_instances, 12))
    T[:, 0] = sc * np.cos(ang); T[:, 1] = -sc * np.sin(ang); T[:, 3] = tx
    T[:, 4] = sc * np.sin(ang); T[:, 5] = sc * np.cos(ang); T[:, 7] = ty
    T[:, 10] = sc; T[:, 11] = tz
    inst["mesh_id"][:num_instances] = 0
    inst["transform"][:num_instances] = T.astype(np.float32)
    # mesh 1: one emissive quad over the field, instanced once with identity
    L = side * 0.3; c = side * 0.5; z = 25.0
    light = _Builder(); light.add(*_quad((c - L, c + L, z), (c + L, c + L, z), (c + L, c - L, z), (c - L, c - L, z)), 1)
    inst["mesh_id"][num_instances] = 1
    inst["transform"][num_instances] = np.asarray([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    mats = np.asarray([_mat((0.5, 0.55, 0.4), 0.6), _mat((0, 0, 0), 1.0, (35.0, 33.0, 30.0))], np.float32)
    view = look_at((-0.15 * side, -0.25 * side, 0.35 * side), (0.5 * side, 0.5 * side, 0.0), (0, 0, 1))
    proj = infinite_perspective(math.radians(45.0), width / height, 0.1)
    return Scene(f"instanced_{num_instances}x{patch}", [patch_mesh, light.mesh()], mats, inst, view, proj, width, height, (0.3, 0.4, 0.6))


# ----------------------------------------------------------------------------------------------
# ray sets
# ----------------------------------------------------------------------------------------------
def scene_bounds(scene: Scene):
    los, his = [], []
    if scene.instances is None:
        for m in scene.meshes:
            los.append(m.positions.min(0)); his.append(m.positions.max(0))
    else:
        for rec in scene.instances[: min(len(scene.instances), 20000)]:
            m = scene.meshes[int(rec["mesh_id"])]
            T = rec["transform"].reshape(3, 4).astype(np.float64)
            lo, hi = m.positions.min(0), m.positions.max(0)
            corners = np.asarray([[(hi if (k >> a) & 1 else lo)[a] for a in range(3)] for k in range(8)], np.float64)
            w = corners @ T[:, :3].T + T[:, 3]
            los.append(w.min(0)); his.append(w.max(0))
    return np.min(los, 0).astype(np.float64), np.max(his, 0).astype(np.float64)


def incoherent_rays(lo, hi, count: int, seed: int = 4, inflate: float = 0.1) -> np.ndarray:
    """origins uniform in the AABB inflated by 10 %, directions uniform on the sphere,
    tmin = 0, tmax = +inf."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo = np.asarray(lo, np.float64); hi = np.asarray(hi, np.float64)
    c = (lo + hi) / 2; half = (hi - lo) / 2 * (1.0 + inflate)
    rays = np.empty(count, RAY_DTYPE)
    chunk = 1 << 22
    for b in range(0, count, chunk):
        e = min(count, b + chunk); n = e - b
        rays["origin"][b:e] = (c + (rng.random((n, 3)) * 2 - 1) * half).astype(np.float32)
        z = rng.random(n) * 2 - 1; phi = rng.random(n) * (2 * math.pi); r = np.sqrt(np.maximum(0.0, 1 - z * z))
        rays["direction"][b:e] = np.stack([r * np.cos(phi), r * np.sin(phi), z], 1).astype(np.float32)
    rays["tmin"] = 0.0
    rays["tmax"] = np.inf
    return rays


def camera_rays(scene: Scene, count: Optional[int] = None, seed: int = 7) -> np.ndarray:
    """Coherent primary rays through random pixels (numpy restatement of the unprojection; test helper)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    W, H = scene.width, scene.height
    n = count or W * H
    px = rng.random(n) * W; py = rng.random(n) * H
    V = scene.view.astype(np.float64).reshape(4, 4).T; P = scene.proj.astype(np.float64).reshape(4, 4).T
    inv = np.linalg.inv(P @ V); eye = np.linalg.inv(V)[:3, 3]
    ndc = np.stack([2 * px / W - 1, 2 * py / H - 1, np.zeros(n), np.ones(n)], 1)
    wpt = ndc @ inv.T; wpt = wpt[:, :3] / wpt[:, 3:4]
    d = wpt - eye; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.empty(n, RAY_DTYPE)
    rays["origin"] = eye.astype(np.float32); rays["direction"] = d.astype(np.float32); rays["tmin"] = 0.0; rays["tmax"] = np.inf
    return rays


def secondary_rays(scene: Scene, rays: np.ndarray, hits: np.ndarray, seed: int = 12) -> np.ndarray:
    """Bounce rays of a FLAT scene: for every ray of `rays` that hit (per `hits`, HIT_DTYPE), a ray that starts at the hit point (offset
    along the geometric normal on the side the ray came from) with a cosine-weighted direction about that normal — the distribution the
    Lambert lobe of the render produces at bounce 1.  Used by bench.py for the `secondary` ray set (surface-started, incoherent)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = scene.meshes[0]
    ok = np.nonzero(hits["prim"] != 0xFFFFFFFF)[0]
    out = np.empty(len(ok), RAY_DTYPE)
    lo, hi = scene_bounds(scene)
    eps = float(np.max(hi - lo)) * 2.0 ** -16
    chunk = 1 << 21
    for b in range(0, len(ok), chunk):
        k = ok[b:b + chunk]
        tri = m.indices[hits["prim"][k]]
        v0 = m.positions[tri[:, 0]].astype(np.float64); e1 = m.positions[tri[:, 1]] - v0; e2 = m.positions[tri[:, 2]] - v0
        n = np.cross(e1, e2); n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        d_in = rays["direction"][k].astype(np.float64)
        n *= np.where((n * d_in).sum(1, keepdims=True) > 0, -1.0, 1.0)
        p = rays["origin"][k].astype(np.float64) + d_in * hits["t"][k].astype(np.float64)[:, None]
        u1, u2 = rng.random(len(k)), rng.random(len(k))
        r, phi = np.sqrt(u1), 2 * math.pi * u2
        a = np.where(np.abs(n[:, :1]) > 0.9, np.asarray([[0.0, 1.0, 0.0]]), np.asarray([[1.0, 0.0, 0.0]]))
        t = np.cross(a, n); t /= np.linalg.norm(t, axis=1, keepdims=True)
        bt = np.cross(n, t)
        d = t * (r * np.cos(phi))[:, None] + bt * (r * np.sin(phi))[:, None] + n * np.sqrt(np.maximum(0.0, 1 - u1))[:, None]
        out["origin"][b:b + chunk] = (p + eps * n).astype(np.float32)
        out["direction"][b:b + chunk] = d.astype(np.float32)
    out["tmin"] = 0.0
    out["tmax"] = np.inf
    return out


def stress_rays(scene: Scene, count: int, seed: int = 11) -> np.ndarray:
    """Rays aimed exactly at vertices, edge midpoints and centroids of random triangles of mesh 0 of a FLAT scene —
    the tie cases (shared edges / vertices) that must resolve by primitive index."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = scene.meshes[0]
    tri = m.indices[rng.integers(0, m.indices.shape[0], count)]
    p = m.positions[tri].astype(np.float64)                # (n,3,3)
    kind = rng.integers(0, 3, count)
    w = np.zeros((count, 3))
    vsel = rng.integers(0, 3, count)
    w[np.arange(count), vsel] = 1.0                                        # vertex
    e = kind == 1; w[e] = 0.5; w[e, vsel[e]] = 0.0                         # edge midpoint
    c = kind == 2; w[c] = 1.0 / 3.0                                        # centroid
    target = (p * w[:, :, None]).sum(1)
    lo, hi = scene_bounds(scene)
    origin = ((lo + hi) / 2 + (rng.random((count, 3)) * 2 - 1) * (hi - lo) * 0.7)
    rays = np.empty(count, RAY_DTYPE)
    rays["origin"] = origin.astype(np.float32)
    rays["direction"] = (target.astype(np.float32) - rays["origin"])       # unnormalised on purpose: hit near t = 1
    rays["tmin"] = 0.0; rays["tmax"] = np.inf
    return rays


def save_scene(scene: Scene, path: str) -> None:
    """Writes the "FPTS" container read by the C++ host (foundation_b200/renderer/Renderer.cpp SceneDesc::Load):
    header 8 x u32 {magic 'FPTS', version 1, width, height, meshes, materials, instances, 0}, view, proj, background, materials
    (8 x f32 each), instances (64 B each), then per mesh {nverts, ntris, positions f32x3, indices u32x3, material ids u32}."""
    inst = scene.instances if scene.instances is not None else np.zeros(0, INSTANCE_DTYPE)
    with open(path, "wb") as f:
        f.write(np.asarray([0x53545046, 1, scene.width, scene.height, len(scene.meshes), scene.materials.shape[0], len(inst), 0], np.uint32).tobytes())
        f.write(np.ascontiguousarray(scene.view, np.float32).tobytes()); f.write(np.ascontiguousarray(scene.proj, np.float32).tobytes())
        f.write(np.asarray(scene.background, np.float32).tobytes())
        f.write(np.ascontiguousarray(scene.materials, np.float32).tobytes()); f.write(np.ascontiguousarray(inst).tobytes())
        for m in scene.meshes:
            f.write(np.asarray([m.positions.shape[0], m.indices.shape[0]], np.uint32).tobytes())
            f.write(np.ascontiguousarray(m.positions, np.float32).tobytes()); f.write(np.ascontiguousarray(m.indices, np.uint32).tobytes())
            f.write(np.ascontiguousarray(m.material_ids, np.uint32).tobytes())


def checker_texture(size: int = 64, cells: int = 8, seed: int = 9) -> np.ndarray:
    """Procedural RGBA8 stand-in for the JPEG the reference downloads at configure time (src/Renderer/CMakeLists.txt:109-115, not
    available offline): coloured checkerboard with per-cell noise so that bilinear filtering and the REPEAT seam are both exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cell = np.add.outer(np.arange(size) // (size // cells), np.arange(size) // (size // cells)) % 2
    base = np.where(cell[..., None] == 0, np.asarray([230, 220, 200]), np.asarray([40, 60, 110]))
    img = np.clip(base + rng.integers(-25, 26, (size, size, 3)), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(np.concatenate([img, np.full((size, size, 1), 255, np.uint8)], -1))


def reference_quad(width: int = 320, height: int = 180) -> Scene:
    """The scene the reference actually draws: its 4-vertex quad (positions, colours and texture coordinates of
    src/Renderer/Renderer.cpp:153-157, indices 0 1 2 2 3 0 of :175), its camera (Renderer.cpp:373-380) and a texture bound to the quad's
    material, path traced under a constant white environment instead of rasterised: the primary hit's base colour is texel x vertex colour,
    the reference's fragment shader (src/Renderer/Triangle.slang:34-37)."""
    pos = np.asarray([(-0.5, -0.5, 0), (0.5, -0.5, 0), (0.5, 0.5, 0), (-0.5, 0.5, 0)], np.float32)
    col = np.asarray([(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)], np.float32)
    uv = np.asarray([(1, 0), (0, 0), (0, 1), (1, 1)], np.float32)
    idx = np.asarray([(0, 1, 2), (2, 3, 0)], np.uint32)
    mesh = Mesh(pos, idx, np.zeros(2, np.uint32), uv, col)
    mats = np.asarray([_mat((1.0, 1.0, 1.0), 1.0)], np.float32)
    view, proj = reference_camera(width / height)
    return Scene("reference_quad", [mesh], mats, None, view, proj, width, height, (1.0, 1.0, 1.0), [checker_texture()], np.asarray([0], np.uint32))


def by_name(name: str, **kw) -> Scene:
    return {"cornell": cornell_box, "spheres": sphere_field, "terrain": fractal_terrain, "instanced": instanced_patches}[name](**kw)

"""CPU tier: pins the oracle (known-answer vectors, exhaustive ground truth, independent float64 check, golden
fixtures) and checks the product's per-item device bodies, executed by the CPU emulation harness, against it."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from foundation_b200 import scenes
from oracle import pt_oracle as orc
from oracle.pt_oracle import HIT_DTYPE, NODE_DTYPE, OracleScene
from tests.util import SMALL_SCENES, assert_hits_equal, multi_mesh_scene, ray_mix

HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------- known answers
def test_pcg32_known_answer_vector():
    """O'Neill's pcg32 demo (pcg-c-basic, seed 42 / stream 54): the published first six outputs."""
    want = [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
    assert list(orc.pcg_raw(42, 54, 6)) == want


def test_pcg_streams_are_distinct_per_pixel_and_sample():
    a = orc.pcg(1, 10, 0, 8); b = orc.pcg(1, 11, 0, 8); c = orc.pcg(1, 10, 1, 8); d = orc.pcg(2, 10, 0, 8)
    assert len({a.tobytes(), b.tobytes(), c.tobytes(), d.tobytes()}) == 4
    assert np.array_equal(a, orc.pcg(1, 10, 0, 8))


def test_morton63_is_bit_interleave():
    rng = np.random.default_rng(0)
    lo = np.zeros(3, np.float32); inv = np.full(3, 2097152.0, np.float32)       # unit cube -> 21-bit grid
    L = orc.lib()
    for _ in range(200):
        c = rng.random(3).astype(np.float32)
        q = np.minimum((c * np.float32(2097152.0)).astype(np.uint64), 2097151)
        want = 0
        for bit in range(21):
            for axis in range(3):
                want |= ((int(q[axis]) >> bit) & 1) << (3 * bit + (2 - axis))
        got = L.orc_morton(c.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p))
        assert got == want


def test_sincos2pi_accuracy():
    u = np.linspace(0, 1, 4097, endpoint=False, dtype=np.float32)
    sc = np.asarray([orc.sincos2pi(float(x)) for x in u])
    assert np.abs(sc[:, 0] - np.sin(2 * np.pi * u.astype(np.float64))).max() < 4e-7
    assert np.abs(sc[:, 1] - np.cos(2 * np.pi * u.astype(np.float64))).max() < 4e-7


# ---------------------------------------------------------------------------------------------- scenes
def test_scene_triangle_counts_match_baseline_configs():
    assert scenes.cornell_box().num_triangles == 32
    assert scenes.sphere_field(num_spheres=3, subdiv=3).num_triangles == 3 * 1280 + 4
    assert scenes.fractal_terrain(n=50).num_triangles == 2 * 50 * 50 + 2
    assert 2 * 2236 ** 2 + 2 == 9_999_394                                     # config 3 at full size
    s = scenes.instanced_patches(num_instances=20, patch=71)
    assert s.meshes[0].indices.shape[0] == 10082 and s.effective_triangles == 20 * 10082 + 2


def test_reference_camera_conventions():
    """Renderer::Draw's camera (Renderer.cpp:373-380): eye (2,2,2), +Z up, Y flipped: the centre ray points at the origin,
    pixel row 0 is the top of the image."""
    view, proj = scenes.reference_camera()
    o = OracleScene(scenes.cornell_box()); o.camera_set(view, proj)
    cam = o.camera_get()
    eye, d0, dx, dy = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    assert np.allclose(eye, [2, 2, 2], atol=1e-5)
    c = d0 / np.linalg.norm(d0)
    assert np.allclose(c, -np.ones(3) / np.sqrt(3), atol=1e-5)
    assert dy[2] < 0                      # NDC +y (down the image) moves the ray towards -Z
    assert abs(np.dot(dx, dy)) < 1e-5 and abs(np.linalg.norm(dx) / np.linalg.norm(dy) - 1920 / 1080) < 1e-4


# ---------------------------------------------------------------------------------------------- traversal ground truth
@pytest.fixture(scope="module", params=list(SMALL_SCENES))
def small(request):
    sc = SMALL_SCENES[request.param]()
    return request.param, sc, OracleScene(sc)


def test_bvh_traversal_equals_exhaustive_search(small):
    name, sc, o = small
    rays = ray_mix(sc, 2048)
    h, i = o.trace_closest(rays)
    hb, ib = o.trace_closest(rays, brute=True)
    assert_hits_equal(h, i, hb, ib, name)
    assert h.tobytes() == hb.tobytes()
    assert np.array_equal(o.trace_any(rays), o.trace_any(rays, brute=True))
    assert (h["prim"] != 0xFFFFFFFF).mean() > 0.05


def test_hits_agree_with_independent_float64_intersection(small):
    """Independent of every shared header: numpy float64 Moller-Trumbore of the reported triangle."""
    name, sc, o = small
    if sc.instances is not None:
        pytest.skip("flat scenes only")
    rays = ray_mix(sc, 1024)
    h, _ = o.trace_closest(rays)
    m = sc.meshes[0]
    hit = h["prim"] != 0xFFFFFFFF
    tri = m.positions[m.indices[h["prim"][hit]]].astype(np.float64)
    ro = rays["origin"][hit].astype(np.float64); rd = rays["direction"][hit].astype(np.float64)
    e1 = tri[:, 1] - tri[:, 0]; e2 = tri[:, 2] - tri[:, 0]
    p = np.cross(rd, e2); det = (e1 * p).sum(1)
    tv = ro - tri[:, 0]; q = np.cross(tv, e1)
    t = (e2 * q).sum(1) / det; u = (tv * p).sum(1) / det; v = (rd * q).sum(1) / det
    ok = np.abs(det) > 1e-9 * np.linalg.norm(e1, axis=1) * np.linalg.norm(e2, axis=1) * np.linalg.norm(rd, axis=1)
    assert np.allclose(t[ok], h["t"][hit][ok], rtol=2e-3, atol=1e-4)
    assert (u[ok] > -1e-3).all() and (v[ok] > -1e-3).all() and (u[ok] + v[ok] < 1 + 1e-3).all()
    assert np.allclose(u[ok], h["u"][hit][ok], atol=2e-3) and np.allclose(v[ok], h["v"][hit][ok], atol=2e-3)


def test_ties_resolve_to_lowest_primitive_index():
    """Two coincident triangles and a shared edge: the reported id is the smaller one whatever the BVH does."""
    pos = np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    idx = np.asarray([[1, 3, 2], [0, 1, 2], [0, 1, 2], [0, 1, 2]], np.uint32)      # prims 1,2,3 coincide; 0 shares the diagonal
    sc = scenes.Scene("ties", [scenes.Mesh(pos, idx, np.zeros(4, np.uint32))], np.asarray([[.8, .8, .8, .5, 0, 0, 0, 0]], np.float32))
    o = OracleScene(sc)
    r = np.zeros(2, scenes.RAY_DTYPE); r["tmax"] = np.inf
    r["origin"] = [(0.25, 0.25, 1), (0.5, 0.5, 1)]; r["direction"] = [(0, 0, -1), (0, 0, -1)]
    for brute in (False, True):
        h, _ = o.trace_closest(r, brute=brute)
        assert list(h["prim"]) == [1, 0] and list(h["t"]) == [1.0, 1.0]


# ---------------------------------------------------------------------------------------------- product bodies under emulation
def _emu():
    so = os.path.join(HERE, "emu", "libpt_emu.so")
    subprocess.run(["make", "-C", os.path.join(HERE, "emu")], check=True, capture_output=True)
    E = C.CDLL(so)
    E.emu_build.restype = C.c_uint32
    E.emu_build_agglomerative.restype = C.c_uint32
    return E


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def test_product_build_and_traversal_bodies_match_oracle(small):
    """pt_build.h (Karras emit, collapse, quantisation) and pt_traverse.h (group-stack traversal) — the code the CUDA
    kernels run per work item — executed on the CPU: byte-identical BVH8, bit-identical hits, identical visit counters."""
    name, sc, _ = small
    E = _emu()
    for leaf in (1, 3):
        _check_product_bodies(name, sc, OracleScene(sc, max_leaf=leaf), E, leaf)


def _check_product_bodies(name, sc, o, E, leaf):
    flat_nodes, flat_tris = [], []
    for mid, mesh in enumerate(sc.meshes):
        v = mesh.positions[mesh.indices].astype(np.float32)
        box = np.ascontiguousarray(np.concatenate([v.min(1), v.max(1)], 1), np.float32)
        cent = np.ascontiguousarray(((v[:, 0] + v[:, 1]) + v[:, 2]) * np.float32(0.333333343267440796), np.float32)
        n = len(v)
        nodes = np.zeros(n + 1, NODE_DTYPE); seq = np.zeros(n, np.uint32); order = np.zeros(n, np.uint32)
        lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32)
        nn = E.emu_build(_p(box), _p(cent), C.c_uint32(n), C.c_uint32(leaf), _p(lo), _p(hi), _p(nodes), _p(seq), _p(order))
        on, ot, oo = o.blas(mid)
        assert nn == len(on) and nodes[:nn].tobytes() == on.tobytes(), f"{name} mesh {mid}: nodes differ"
        assert np.array_equal(order, oo) and np.array_equal(ot["prim"], order[seq])
        # the product's default build: bottom-up agglomerative joins (pt_join / pt_join_is_local), emulated tile by tile like the kernels,
        # for several tile sizes (tiny tiles push almost every node through the "global" path, one big tile keeps everything local)
        for tile in (2, 5, 64, 256, 1 << 30):
            n2 = np.zeros(n + 1, NODE_DTYPE); s2 = np.zeros(n, np.uint32); o2 = np.zeros(n, np.uint32)
            m2 = E.emu_build_agglomerative(_p(box), _p(cent), C.c_uint32(n), C.c_uint32(leaf), _p(lo), _p(hi), _p(n2), _p(s2), _p(o2), C.c_uint32(tile))
            assert m2 == nn and n2[:m2].tobytes() == on.tobytes() and np.array_equal(s2, seq) and np.array_equal(o2, order), f"{name} mesh {mid}: agglomerative build, tile {tile}"
        flat_nodes.append(on); flat_tris.append(ot)
    rays = ray_mix(sc, 2048)
    if sc.instances is None:
        nodes, tris, inst, two = flat_nodes[0], flat_tris[0], None, 0
    else:
        tn, _, rec = o.tlas()
        nodes = np.concatenate([tn] + flat_nodes); tris = np.concatenate(flat_tris); inst = rec; two = 1
    h, i, cnt = o.trace_closest(rays, counters=True, threads=1)
    hits = np.zeros(len(rays), HIT_DTYPE); ins = np.zeros(len(rays), np.uint32); c2 = np.zeros(3, np.uint64)
    ok = E.emu_trace(_p(nodes), _p(tris), _p(inst), two, _p(rays), C.c_uint64(len(rays)), _p(hits), _p(ins), None, 0, _p(c2))
    assert ok == 1 and hits.tobytes() == h.tobytes() and np.array_equal(ins, i)
    assert np.array_equal(cnt, c2), f"{name}: visit counters differ {cnt} vs {c2}"
    occ = np.zeros(len(rays), np.uint8)
    E.emu_trace(_p(nodes), _p(tris), _p(inst), two, _p(rays), C.c_uint64(len(rays)), None, None, _p(occ), 1, None)
    assert np.array_equal(occ, o.trace_any(rays))


# ---------------------------------------------------------------------------------------------- shading model self-checks
MATS = [[0.8, 0.8, 0.8, 1.0, 0, 0, 0, 0.0], [1.0, 1.0, 1.0, 0.3, 0, 0, 0, 0.0], [1.0, 1.0, 1.0, 0.5, 0, 0, 0, 1.0], [0.9, 0.6, 0.3, 0.15, 0, 0, 0, 1.0]]


@pytest.mark.parametrize("mat", MATS)
def test_bsdf_reciprocity_energy_and_pdf(mat):
    rng = np.random.default_rng(5)

    def hemi(n):
        z = rng.random(n); phi = rng.random(n) * 2 * np.pi; r = np.sqrt(1 - z * z)
        return np.stack([r * np.cos(phi), r * np.sin(phi), z], 1).astype(np.float32)

    for wo in hemi(6):
        if wo[2] < 0.1:
            continue
        wis = hemi(4000)
        pdf_int = 0.0
        for k, wi in enumerate(wis):
            f1, p1 = orc.bsdf_eval(mat, wo, wi)
            if k < 300:
                f2, _ = orc.bsdf_eval(mat, wi, wo)
                assert np.allclose(f1, f2, rtol=2e-4, atol=1e-6)              # Helmholtz reciprocity
            pdf_int += p1 * 2 * np.pi / len(wis)
        # the mixture pdf integrates to <= 1 over the upper hemisphere (VNDF samples below the horizon are rejected);
        # uniform-hemisphere Monte Carlo of a peaked lobe is noisy, so this only catches gross normalisation errors
        if mat[3] >= 0.3:                                                     # sharper lobes need far more uniform samples than is worth it here
            assert 0.55 < pdf_int < 1.2, pdf_int
        est = []
        for _ in range(3000):                                                  # albedo via the model's own importance sampling
            ok, wi = orc.bsdf_sample(mat, wo, float(rng.random()), float(rng.random()), float(rng.random()))
            if not ok:
                est.append(0.0); continue
            f, p = orc.bsdf_eval(mat, wo, wi)
            est.append(float(f.max()) * wi[2] / p if p > 0 else 0.0)
        assert np.mean(est) <= 1.06, f"energy gain: albedo {np.mean(est)}"   # 3000-sample MC estimate, sigma ~0.02


def test_nee_and_bsdf_sampling_estimators_agree():
    """Same expectation from light sampling only and BSDF sampling only (Cornell, low resolution, many samples)."""
    sc = scenes.cornell_box(24, 24)
    o = OracleScene(sc)
    spp = 384
    nee = o.render(24, 24, 11, 0, spp, 3, flags=4)[..., :3].mean() / spp      # NO_BSDF_EMISSION
    bsdf = o.render(24, 24, 12, 0, spp, 3, flags=2)[..., :3].mean() / spp     # NO_NEE
    mis = o.render(24, 24, 13, 0, spp, 3, flags=0)[..., :3].mean() / spp
    assert abs(nee - mis) / mis < 0.06 and abs(bsdf - mis) / mis < 0.12, (nee, bsdf, mis)


def test_render_is_deterministic_and_progressive():
    sc = scenes.cornell_box(32, 32)
    o = OracleScene(sc)
    a = o.render(32, 32, 5, 0, 3, 4)
    b = o.render(32, 32, 5, 0, 1, 4); b = o.render(32, 32, 5, 1, 2, 4, accum=b)
    assert np.array_equal(a, b)
    assert np.array_equal(a, o.render(32, 32, 5, 0, 3, 4, threads=1))
    brute = o.render(32, 32, 5, 0, 3, 4, brute=True)
    assert np.array_equal(a, brute)                                           # the BVH never changes a path


# ---------------------------------------------------------------------------------------------- golden fixtures
def test_golden_fixtures():
    """Fixtures committed under tests/golden (generated by tests/golden/make_golden.py from this oracle): a regression pin
    for the arithmetic contract — the GPU tests compare against the same files."""
    g = json.load(open(os.path.join(HERE, "golden", "golden.json")))
    for name, want in g["scenes"].items():
        sc = SMALL_SCENES[name]()
        o = OracleScene(sc)
        nodes, tris, order = o.blas(0)
        assert hashlib.sha256(nodes.tobytes()).hexdigest() == want["blas0_nodes_sha256"], name
        rays = ray_mix(sc, 512)
        h, i = o.trace_closest(rays)
        assert hashlib.sha256(h.tobytes() + i.tobytes()).hexdigest() == want["hits_sha256"], name
        img = o.render(sc.width, sc.height, 7, 0, 1, 3, background=sc.background)
        assert hashlib.sha256(img.tobytes()).hexdigest() == want["image_sha256"], name
    cornell = np.load(os.path.join(HERE, "golden", "cornell_hits.npz"))
    sc = SMALL_SCENES["cornell"](); o = OracleScene(sc)
    h, _ = o.trace_closest(cornell["rays"].view(scenes.RAY_DTYPE).reshape(-1))
    assert np.array_equal(h["prim"], cornell["prim"]) and np.array_equal(h["t"].view(np.uint32), cornell["t_bits"])


def test_far_ray_origins_stay_conservative():
    """Rays starting ~100 scene radii away (instanced BLAS see them in object space, where the mesh is tiny compared with the
    distance): the ray-dependent slab slack (PT_SLAB_EPS) keeps BVH culling consistent with exhaustive search."""
    sc = scenes.instanced_patches(num_instances=60, patch=8, width=64, height=64)
    lo, hi = scenes.scene_bounds(sc)
    rng = np.random.default_rng(5)
    n = 30000
    c = (lo + hi) / 2; half = (hi - lo) / 2
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    org = c + dirs * np.linalg.norm(half) * 100.0
    tgt = c + (rng.random((n, 3)) * 2 - 1) * half
    rays = np.zeros(n, scenes.RAY_DTYPE)
    rays["origin"] = org.astype(np.float32); d = tgt - org; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32); rays["tmax"] = np.inf
    for leaf in (1, 3):
        o = OracleScene(sc, max_leaf=leaf)
        h, i = o.trace_closest(rays); hb, ib = o.trace_closest(rays, brute=True)
        assert_hits_equal(h, i, hb, ib, f"far origins, leaf {leaf}")


def test_several_meshes_without_instances_are_instanced_once_each():
    """Implicit identity instances: the split Cornell box traces and renders like the single-mesh one (prim ids are per mesh)."""
    sc3 = multi_mesh_scene(); sc1 = scenes.cornell_box(96, 96)
    o3 = OracleScene(sc3); o1 = OracleScene(sc1)
    rays = ray_mix(sc1, 2048)
    h3, i3 = o3.trace_closest(rays); h1, _ = o1.trace_closest(rays); hb, ib = o3.trace_closest(rays, brute=True)
    assert_hits_equal(h3, i3, hb, ib, "3 meshes: BVH vs exhaustive")
    base = np.asarray([0, 10, 20])
    hit = h3["prim"] != 0xFFFFFFFF
    assert np.array_equal(hit, h1["prim"] != 0xFFFFFFFF) and np.array_equal(h3["t"], h1["t"])
    assert np.array_equal(base[i3[hit]] + h3["prim"][hit], h1["prim"][hit])


def degenerate_scene():
    """Zero-area triangles (repeated vertex, collinear vertices), duplicated triangles and a far outlier: ragged input the build must
    survive without NaNs in the tree; degenerate triangles can never be hit (|det| = 0 is rejected)."""
    rng = np.random.default_rng(2)
    pos = rng.random((60, 3)).astype(np.float32)
    idx = rng.integers(0, 60, (200, 3)).astype(np.uint32)
    idx[::7, 1] = idx[::7, 0]                              # repeated vertex
    pos[50] = pos[51] * 0.5 + pos[52] * 0.5; idx[5] = (51, 50, 52)    # collinear
    idx[100:110] = idx[90:100]                             # duplicates
    pos[59] = (40.0, -30.0, 25.0)                          # outlier stretches the Morton grid
    return scenes.Scene("degenerate", [scenes.Mesh(pos, idx, np.zeros(200, np.uint32))], np.asarray([[.8, .8, .8, .5, 0, 0, 0, 0]], np.float32))


def test_degenerate_and_duplicate_triangles():
    sc = degenerate_scene()
    o = OracleScene(sc)
    nodes, tris, order = o.blas(0)
    assert np.isfinite(nodes["p"]).all() and sorted(order.tolist()) == list(range(200))
    lo, hi = scenes.scene_bounds(sc)
    rays = np.concatenate([scenes.incoherent_rays(lo, hi, 6000, 3), scenes.incoherent_rays(np.zeros(3), np.ones(3), 6000, 4)])
    h, i = o.trace_closest(rays); hb, ib = o.trace_closest(rays, brute=True)
    assert_hits_equal(h, i, hb, ib, "degenerate mesh")
    m = sc.meshes[0]
    hit = h["prim"][h["prim"] != 0xFFFFFFFF]
    tri = m.positions[m.indices[hit]].astype(np.float64)
    area = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    assert (area > 0).all() and len(hit) > 100
    dup_hit = hit[(hit >= 100) & (hit < 110)]
    assert len(dup_hit) == 0, "a duplicated triangle must lose the tie to its lower-index twin"


def test_product_bodies_on_degenerate_and_chain_like_inputs():
    """Equal Morton keys (duplicated / coincident triangles: the radix tree is decided by the index tie-break) through the emulated product
    build, Karras and agglomerative, at several tile sizes."""
    E = _emu()
    rng = np.random.default_rng(5)
    tri = np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    pos = np.concatenate([np.tile(tri, (300, 1)), (rng.uniform(-50, 50, (7, 1, 3)) + tri[None]).reshape(-1, 3).astype(np.float32)])
    idx = np.arange(pos.shape[0], dtype=np.uint32).reshape(-1, 3)
    chain = scenes.Scene("chain", [scenes.Mesh(np.ascontiguousarray(pos), idx, np.zeros(idx.shape[0], np.uint32))], np.asarray([[.8, .8, .8, .5, 0, 0, 0, 0]], np.float32))
    for sc in (degenerate_scene(), chain):
        sc.view = None
        o = OracleScene(sc)
        for mid, mesh in enumerate(sc.meshes):
            v = mesh.positions[mesh.indices].astype(np.float32)
            box = np.ascontiguousarray(np.concatenate([v.min(1), v.max(1)], 1), np.float32)
            cent = np.ascontiguousarray(((v[:, 0] + v[:, 1]) + v[:, 2]) * np.float32(0.333333343267440796), np.float32)
            n = len(v)
            on, ot, oo = o.blas(mid)
            lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32)
            for tile in (0, 3, 64, 256):
                nodes = np.zeros(n + 1, NODE_DTYPE); seq = np.zeros(n, np.uint32); order = np.zeros(n, np.uint32)
                if tile == 0:
                    nn = E.emu_build(_p(box), _p(cent), C.c_uint32(n), C.c_uint32(1), _p(lo), _p(hi), _p(nodes), _p(seq), _p(order))
                else:
                    nn = E.emu_build_agglomerative(_p(box), _p(cent), C.c_uint32(n), C.c_uint32(1), _p(lo), _p(hi), _p(nodes), _p(seq), _p(order), C.c_uint32(tile))
                assert nn == len(on) and nodes[:nn].tobytes() == on.tobytes() and np.array_equal(order, oo), (sc.name, tile)


def test_bsdf_matches_independent_float64_formula():
    """The surface model restated in numpy float64 from its definition (Lambert base under (1-F)(1-F), GGX D, height-correlated
    Smith G2, Schlick F; VNDF pdf) — independent of pt_shading.h — agrees with the shared float32 arithmetic."""
    rng = np.random.default_rng(9)

    def ref(mat, wo, wi):
        base = np.asarray(mat[:3], np.float64); rough, metal = float(mat[3]), float(mat[7])
        kd = base * (1 - metal); f0 = 0.04 + (base - 0.04) * metal
        alpha = max(rough * rough, 1e-3); a2 = alpha * alpha
        h = (wo + wi) / np.linalg.norm(wo + wi)
        D = a2 / (np.pi * ((h[2] ** 2) * (a2 - 1) + 1) ** 2)
        lam = lambda c: 0.5 * (np.sqrt(1 + a2 * (1 - c * c) / (c * c)) - 1)
        G2 = 1 / (1 + lam(wo[2]) + lam(wi[2])); G1 = 1 / (1 + lam(wo[2]))
        F = f0 + (1 - f0) * (1 - max(wo @ h, 0)) ** 5
        Fo = f0 + (1 - f0) * (1 - wo[2]) ** 5; Fi = f0 + (1 - f0) * (1 - wi[2]) ** 5
        f = kd / np.pi * (1 - Fo) * (1 - Fi) + F * D * G2 / (4 * wo[2] * wi[2])
        p_spec = 0.5 + 0.5 * metal
        pdf = p_spec * G1 * D / (4 * wo[2]) + (1 - p_spec) * wi[2] / np.pi
        return f, pdf

    for mat in MATS + [[0.3, 0.6, 0.9, 0.7, 0, 0, 0, 0.5]]:
        for _ in range(200):
            v = rng.normal(size=(2, 3)); v[:, 2] = np.abs(v[:, 2]) + 0.05; v /= np.linalg.norm(v, axis=1, keepdims=True)
            f, pdf = orc.bsdf_eval(mat, v[0].astype(np.float32), v[1].astype(np.float32))
            fr, pr = ref(mat, v[0].astype(np.float32).astype(np.float64), v[1].astype(np.float32).astype(np.float64))
            assert np.allclose(f, fr, rtol=2e-3, atol=1e-6) and abs(pdf - pr) <= 2e-3 * max(pr, 1e-3), (mat, f, fr, pdf, pr)


def test_sobol02_is_a_02_sequence_and_known_values():
    """Known first points of the (0,2)-sequence (van der Corput base 2; second Sobol dimension 0, .5, .25, .75, .125 ...  and
    0, .5, .75, .25, .625 ...), and the defining stratification: every aligned block of 2^k samples hits each elementary interval
    of area 2^-k exactly once, with and without XOR scrambling."""
    pts = np.asarray([orc.sobol02(i) for i in range(8)])
    assert np.allclose(pts[:, 0], [0, .5, .25, .75, .125, .625, .375, .875])
    assert np.allclose(pts[:, 1], [0, .5, .75, .25, .625, .125, .375, .875])
    for k0, k1 in ((0, 0), (0x9E3779B9, 0x7F4A7C15)):
        for k in (4, 6):
            n = 1 << k
            for start in (0, n, 5 * n):
                p = np.asarray([orc.sobol02(start + i, k0, k1) for i in range(n)])
                for a in range(k + 1):                       # grid 2^a x 2^(k-a)
                    cells = np.floor(p[:, 0] * (1 << a)).astype(int) * (1 << (k - a)) + np.floor(p[:, 1] * (1 << (k - a))).astype(int)
                    assert len(set(cells.tolist())) == n, (k0, k, start, a)


def test_sobol_jitter_lowers_pixel_variance_and_keeps_the_mean():
    sc = scenes.cornell_box(24, 24)
    o = OracleScene(sc)
    spp = 64
    a = o.render(24, 24, 3, 0, spp, 2, flags=0)[..., :3] / spp
    b = o.render(24, 24, 3, 0, spp, 2, flags=16)[..., :3] / spp
    assert abs(a.mean() - b.mean()) / a.mean() < 0.05 and not np.array_equal(a, b)


def test_padded_sobol_path_dimensions_are_nets_and_decorrelated():
    """PT_FLAG_SOBOL_PATH: every (pixel, bounce, decision) key gives a sequence whose aligned 2^k-sample blocks are (0,2)-nets (index
    shuffle + Owen scrambling of both coordinates are nested permutations), and two decisions of the same vertex fill the joint
    domain like independent samples (padding), not like one sequence XOR-ed with itself."""
    keys = (0x123456789abcdef, orc.path_dim_key(7, 100, 0, 0), orc.path_dim_key(7, 100, 3, 1))
    assert len(set(keys)) == 3
    for key in keys:
        for k in (3, 6, 8):
            n = 1 << k
            for start in (0, 3 * n):
                p = np.asarray([orc.sobol02_padded(start + i, key) for i in range(n)])
                assert p.min() >= 0.0 and p.max() < 1.0
                for a in range(k + 1):
                    cells = np.floor(p[:, 0] * (1 << a)).astype(int) * (1 << (k - a)) + np.floor(p[:, 1] * (1 << (k - a))).astype(int)
                    assert len(set(cells.tolist())) == n, (hex(key), k, start, a)
    a = np.asarray([orc.sobol02_padded(i, orc.path_dim_key(7, 100, 0, 0)) for i in range(1024)])
    b = np.asarray([orc.sobol02_padded(i, orc.path_dim_key(7, 100, 0, 1)) for i in range(1024)])
    assert abs(np.corrcoef(a[:, 0], b[:, 0])[0, 1]) < 0.1 and abs(np.corrcoef(a[:, 1], b[:, 1])[0, 1]) < 0.1
    joint = np.floor(a[:, 0] * 32).astype(int) * 32 + np.floor(b[:, 0] * 32).astype(int)
    assert 560 < len(set(joint.tolist())) < 740        # independent points occupy ~647 of 1024 cells; a shared sequence would occupy 32 or 1024


def test_sobol_path_dimensions_keep_the_mean_and_do_not_raise_the_error():
    sc = scenes.cornell_box(24, 24)
    o = OracleScene(sc)
    ref = o.render(24, 24, 99, 0, 1024, 3, flags=0)[..., :3] / 1024
    err = {}
    for fl in (0, 32, 48):
        e = []
        for seed in range(4):
            im = o.render(24, 24, seed, 0, 32, 3, flags=fl)[..., :3] / 32
            assert abs(im.mean() - ref.mean()) / ref.mean() < 0.08, (fl, seed)
            e.append(float(np.sqrt(np.mean((im - ref) ** 2))))
        err[fl] = float(np.mean(e))
    assert err[32] < 1.1 * err[0] and err[48] < 0.8 * err[0], err
    assert not np.array_equal(o.render(24, 24, 1, 0, 2, 3, flags=0), o.render(24, 24, 1, 0, 2, 3, flags=32))


@pytest.mark.parametrize("rough,metallic,lo,hi", [(0.1, 1.0, 0.99, 1.005), (0.5, 1.0, 0.85, 1.0), (1.0, 0.0, 0.84, 0.95), (0.5, 0.0, 0.86, 0.97)])
def test_white_furnace(rough, metallic, lo, hi):
    """spheres and a ground plane with white, emission-free materials inside a uniform environment of radiance 1,
    32 bounces, BSDF sampling + Russian roulette only (there is no light to sample).  Energy can only be lost, never gained: a polished
    white metal (F = 1, negligible single-scattering loss) must return the environment exactly; rough GGX loses what single scattering
    cannot return; the diffuse base loses the (1 - F)(1 - F) coupling.  The bounds bracket what the model gives; a wrong cosine, pdf,
    roulette weight or a self-intersecting bounce ray moves the mean far outside them."""
    sc = scenes.sphere_field(num_spheres=12, subdiv=2, width=48, height=27)
    mats = sc.materials.copy()
    mats[:, 0:3] = 1.0; mats[:, 3] = rough; mats[:, 4:7] = 0.0; mats[:, 7] = metallic
    sc2 = scenes.Scene("furnace", sc.meshes, mats, sc.instances, sc.view, sc.proj, 48, 27, (1.0, 1.0, 1.0))
    o = OracleScene(sc2)
    spp = 64
    img = o.render(48, 27, 5, 0, spp, 32, background=(1.0, 1.0, 1.0))[..., :3] / spp
    assert lo <= img.mean() <= hi, (rough, metallic, float(img.mean()))
    assert img.max() <= 1.15, float(img.max())                     # per-pixel 64-spp noise only


def test_oracle_under_address_and_undefined_behaviour_sanitizers(tmp_path):
    """the checker itself is checked.  The oracle is rebuilt with -fsanitize=address,undefined and drives a flat scene,
    an instanced scene and the degenerate mesh through build, both traversals, exhaustive search and a render in a child process."""
    cxx = asan = None
    for cand in (os.environ.get("CXX"), "g++", "/usr/bin/g++"):          # the first compiler that ships the sanitizer runtime
        if not cand:
            continue
        try:
            p = subprocess.run([cand, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
        except OSError:
            continue
        if os.path.isabs(p) and os.path.exists(p):
            cxx, asan = cand, p
            break
    if not asan:
        pytest.skip("no libasan in this toolchain")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libpt_oracle_san.so")
    subprocess.run(["make", "-C", os.path.join(root, "oracle"), "asan", f"SAN_OUT={so}", f"CXX={cxx}"], check=True, capture_output=True)
    code = """
import numpy as np
from foundation_b200 import scenes
from oracle.pt_oracle import OracleScene
from tests.util import SMALL_SCENES, ray_mix
from tests.test_oracle import degenerate_scene
for make in (SMALL_SCENES['cornell'], SMALL_SCENES['instanced'], degenerate_scene):
    sc = make(); o = OracleScene(sc)
    lo, hi = scenes.scene_bounds(sc)
    rays = ray_mix(sc, 256) if sc.view is not None else scenes.incoherent_rays(lo, hi, 1024, 3)
    h, i = o.trace_closest(rays); b, _ = o.trace_closest(rays[:64], brute=True)
    assert (h['prim'][:64] == b['prim']).all()
    o.trace_any(rays)
    o.blas(0)
    if sc.view is not None:
        o.render(16, 16, 3, 0, 2, 3, flags=48)
print('sanitized run ok')
"""
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1", PT_ORACLE_LIB=so, PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=600)
    assert r.returncode == 0 and "sanitized run ok" in r.stdout, r.stderr[-3000:]
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]


def test_randomised_meshes_through_both_emulated_builds():
    """120 random triangle soups — uniform, snapped to a coarse grid (many equal Morton keys), duplicated triangles, extreme aspect ratio —
    with 1 .. 2000 triangles, both leaf sizes: the emulated Karras build and the emulated agglomerative build at tile sizes 2, 7, 64, 256
    reproduce the oracle's BVH8 byte for byte."""
    E = _emu()
    rng = np.random.default_rng(123)
    for trial in range(120):
        n = int(rng.choice([1, 2, 3, 4, 7, 31, 64, 255, 256, 257, 600, 2000]))
        mode = trial % 4
        if mode == 0:
            pos = rng.random((3 * n, 3)).astype(np.float32)
        elif mode == 1:
            pos = np.round(rng.random((3 * n, 3)) * 4).astype(np.float32) / 4
        elif mode == 2:
            base = rng.random((max(1, n // 8) * 3, 3)).astype(np.float32)
            pos = np.tile(base, (-(-n * 3 // len(base)), 1))[:3 * n].copy()
        else:
            pos = (rng.random((3 * n, 3)) * np.asarray([1000, 1e-3, 1])).astype(np.float32)
        idx = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
        sc = scenes.Scene("soup", [scenes.Mesh(np.ascontiguousarray(pos), idx, np.zeros(n, np.uint32))], np.asarray([[.8, .8, .8, .5, 0, 0, 0, 0]], np.float32))
        v = pos[idx]
        box = np.ascontiguousarray(np.concatenate([v.min(1), v.max(1)], 1), np.float32)
        cent = np.ascontiguousarray(((v[:, 0] + v[:, 1]) + v[:, 2]) * np.float32(0.333333343267440796), np.float32)
        lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32)
        for leaf in (1, 3):
            on, ot, oo = OracleScene(sc, max_leaf=leaf).blas(0)
            for tile in (0, 2, 7, 64, 256):
                nodes = np.zeros(n + 1, NODE_DTYPE); seq = np.zeros(n, np.uint32); order = np.zeros(n, np.uint32)
                if tile == 0:
                    nn = E.emu_build(_p(box), _p(cent), C.c_uint32(n), C.c_uint32(leaf), _p(lo), _p(hi), _p(nodes), _p(seq), _p(order))
                else:
                    nn = E.emu_build_agglomerative(_p(box), _p(cent), C.c_uint32(n), C.c_uint32(leaf), _p(lo), _p(hi), _p(nodes), _p(seq), _p(order), C.c_uint32(tile))
                assert nn == len(on) and nodes[:nn].tobytes() == on.tobytes() and np.array_equal(order, oo), (trial, n, mode, leaf, tile)


def test_sobol02_prefixes_equal_the_published_joe_kuo_sequence():
    """External pin: with zero scramble keys every 2^k-point prefix of pt_sobol02 is, as a SET, the first 2^k points of the 2-D Sobol
    sequence with Joe & Kuo's direction numbers as shipped in scipy.stats.qmc (scipy enumerates in Gray-code order, which permutes points
    only inside such prefixes)."""
    from scipy.stats import qmc
    ref = qmc.Sobol(d=2, scramble=False).random_base2(10)
    ours = np.asarray([orc.sobol02(i, 0, 0) for i in range(1024)])
    for k in range(0, 11):
        a = {tuple(np.round(p * (1 << 24)).astype(np.int64)) for p in ours[: 1 << k]}
        b = {tuple(np.round(p * (1 << 24)).astype(np.int64)) for p in ref[: 1 << k]}
        assert a == b, k


def test_direct_light_under_a_square_emitter_matches_the_analytic_form_factor():
    """External pin of the light-transport scale (pi factors, area <-> solid-angle pdf, MIS weights summing to one): radiance leaving a
    rough white floor point straight below the centre of a square diffuse emitter.  Irradiance has a closed form — the differential-area to
    parallel-rectangle configuration factor (Howell catalogue B-4): E = L_e * pi * F, F = 4 F_corner(a/2, a/2, h),
    F_corner = 1/(2 pi) [ X/sqrt(1+X^2) atan(Y/sqrt(1+X^2)) + Y/sqrt(1+Y^2) atan(X/sqrt(1+Y^2)) ], X = a/h, Y = b/h.
    The surface model is Lambert under a dielectric coat (F0 = 0.04), not pure Lambert, so the rendered value may sit a few per cent off
    albedo / pi * E; the test allows 8 % — a lost or doubled pi, a wrong pdf conversion or a broken MIS weight would be off by 2x or more."""
    a, h, Le, albedo = 1.0, 1.5, 10.0, 0.8
    X = Y = (a / 2) / h
    Fc = (X / np.sqrt(1 + X * X) * np.arctan(Y / np.sqrt(1 + X * X)) + Y / np.sqrt(1 + Y * Y) * np.arctan(X / np.sqrt(1 + Y * Y))) / (2 * np.pi)
    E = Le * np.pi * 4 * Fc
    want = albedo / np.pi * E
    b = scenes._Builder()
    b.add(*scenes._quad((-50, -50, 0), (50, -50, 0), (50, 50, 0), (-50, 50, 0)), 0)
    b.add(*scenes._quad((-a / 2, a / 2, h), (a / 2, a / 2, h), (a / 2, -a / 2, h), (-a / 2, -a / 2, h)), 1)          # faces -z
    mats = np.asarray([scenes._mat((albedo,) * 3, 1.0), scenes._mat((0, 0, 0), 1.0, (Le,) * 3)], np.float32)
    W = 16
    # a narrow camera just under the emitter's height, off to the side, looking at the floor point below the emitter's centre
    view = scenes.look_at((0.6, -0.8, 1.2), (0, 0, 0), (0, 0, 1)); proj = scenes.infinite_perspective(np.radians(0.5), 1.0, 0.1)
    sc = scenes.Scene("form_factor", [b.mesh()], mats, None, view, proj, W, W)
    o = OracleScene(sc)
    spp = 256
    for flags in (0, 2, 4):                                   # MIS, BSDF sampling only, NEE only: three estimators of the same integral
        img = o.render(W, W, 11, 0, spp, 1, flags=flags)
        got = img[..., :3].mean() / spp
        assert abs(got / want - 1) < 0.08, (flags, got, want)

"""CPU tier: an INDEPENDENT ground truth for MISSED hits.  Everything else in the suite (CPU brute force, oracle BVH, device BVH, device
exhaustive search) calls the shared `pt_ray_tri`, so a leak in that routine would be invisible to every parity check (round-1 verdict).
Here the reference is a numpy float64 exhaustive search written from scratch (no shared header).  What is asserted, with the epsilons
written out:

  * every ray whose float64 closest hit lies inside a triangle by a barycentric margin > EPS_IN is reported with the same primitive and
    t within T_REL / T_ABS, and it never ends up BEHIND that surface; every hit the oracle reports exists in float64 (edges widened by
    EPS_IN): no phantom surfaces;
  * the Moller-Trumbore form is NOT watertight, and the test says exactly how much: a ray can be lost only inside a band of barycentric
    width EPS_BAND around a shared edge / vertex (every lost ray's float64 margin is below it), and the fraction of rays AIMED at an edge
    midpoint or a vertex to the last float bit that fall between the triangles is measured and bounded (7 % on the terrain).  For rays
    that are not constructed to hit an edge the band has measure ~1e-5 of a surface crossing.
Two watertight forms were built and measured in round 2 and rejected on cost: a 3-D triple product with exactly
antisymmetric edge functions (0 edge leaks, but ill-conditioned for small far triangles: BVH and exhaustive search disagreed on 3 of 2^18
rays; -12 % Mrays/s) and a Woop-style test on vertices projected into a per-ray frame kept in shared memory (0 edge AND 0 vertex leaks in
this very file; -18 % Mrays/s).
No reference counterpart: the reference ships no intersection code (mos9527/Foundation src/Renderer/Triangle.slang:23-37 is a textured quad)."""
import numpy as np
import pytest

from foundation_b200 import scenes
from oracle.pt_oracle import OracleScene

EPS_IN = 1e-4      # barycentric margin beyond which float64 and float32 must agree on the primitive
EPS_BAND = 1e-4    # a ray can only be lost within this barycentric distance of a shared edge / vertex (small scenes: distance / triangle size < ~50)
T_REL = 2e-5       # relative tolerance on the hit distance ...
T_ABS = 1e-6       # ... plus an absolute one (t is in units of |direction|; origins a hair away from a surface give t ~ 1e-3)
MISS = 0xFFFFFFFF


def f64_exhaustive(mesh, rays, eps=1e-9, chunk=256):
    """closest hit of every ray over ALL triangles in float64; returns t, primitive (-1 = miss) and the barycentric margin of the hit
    (> 0 strictly inside, ~0 on an edge); edges are inclusive by `eps`"""
    P = mesh.positions.astype(np.float64); I = mesh.indices
    v0 = P[I[:, 0]]; e1 = P[I[:, 1]] - v0; e2 = P[I[:, 2]] - v0
    n = len(rays)
    tb = np.full(n, np.inf); pb = np.full(n, -1, np.int64); mb = np.full(n, -np.inf); fb = np.zeros(n)
    for b in range(0, n, chunk):
        o = rays["origin"][b:b + chunk].astype(np.float64)[:, None, :]; d = rays["direction"][b:b + chunk].astype(np.float64)[:, None, :]
        p = np.cross(d, e2[None]); det = (e1[None] * p).sum(-1)
        ok = det != 0
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = o - v0[None]
        u = (tv * p).sum(-1) * inv
        q = np.cross(tv, e1[None]); v = (d * q).sum(-1) * inv
        t = (e2[None] * q).sum(-1) * inv
        m = np.minimum(np.minimum(u, v), 1.0 - u - v)
        hit = ok & (m >= -eps) & (t > rays["tmin"][b:b + chunk, None]) & (t < rays["tmax"][b:b + chunk, None])
        tt = np.where(hit, t, np.inf)
        k = tt.argmin(1); r = np.arange(len(k))
        tb[b:b + chunk] = tt[r, k]; pb[b:b + chunk] = np.where(np.isfinite(tt[r, k]), k, -1); mb[b:b + chunk] = m[r, k]; fb[b:b + chunk] = np.sign(det[r, k])
    return tb, pb, mb, fb


def shared_edges(mesh):
    """(i, j, tri_a, tri_b) for every edge used by exactly two triangles (vertex indices, i < j)"""
    I = mesh.indices.astype(np.int64)
    e = np.concatenate([I[:, [0, 1]], I[:, [1, 2]], I[:, [2, 0]]]); tri = np.tile(np.arange(len(I)), 3)
    key = np.sort(e, 1); order = np.lexsort((key[:, 1], key[:, 0])); key = key[order]; tri = tri[order]
    same = (key[1:] == key[:-1]).all(1)
    first = np.nonzero(same)[0]
    alone = np.ones(len(key), bool); alone[first] = False; alone[first + 1] = False
    # edges used three times or more do not occur in these meshes
    return key[first, 0], key[first, 1], tri[first], tri[first + 1]


def edge_rays(sc, count, seed):
    """rays from random origins aimed at midpoints of shared edges whose two triangles face the origin the same way (no silhouettes)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = sc.meshes[0]
    ei, ej, ta, tb = shared_edges(m)
    pick = rng.integers(0, len(ei), count)
    P = m.positions.astype(np.float64)
    mid = 0.5 * (P[ei[pick]] + P[ej[pick]])
    lo, hi = scenes.scene_bounds(sc)
    org = (lo + hi) / 2 + (rng.random((count, 3)) * 2 - 1) * (hi - lo) * 0.6

    def facing(t):                                     # cosine between the triangle normal and the direction back to the origin
        I = m.indices[t]
        n = np.cross(P[I[:, 1]] - P[I[:, 0]], P[I[:, 2]] - P[I[:, 0]])
        w = org - mid
        return (n * w).sum(1) / (np.linalg.norm(n, axis=1) * np.linalg.norm(w, axis=1))
    fa, fb = facing(ta[pick]), facing(tb[pick])
    keep = (fa * fb > 0) & (np.abs(fa) > 0.05) & (np.abs(fb) > 0.05)               # same side, not grazing (a grazing ray is moved off the edge by the float rounding of its own direction)
    rays = np.zeros(int(keep.sum()), scenes.RAY_DTYPE)
    rays["origin"] = org[keep].astype(np.float32)
    rays["direction"] = (mid[keep].astype(np.float32) - rays["origin"])
    rays["tmax"] = np.inf
    return rays


def small_scenes():
    yield "cornell", scenes.cornell_box(32, 32)
    yield "terrain", scenes.fractal_terrain(n=24, width=32, height=32)
    yield "spheres", scenes.sphere_field(num_spheres=6, subdiv=2, width=32, height=32)


@pytest.mark.parametrize("name,sc", list(small_scenes()), ids=[n for n, _ in small_scenes()])
def test_rays_aimed_at_shared_edges_are_lost_only_inside_the_stated_band(name, sc):
    orc = OracleScene(sc)
    rays = edge_rays(sc, 30000, 41)
    assert len(rays) > 10000
    # the direction is (edge midpoint - origin), unnormalised: the targeted edge is at t = 1
    t64, p64, mg, _ = f64_exhaustive(sc.meshes[0], rays, eps=EPS_IN)
    assert np.isfinite(t64).all() and (t64 <= 1 + T_REL).all(), "construction: every ray passes through an edge of the mesh at t = 1"
    for brute in (False, True):
        h, _ = orc.trace_closest(rays, brute=brute)
        t32 = h["t"].astype(np.float64)
        lost = (h["prim"] == MISS) | (t32 > 1 + T_REL)          # fell between the two triangles: nothing, or something BEHIND the edge, is reported
        print(f"{name}: {lost.mean():.4f} of the rays aimed exactly at a shared edge fall between its two triangles (brute={brute})")
        assert lost.mean() < 0.15
        # exactly ON the edge in float64 terms (|margin| ~ 1e-8): inside the stated band by construction, which is the point
        print(f"   largest |margin| of a lost ray: {np.abs(mg[lost]).max() if lost.any() else 0:.2e}")
        assert (np.abs(mg[lost]) < EPS_BAND).all()
        # whatever is reported in FRONT of the edge is a real surface
        closer = ~lost & (t32 < 1 - T_REL)
        assert (t64[closer] <= t32[closer] * (1 + T_REL) + T_ABS).all(), f"{name}: phantom occluders in front of the targeted edge"


@pytest.mark.parametrize("name,sc", list(small_scenes()), ids=[n for n, _ in small_scenes()])
def test_random_and_camera_rays_agree_with_float64_exhaustive_search(name, sc):
    orc = OracleScene(sc)
    lo, hi = scenes.scene_bounds(sc)
    rays = np.concatenate([scenes.incoherent_rays(lo, hi, 12000, 42), scenes.camera_rays(sc, 6000, 43)])
    t64, p64, mg, _ = f64_exhaustive(sc.meshes[0], rays, eps=EPS_IN)
    h, _ = orc.trace_closest(rays)
    hit32 = h["prim"] != MISS
    inside = np.isfinite(t64) & (mg > EPS_IN)
    # (1) nothing well inside a triangle is missed or attributed to another surface
    assert hit32[inside].all(), f"{name}: {int((~hit32[inside]).sum())} interior hits missed"
    rel = np.abs(h["t"][inside].astype(np.float64) - t64[inside]) / np.abs(t64[inside])
    # a hit within the EPS band of another, closer triangle is allowed to win; count only rays that end up FARTHER
    farther = h["t"][inside].astype(np.float64) > t64[inside] * (1 + T_REL) + T_ABS
    assert not farther.any(), f"{name}: {int(farther.sum())} rays passed through a surface float64 hits in its interior"
    same = h["prim"][inside] == p64[inside]
    assert same.mean() > 0.999 and (np.abs(h["t"][inside].astype(np.float64) - t64[inside])[same] <= T_REL * np.abs(t64[inside][same]) + T_ABS).all()
    # (1b) whatever float64 hits and the oracle loses lies inside the stated band around an edge
    t64n, _, mgn, _ = f64_exhaustive(sc.meshes[0], rays, eps=1e-12)
    lost = np.isfinite(t64n) & ~hit32
    assert (mgn[lost] < EPS_BAND).all(), f"{name}: a hit with margin {mgn[lost].max():.2e} was lost"
    # (2) no phantom hits: where float64 (edges widened by EPS_IN) sees nothing, the oracle sees nothing
    phantom = hit32 & ~np.isfinite(t64)
    assert not phantom.any(), f"{name}: {int(phantom.sum())} hits float64 cannot find"


def test_rays_aimed_exactly_at_shared_vertices_leak_rate_is_bounded():
    """Steep rays from above a height field through an interior grid vertex to the last float bit: no silhouettes, so a miss can only be a
    leak between the six triangles of the fan.  Moller-Trumbore decides each of them with its own roundings; the rate is printed and bounded.
    (The projected watertight form built in round 2 passed this construction with 0 of 30,000 lost.)"""
    sc = scenes.fractal_terrain(n=24, width=32, height=32, with_light=False)
    orc = OracleScene(sc)
    m = sc.meshes[0]
    rng = np.random.Generator(np.random.PCG64(44))
    n = 24
    vi = rng.integers(1, n, 30000) * (n + 1) + rng.integers(1, n, 30000)         # interior grid vertices
    tgt = m.positions[vi].astype(np.float64)
    height = 30 + 40 * rng.random(len(vi))
    org = tgt + np.stack([(rng.random(len(vi)) - 0.5) * 0.5 * height, (rng.random(len(vi)) - 0.5) * 0.5 * height, height], 1)   # within ~14 degrees of vertical
    rays = np.zeros(len(vi), scenes.RAY_DTYPE)
    rays["origin"] = org.astype(np.float32); rays["direction"] = m.positions[vi] - rays["origin"]; rays["tmax"] = np.inf
    h, _ = orc.trace_closest(rays)
    leak = float((h["prim"] == MISS).mean())
    print(f"exact-vertex rays falling through the fan: {leak:.4f}")
    assert leak < 0.30                                      # measured 0.21: six triangles, each decided with its own roundings
    # rays that merely pass NEAR the vertex (1e-3 of the cell size away, off the grid and diagonal directions) are never lost
    rays["direction"][:, 0] += np.float32(4e-3); rays["direction"][:, 1] += np.float32(1.7e-3)
    h2, _ = orc.trace_closest(rays)
    assert not (h2["prim"] == MISS).any()

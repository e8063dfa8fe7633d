"""Shared helpers for the parity tests: small versions of BASELINE.json's five configs and comparison utilities."""
import numpy as np

from foundation_b200 import scenes

SMALL_SCENES = {
    "cornell": lambda: scenes.cornell_box(128, 128),
    "spheres": lambda: scenes.sphere_field(num_spheres=60, subdiv=2, width=160, height=90),
    "terrain": lambda: scenes.fractal_terrain(n=160, width=160, height=90),
    "instanced": lambda: scenes.instanced_patches(num_instances=150, patch=12, width=160, height=90),
}


def multi_mesh_scene():
    """Three meshes, no instance list (implicit identity instances): the Cornell room, and its two boxes as separate meshes."""
    sc = scenes.cornell_box(96, 96)
    m = sc.meshes[0]
    parts = [(0, 10), (10, 20), (20, 30), (30, 32)]      # room, tall box, short box, light
    meshes = []
    for a, b in [(0, 10), (10, 20), (20, 32)]:
        idx = m.indices[a:b]
        used, inv = np.unique(idx.reshape(-1), return_inverse=True)
        meshes.append(scenes.Mesh(np.ascontiguousarray(m.positions[used]), np.ascontiguousarray(inv.reshape(-1, 3).astype(np.uint32)),
                                  np.ascontiguousarray(m.material_ids[a:b])))
    return scenes.Scene("cornell_3_meshes", meshes, sc.materials, None, sc.view, sc.proj, sc.width, sc.height)


def ray_mix(scene, n_each=4096, seed=4):
    lo, hi = scenes.scene_bounds(scene)
    parts = [scenes.incoherent_rays(lo, hi, n_each, seed), scenes.camera_rays(scene, n_each, seed + 1)]
    if scene.instances is None and len(scene.meshes) == 1:
        parts.append(scenes.stress_rays(scene, n_each, seed + 2))
    # bounded segments exercise tmin / tmax
    seg = scenes.incoherent_rays(lo, hi, n_each, seed + 3)
    ext = float(np.max(hi - lo))
    seg["tmin"] = 0.05 * ext
    seg["tmax"] = 0.4 * ext
    parts.append(seg)
    return np.concatenate(parts)


def assert_hits_equal(a_hits, a_inst, b_hits, b_inst, what=""):
    bad_id = int((a_hits["prim"] != b_hits["prim"]).sum()) + (int((a_inst != b_inst).sum()) if a_inst is not None and b_inst is not None else 0)
    ta = a_hits["t"].view(np.uint32).astype(np.int64); tb = b_hits["t"].view(np.uint32).astype(np.int64)
    max_ulp = int(np.abs(ta - tb).max()) if len(ta) else 0
    assert bad_id == 0, f"{what}: {bad_id} hit-ID mismatches"
    assert max_ulp <= 2, f"{what}: hit t differs by {max_ulp} ulp (north_star allows 2)"
    return max_ulp


def rmse(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt(np.mean(d * d)))

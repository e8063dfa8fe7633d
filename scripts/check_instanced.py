"""BVH vs exhaustive search on the device for the full instanced scene (config 4), incoherent + camera + surface-started rays."""
import sys
import numpy as np
sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

sc = scenes.instanced_patches()
n = 1 << 15
for leaf in (1, 3):
    tr = pt.PathTracer(sc.width, sc.height, background=sc.background, max_leaf_tris=leaf)
    bs = tr.load(sc)
    lo, hi = np.asarray(bs.scene_lo[:]), np.asarray(bs.scene_hi[:])
    cam = scenes.camera_rays(sc, n, 3)
    inc = scenes.incoherent_rays(lo, hi, n, 4)
    # secondary rays: start at camera hit points, offset along +z, random upward directions (like diffuse bounces off the patches)
    ch, _ = tr.trace_closest(cam)
    ok = ch["prim"] != 0xFFFFFFFF
    p = cam["origin"][ok] + cam["direction"][ok] * ch["t"][ok, None]
    rng = np.random.default_rng(9)
    d = rng.normal(size=(ok.sum(), 3)); d[:, 2] = np.abs(d[:, 2]) * 0.3; d /= np.linalg.norm(d, axis=1, keepdims=True)
    sec = np.zeros(ok.sum(), scenes.RAY_DTYPE); sec["origin"] = (p + np.asarray([0, 0, 1.5e-3])).astype(np.float32); sec["direction"] = d.astype(np.float32); sec["tmax"] = np.inf
    rays = np.concatenate([cam, inc, sec])
    tr.rays_upload(rays)
    tr.rays_trace_closest(); gh, gi = tr.rays_download_hits()
    tr.rays_trace_brute(); bh, bi = tr.rays_download_hits()
    print("brute ms", tr.stats().last_ms)
    bad = (gh["prim"] != bh["prim"]) | (gi != bi)
    print(f"leaf {leaf}: rays {len(rays)} hitfrac {float((bh['prim'] != 0xFFFFFFFF).mean()):.3f} mismatches {int(bad.sum())} (cam {int(bad[:n].sum())}, inc {int(bad[n:2*n].sum())}, sec {int(bad[2*n:].sum())})")
    for k in np.nonzero(bad)[0][:8]:
        print("  ray", k, "bvh", gh[k], gi[k], "brute", bh[k], bi[k], "o", rays["origin"][k], "d", rays["direction"][k])
    tr.close()

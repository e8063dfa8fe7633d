"""Quick GPU probe: build a scene, trace an incoherent ray set, print build time and Mrays/s (dev tool, not the bench)."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="terrain")
ap.add_argument("--n", type=int, default=2236)
ap.add_argument("--rays", type=int, default=1 << 24)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--bounces", type=int, default=8)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--leaf", type=int, default=0)
ap.add_argument("--sort", type=int, default=0, help="pre-sort rays on the host by a Morton key of the origin with this many bits per axis (experiment)")
a = ap.parse_args()
t0 = time.time()
if a.scene == "terrain":
    sc = scenes.fractal_terrain(n=a.n)
elif a.scene == "spheres":
    sc = scenes.sphere_field()
elif a.scene == "instanced":
    sc = scenes.instanced_patches()
else:
    sc = scenes.cornell_box()
print(f"scene {sc.name}: {sc.num_triangles} tris ({sc.effective_triangles} effective) generated in {time.time() - t0:.1f}s", flush=True)
tr = pt.PathTracer(sc.width, sc.height, background=sc.background, flags=a.flags, max_leaf_tris=a.leaf)
t0 = time.time()
bs = tr.load(sc)
print(f"commit wall {time.time() - t0:.2f}s  build_ms={bs.build_ms:.2f} sort_ms={bs.sort_ms:.2f} nodes8={bs.num_nodes8} bytes={bs.device_bytes / 1e6:.1f}MB", flush=True)
lo, hi = np.asarray(bs.scene_lo[:]), np.asarray(bs.scene_hi[:])
rays = scenes.incoherent_rays(lo, hi, a.rays)
if a.sort:
    b = a.sort
    q = np.clip(((rays["origin"].astype(np.float64) - lo) / (hi - lo) * (1 << b)).astype(np.int64), 0, (1 << b) - 1)
    key = np.zeros(len(rays), np.int64)
    for bit in range(b):
        for ax in range(3):
            key |= ((q[:, ax] >> bit) & 1) << (3 * bit + (2 - ax))
    octant = (rays["direction"][:, 0] < 0).astype(np.int64) * 4 + (rays["direction"][:, 1] < 0) * 2 + (rays["direction"][:, 2] < 0)
    key = (key << 3) | octant
    t0 = time.time(); order = np.argsort(key, kind="stable"); rays = rays[order]; print(f"host sort {time.time() - t0:.1f}s", flush=True)
tr.rays_upload(rays)
for r in range(a.reps):
    tr.rays_trace_closest()
    ms = tr.stats().last_ms
    print(f"closest: {ms:.2f} ms  {a.rays / ms / 1e3:.1f} Mrays/s", flush=True)
tr.rays_trace_any()
ms = tr.stats().last_ms
print(f"any: {ms:.2f} ms  {a.rays / ms / 1e3:.1f} Mrays/s", flush=True)
h, _ = tr.rays_download_hits()
print("hit fraction", float((h['prim'] != 0xFFFFFFFF).mean()))
if a.spp:
    tr.render(0, 1, a.bounces)
    t0 = time.time()
    tr.render(1, a.spp, a.bounces)
    st = tr.stats()
    print(f"render {a.spp} spp: {st.last_ms:.1f} ms -> {a.spp / st.last_ms * 1e3:.2f} spp/s; rays ext {st.rays_extend} shadow {st.rays_shadow} -> "
          f"{(st.rays_extend + st.rays_shadow) / st.last_ms / 1e3:.1f} Mrays/s; launches {st.kernel_launches}", flush=True)
    if a.flags & pt.FLAG_STAGE_TIMING:
        print("stages (ms): " + ", ".join(f"{n} {v:.2f}" for n, v in zip(pt.STAGE_NAMES, st.stage_ms)), flush=True)

"""Round-2 A/B probe (dev tool): steady-state build of the 10 M-triangle terrain with a hash of the BVH8, then the three ray sets of
bench.py (headline incoherent set, its hitting subset, surface-started bounce rays) and a short render.  One line per measurement.
FOUNDATION_PT_LIB=<variant.so> selects the library."""
import argparse
import hashlib
import sys

import numpy as np

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2-rays", type=int, default=24)
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--hash", action="store_true")
ap.add_argument("--builds", type=int, default=3)
a = ap.parse_args()
sc = scenes.fractal_terrain()
out = {}
for i in range(a.builds):
    tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
    bs = tr.load(sc)
    out["build_ms"], out["sort_ms"] = round(float(bs.build_ms), 3), round(float(bs.sort_ms), 3)
    if i + 1 < a.builds:
        tr.close()
if a.hash:
    n, t, o = tr.blas_download(0)
    out["bvh_sha"] = hashlib.sha256(n.tobytes() + t.tobytes() + o.tobytes()).hexdigest()[:16]
lo, hi = np.asarray(bs.scene_lo[:]), np.asarray(bs.scene_hi[:])
rays = scenes.incoherent_rays(lo, hi, 1 << a.log2_rays, 4)


def run(label, rs, reps=4):
    tr.rays_upload(rs)
    best = 1e9
    for _ in range(reps):
        tr.rays_trace_closest(); best = min(best, tr.stats().last_ms)
    out[label] = round(len(rs) / best / 1e3, 1)


run("incoherent", rays)
h, _ = tr.rays_download_hits()
out["hits_sha"] = hashlib.sha256(h.tobytes()).hexdigest()[:12]
run("hits_only", rays[h["prim"] != 0xFFFFFFFF])
cam = scenes.camera_rays(sc, 1 << 22, 3)
ch, _ = tr.trace_closest(cam)
run("secondary", scenes.secondary_rays(sc, cam, ch))
tr.rays_upload(rays[: 1 << 22]); tr.rays_trace_any(); out["any"] = round((1 << 22) / tr.stats().last_ms / 1e3, 1)
if a.spp:
    tr.render(0, 1, 8); tr.render(1, a.spp, 8)
    st = tr.stats()
    out["spp_per_s"] = round(a.spp / st.last_ms * 1e3, 1)
    out["frame_sha"] = hashlib.sha256(tr.read_accum().tobytes()).hexdigest()[:12]
print(out, flush=True)

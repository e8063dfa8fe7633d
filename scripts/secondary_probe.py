"""Experiment: do surface-started (secondary) rays trace faster when sorted by origin?  Builds the terrain, shoots camera rays, starts cosine-ish
bounce rays at the hit points and traces them in (a) wavefront order (pixel order, as k_shade's compaction leaves them), (b) shuffled, (c) sorted
by a Morton key of the origin."""
import sys
import numpy as np
sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

sc = scenes.fractal_terrain()
tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
bs = tr.load(sc)
lo, hi = np.asarray(bs.scene_lo[:], np.float64), np.asarray(bs.scene_hi[:], np.float64)
W, H = sc.width, sc.height
# primary rays in pixel order, 4 per pixel
cam = scenes.camera_rays(sc, W * H * 4, 3)
order = np.lexsort((cam["direction"][:, 0], cam["direction"][:, 2]))   # roughly scanline order
cam = cam[order]
h, _ = tr.trace_closest(cam)
ok = h["prim"] != 0xFFFFFFFF
p = cam["origin"][ok].astype(np.float64) + cam["direction"][ok].astype(np.float64) * h["t"][ok, None]
rng = np.random.default_rng(1)
d = rng.normal(size=(ok.sum(), 3)); d[:, 2] = np.abs(d[:, 2]) + 0.2; d /= np.linalg.norm(d, axis=1, keepdims=True)
sec = np.zeros(int(ok.sum()), scenes.RAY_DTYPE)
sec["origin"] = (p + np.asarray([0, 0, 2e-3])).astype(np.float32); sec["direction"] = d.astype(np.float32); sec["tmax"] = np.inf
print("secondary rays:", len(sec))


def run(label, rays):
    tr.rays_upload(rays)
    best = 1e9
    for _ in range(3):
        tr.rays_trace_closest(); best = min(best, tr.stats().last_ms)
    print(f"{label:28s} {best:7.2f} ms  {len(rays) / best / 1e3:8.1f} Mrays/s", flush=True)


run("wavefront (pixel) order", sec)
run("shuffled", sec[rng.permutation(len(sec))])
for bits in (6, 10):
    q = np.clip(((sec["origin"].astype(np.float64) - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    key = np.zeros(len(sec), np.int64)
    for b in range(bits):
        for ax in range(3):
            key |= ((q[:, ax] >> b) & 1) << (3 * b + (2 - ax))
    run(f"sorted by origin ({bits} bits)", sec[np.argsort(key, kind="stable")])

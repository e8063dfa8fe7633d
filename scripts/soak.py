"""Full-size parity soak (BASELINE.json configs 2-4): device BVH vs device exhaustive search, device vs CPU oracle hits,
and 1080p image parity against the CPU oracle.  Writes gpurun_out/soak.json."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402
from oracle.pt_oracle import OracleScene  # noqa: E402

out = {}
for name, make, nbrute in (("config2_sphere_field", scenes.sphere_field, 1 << 20), ("config3_terrain", scenes.fractal_terrain, 1 << 19),
                           ("config4_instanced", scenes.instanced_patches, 1 << 14)):
    t0 = time.time()
    sc = make()
    tr = pt.PathTracer(sc.width, sc.height, seed=5, background=sc.background)
    bs = tr.load(sc)
    orc = OracleScene(sc)
    lo, hi = np.asarray(bs.scene_lo[:]), np.asarray(bs.scene_hi[:])
    n = 1 << 20
    rays = np.concatenate([scenes.incoherent_rays(lo, hi, n // 2, 4), scenes.camera_rays(sc, n // 2, 6)])
    gh, gi = tr.trace_closest(rays)
    oh, oi = orc.trace_closest(rays)
    r = {"triangles": int(bs.num_triangles), "effective_triangles": int(bs.effective_triangles), "rays_vs_oracle": int(n),
         "hit_id_mismatches_vs_oracle": int((gh["prim"] != oh["prim"]).sum() + (gi != oi).sum()),
         "t_bits_differ": int((gh["t"].view(np.uint32) != oh["t"].view(np.uint32)).sum())}
    tr.rays_upload(rays[:nbrute]); tr.rays_trace_brute(); bh, bi = tr.rays_download_hits()
    r["rays_vs_exhaustive"] = int(nbrute); r["exhaustive_ms"] = float(tr.stats().last_ms)
    r["hit_id_mismatches_vs_exhaustive"] = int((gh["prim"][:nbrute] != bh["prim"]).sum() + (gi[:nbrute] != bi).sum())
    tr.render(0, 2, 8)
    g = tr.read_accum()
    o = orc.render(sc.width, sc.height, 5, 0, 2, 8, background=sc.background)
    d = g[..., :3].astype(np.float64) - o[..., :3].astype(np.float64)
    r["image_1080p_2spp_8bounce_rmse"] = float(np.sqrt((d * d).mean())); r["image_pixels_differing"] = int((g != o).any(-1).sum())
    r["seconds"] = round(time.time() - t0, 1)
    out[name] = r
    print(name, r, flush=True)
    tr.close()
json.dump(out, open("gpurun_out/soak.json", "w"), indent=1)

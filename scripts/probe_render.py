"""Render-only A/B probe (dev tool): config-3 terrain, N spp at 1080p, per-stage device times (FOUNDATION_PT_FLAG_STAGE_TIMING)."""
import hashlib
import sys

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "terrain"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = scenes.by_name(name)
out = {}
for timing in (0, pt.FLAG_STAGE_TIMING):
    tr = pt.PathTracer(sc.width, sc.height, background=sc.background, flags=timing)
    tr.load(sc)
    tr.render(0, 1, 8); tr.render(1, spp, 8)
    st = tr.stats()
    if timing:
        out["stages_ms"] = {n: round(v, 2) for n, v in zip(pt.STAGE_NAMES, st.stage_ms)}
    else:
        out["spp_per_s"] = round(spp / st.last_ms * 1e3, 1); out["mrays_per_s"] = round((st.rays_extend + st.rays_shadow) / st.last_ms / 1e3, 1)
        out["frame_sha"] = hashlib.sha256(tr.read_accum().tobytes()).hexdigest()[:12]
    tr.close()
print(name, out, flush=True)

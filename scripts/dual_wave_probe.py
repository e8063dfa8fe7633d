"""Dev tool: spp/s of a 64-spp render with one / two waves in flight and several wave sizes; prints a hash of the frame (must not change)."""
import hashlib
import os
import subprocess
import sys

if len(sys.argv) > 2:
    sys.path.insert(0, ".")
    from foundation_b200 import pt, scenes  # noqa: E402
    sc = scenes.by_name(sys.argv[1])
    tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
    tr.load(sc)
    tr.render(0, 16, 8)
    best = 1e9
    for _ in range(3):
        tr.render(0, 64, 8); best = min(best, tr.stats().last_ms)
    st = tr.stats()
    print(f"{sys.argv[1]} {sys.argv[2]}: {best:.1f} ms  {64 / best * 1e3:.1f} spp/s  rays {st.rays_extend + st.rays_shadow}  frame {hashlib.sha256(tr.read_accum().tobytes()).hexdigest()[:12]}", flush=True)
else:
    scene = sys.argv[1] if len(sys.argv) > 1 else "terrain"
    for dual, ws in ((0, 16), (1, 16), (1, 8), (1, 12), (1, 32)):
        env = dict(os.environ, FOUNDATION_PT_DUAL_WAVE=str(dual), FOUNDATION_PT_WAVE_SAMPLES=str(ws))
        subprocess.run([sys.executable, __file__, scene, f"dual={dual} wave_samples={ws}"], env=env, timeout=200)
